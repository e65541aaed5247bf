#!/bin/bash
# round 2, session 3, last call: device pool with best-fit re-use — smoke, allocator / sweep / Split parity, phase breakdown at D = 2000
mkdir -p gpurun_out
timeout 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
timeout 50 python -m pytest tests/test_sigma_gpu.py tests/test_dmrg_gpu.py tests/test_zz_sobject_gpu.py -x -q -m gpu -k "caching_allocator or synthetic_vs_cpu or sweep_energies or split_device" > gpurun_out/r2z_tests.log 2>&1; tail -2 gpurun_out/r2z_tests.log
B2_TIMING=1 timeout 45 python scripts/run_dmrg.py n2_ccpvdz 1000:1,2000:1 2> gpurun_out/r2z_timing.err > gpurun_out/r2z_n2.log
grep "b2_dmrg_sweep" gpurun_out/r2z_timing.err > gpurun_out/r2z_sweep_timing.log; rm -f gpurun_out/r2z_timing.err
cat gpurun_out/r2z_n2.log gpurun_out/r2z_sweep_timing.log; echo "elapsed $SECONDS s"
