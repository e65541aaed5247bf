#!/bin/bash
# round 2, session 3: plan prefetch + cache key — parity subset, then N2/cc-pVDZ sweeps with the prefetch off and on
mkdir -p gpurun_out
( time timeout 240 python -m pytest tests/test_dmrg_gpu.py -x -q -k "prefetch or plan_cache or sweep_energies" ) > gpurun_out/r2u_tests.log 2>&1
tail -4 gpurun_out/r2u_tests.log
B2_PLAN_PREFETCH=0 timeout 150 python scripts/run_dmrg.py n2_ccpvdz 250:1,500:1,1000:2 > gpurun_out/r2u_n2_prefetch_off.log 2>&1
timeout 150 python scripts/run_dmrg.py n2_ccpvdz 250:1,500:1,1000:2 > gpurun_out/r2u_n2_prefetch_on.log 2>&1
echo "--- off"; cat gpurun_out/r2u_n2_prefetch_off.log
echo "--- on"; cat gpurun_out/r2u_n2_prefetch_on.log
