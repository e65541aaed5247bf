#!/bin/bash
# round 2, session 3, final validation on one B200: smoke, the parity tests closest to what changed, a short bench line (with the sweep metric)
mkdir -p gpurun_out
( time timeout 90 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2x_smoke.log 2>&1; tail -4 gpurun_out/r2x_smoke.log
( time timeout 130 python -m pytest tests/test_sigma_gpu.py tests/test_update.py tests/test_svd_gpu.py tests/test_zz_sobject_gpu.py tests/test_dmrg_gpu.py -x -q -m gpu -k "not known_answer and not excited and not 2rdm and not twodm and not correlation and not checkpoint and not spill and not variational and not offload and not sharded and not owner_shards and not join_gpu and not convergence_scheme" ) > gpurun_out/r2x_tests.log 2>&1; tail -4 gpurun_out/r2x_tests.log
( time timeout 150 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-update ) > gpurun_out/r2x_bench_short.json 2> gpurun_out/r2x_bench_short.err; tail -c 2500 gpurun_out/r2x_bench_short.json; tail -3 gpurun_out/r2x_bench_short.err
# if time is left: the complete N2/cc-pVDZ schedule up to the published bond dimension (same command as profiles/r2t_dmrg_n2_ccpvdz_d2000.log)
if [ $SECONDS -lt 235 ]; then
  timeout $((310 - SECONDS)) python scripts/run_dmrg.py n2_ccpvdz 250:1,500:1,1000:1,2000:2 > gpurun_out/r2x_dmrg_n2_ccpvdz_d2000.log 2>&1; cat gpurun_out/r2x_dmrg_n2_ccpvdz_d2000.log
fi
echo "elapsed $SECONDS s"
