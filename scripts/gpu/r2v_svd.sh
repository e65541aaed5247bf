#!/bin/bash
# round 2, session 3: block-Jacobi SVD — parity (LAPACK, reference Split traces, sweeps), then N2/cc-pVDZ sweeps with the scalar and the block kernel
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests/test_svd_gpu.py tests/test_zz_sobject_gpu.py tests/test_trace.py tests/test_checkpoint_interop_gpu.py -x -q -m gpu ) > gpurun_out/r2v_tests.log 2>&1
tail -4 gpurun_out/r2v_tests.log
( time timeout 120 python -m pytest tests/test_dmrg_gpu.py -x -q -k "sweep_energies or known_answer or excited" ) > gpurun_out/r2v_tests2.log 2>&1
tail -4 gpurun_out/r2v_tests2.log
B2_SVD_BLOCK=0 timeout 150 python scripts/run_dmrg.py n2_ccpvdz 500:1,1000:1,2000:1 > gpurun_out/r2v_n2_svd_scalar.log 2>&1
timeout 150 python scripts/run_dmrg.py n2_ccpvdz 500:1,1000:1,2000:1 > gpurun_out/r2v_n2_svd_block.log 2>&1
echo "--- scalar"; cat gpurun_out/r2v_n2_svd_scalar.log
echo "--- block"; cat gpurun_out/r2v_n2_svd_block.log
