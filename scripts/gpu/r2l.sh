#!/bin/bash
# full GPU suite + ncu --set full of the HBM-bound kernels (second capture: Davidson BLAS-1, pre-sum, block scaling, diagonal, mixing, SVD)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=12 ) > gpurun_out/r2l_gpu_tests.log 2>&1
tail -22 gpurun_out/r2l_gpu_tests.log
NCU="ncu --clock-control none"
timeout 600 $NCU --set full -k regex:"k_multi_dot|k_multi_axpy|k_ritz|k_precond|k_diag|k_presum|k_scale|k_lincomb|k_jacobi" -c 48 -o gpurun_out/r2l_hbm python scripts/hbm_kernels.py > gpurun_out/r2l_hbm.log 2>&1
echo "hbm rc $?"
timeout 400 $NCU --set full -k regex:"k_mix_flat|k_axpy_tiles" -c 8 -o gpurun_out/r2l_mix python scripts/update_only.py 2000 > gpurun_out/r2l_mix.log 2>&1
echo "mix rc $?"
ls -la gpurun_out | tail -6
