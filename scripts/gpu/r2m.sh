#!/bin/bash
# ncu --set full of the HBM-bound kernels and of the mixing kernels; the reports are reduced to their raw-page CSV on the box
# (gpurun brings back at most 64 MiB)
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 600 $NCU --set full -k regex:"k_multi_dot|k_multi_axpy|k_ritz|k_precond|k_diag|k_presum|k_scale|k_lincomb|k_jacobi" -c 48 -o /tmp/r2m_hbm python scripts/hbm_kernels.py > gpurun_out/r2m_hbm.log 2>&1
echo "hbm rc $?"
ncu -i /tmp/r2m_hbm.ncu-rep --page raw --csv > gpurun_out/r2m_hbm_raw.csv 2>/dev/null
timeout 400 $NCU --set full -k regex:"k_mix_flat|k_axpy_tiles" -c 8 -o /tmp/r2m_mix python scripts/update_only.py 2000 > gpurun_out/r2m_mix.log 2>&1
echo "mix rc $?"
ncu -i /tmp/r2m_mix.ncu-rep --page raw --csv > gpurun_out/r2m_mix_raw.csv 2>/dev/null
gzip -f gpurun_out/r2m_*_raw.csv
ls -la gpurun_out
