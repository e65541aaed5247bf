#!/bin/bash
# ncu evidence of round 2 (every step under its own timeout; memory footprints kept small so that ncu's save/restore between replay
# passes stays cheap).  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 1. operator update at the bench shape: every launch with its device time and DRAM bytes
timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"k_tiles|k_reduce|k_presum|k_zero|k_axpy" --csv \
     --log-file gpurun_out/r2f_update_launches.csv python scripts/update_only.py > gpurun_out/r2f_update.log 2>&1
echo "step 1 rc $?"
# 2. sigma build: launch list of one build
timeout 900 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"k_tiles|k_reduce|k_presum|k_zero|k_diag" --csv \
     --log-file gpurun_out/r2f_sigma_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-sweep --no-update > gpurun_out/r2f_sigma.log 2>&1
echo "step 2 rc $?"
# 3. --set full of the mixing kernel (D = 2000: 1/4 of the memory)
timeout 600 $NCU --set full --import-source on -k regex:"k_axpy_tiles" -c 2 -o gpurun_out/r2f_k_axpy python scripts/update_only.py 2000 > gpurun_out/r2f_full2.log 2>&1
echo "step 3 rc $?"
# 4. --set full of the HBM-bound kernels (long vectors, small operator arenas)
timeout 900 $NCU --set full -k regex:"k_multi_dot|k_multi_axpy|k_ritz|k_precond|k_diag|k_presum|k_reduce|k_scale|k_jacobi|k_lincomb" -c 40 -o gpurun_out/r2f_hbm python scripts/hbm_kernels.py > gpurun_out/r2f_hbm.log 2>&1
echo "step 4 rc $?"
gzip -f gpurun_out/r2f_*launches.csv
ls -la gpurun_out/; tail -3 gpurun_out/r2f_hbm.log; tail -2 gpurun_out/r2f_update.log
