#!/bin/bash
# round 2, session 3: where a Split spends its time (B2_TIMING lines of split_host / dev_svd_batch), N2/cc-pVDZ D = 1000 -> 2000
mkdir -p gpurun_out
B2_TIMING=1 timeout 120 python scripts/run_dmrg.py n2_ccpvdz 1000:1,2000:1 2> gpurun_out/r2w_timing.err > gpurun_out/r2w_n2.log
grep "dev_svd_batch\|split_host\|b2_dmrg_sweep" gpurun_out/r2w_timing.err > gpurun_out/r2w_split_timing.log
rm -f gpurun_out/r2w_timing.err
cat gpurun_out/r2w_n2.log; grep b2_dmrg_sweep gpurun_out/r2w_split_timing.log; awk 'NR%2==0' gpurun_out/r2w_split_timing.log | tail -44
