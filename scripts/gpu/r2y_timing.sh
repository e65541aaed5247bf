#!/bin/bash
# round 2, session 3: phase breakdown of N2/cc-pVDZ half sweeps after the host block cache (B2_TIMING lines of b2_dmrg_sweep)
mkdir -p gpurun_out
B2_TIMING=1 timeout 100 python scripts/run_dmrg.py n2_ccpvdz 1000:1,2000:1 2> gpurun_out/r2y_timing.err > gpurun_out/r2y_n2.log
grep "b2_dmrg_sweep" gpurun_out/r2y_timing.err > gpurun_out/r2y_sweep_timing.log
grep "b2_heff_create\|b2_update_create\|compile_sigma\|build_sigma_plan" gpurun_out/r2y_timing.err | tail -120 > gpurun_out/r2y_plan_timing.log
rm -f gpurun_out/r2y_timing.err
cat gpurun_out/r2y_n2.log gpurun_out/r2y_sweep_timing.log; tail -30 gpurun_out/r2y_plan_timing.log
