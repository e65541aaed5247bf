#!/bin/bash
mkdir -p gpurun_out
for hyb in 0 1; do
  B2_KHYBRID=$hyb timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sweep --no-update > gpurun_out/r2j_hyb$hyb.json 2> gpurun_out/r2j_hyb$hyb.err
  python -c "
import json
l=json.load(open('gpurun_out/r2j_hyb$hyb.json')); print('hybrid', $hyb, l['ms_per_step'], l['roofline']['frac'], l['roofline']['executed_frac'], l['plan']['sigma_norm'])"
done
B2_KHYBRID=1 timeout 600 python -m pytest tests/test_sigma_gpu.py tests/test_update.py tests/test_zz_sobject_gpu.py -m gpu -q -x 2>&1 | tail -3
