#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sweep > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python -c "
import json
l=json.load(open('gpurun_out/r2i_bench.json')); print(l['ms_per_step'], l['roofline']['frac'], l['roofline']['executed_frac'], l['plan']['sigma_norm']); print(json.dumps(l['operator_update']))"
B2_TIMING=1 timeout 300 python scripts/update_only.py 2>&1 | grep -E "b2_update_run|^update" | tail -4
timeout 900 python -m pytest tests/test_sigma_gpu.py tests/test_update.py tests/test_large_vs_reference_gpu.py tests/test_twodm.py tests/test_zz_sobject_gpu.py -m gpu -q -x 2>&1 | tail -4
