#!/bin/bash
# kernel variants 6-8 (8-warp CTAs), the bench line with the operator-update metric, update / large-D / sweep tests
mkdir -p gpurun_out
for v in 6 7 8; do
  B2_KVARIANT=$v python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-sweep > gpurun_out/r2e_var$v.json 2> gpurun_out/r2e_var$v.err
  python -c "
import json
l=json.load(open('gpurun_out/r2e_var$v.json'))
print($v, l['ms_per_step'], l['roofline']['kernel_ms_per_sigma_build'], l['plan']['sigma_norm'])"
done
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
python -c "
import json
l=json.load(open('gpurun_out/r2e_bench.json')); print(json.dumps(l['operator_update'])); print(json.dumps(l['sweep'])[:2500])"
python -m pytest tests/test_update.py tests/test_large_vs_reference_gpu.py tests/test_dmrg_gpu.py tests/test_sigma_gpu.py -m gpu -q -x 2>&1 | tail -5
