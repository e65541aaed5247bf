#!/bin/bash
# last validation of the round: smoke(), the whole GPU suite, one default bench line
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2s_smoke.log 2>&1; tail -3 gpurun_out/r2s_smoke.log
( time timeout 1300 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r2s_gpu_tests.log 2>&1
tail -16 gpurun_out/r2s_gpu_tests.log
( time timeout 600 python bench.py ) > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python - <<'PY'
import json
b = json.loads([l for l in open("gpurun_out/r2s_bench.json") if l.startswith("{")][-1])
print("bench:", b["value"], "e2e", b["e2e"]["value"], "frac", b["roofline"]["frac"], "cpu", b["cpu_baseline"]["value"], "parity", b["parity_vs_reference"]["max_rel_err"])
print("update:", b["operator_update"]["ms_per_update"], b["operator_update"]["tflops_fp64"], b["operator_update"].get("cpu_reference"))
print("sweep:", b["sweep"]["seconds_per_sweep_at_D"], b["sweep"]["total_seconds"], b["sweep"]["energy"])
PY
tail -3 gpurun_out/r2s_bench.err
