#!/bin/bash
# the two arms exactly as the driver launches them at N = 1
mkdir -p gpurun_out
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2n_ref.json 2> gpurun_out/r2n_ref.err
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
tail -4 gpurun_out/r2n_ref.err; tail -4 gpurun_out/r2n_bench.err
python - <<'PY'
import json
r = json.loads([l for l in open("gpurun_out/r2n_ref.json") if l.startswith("{")][-1])
b = json.loads([l for l in open("gpurun_out/r2n_bench.json") if l.startswith("{")][-1])
print("reference:", r["value"], "builds/s, steps", r["steps"], "ms/step", r["ms_per_step"])
print("ours     :", b["value"], "e2e", b["e2e"]["value"], "ms/step", b["ms_per_step"], "frac", b["roofline"]["frac"], "exec", b["roofline"]["executed_frac"], "traffic", b["roofline"]["traffic"])
print("ratio e2e:", b["e2e"]["value"] / r["value"], "parity", b.get("parity_vs_reference"), "cpu", b.get("cpu_baseline", {}).get("value"))
print("update   :", json.dumps(b.get("operator_update"))[:900])
print("sweep    :", json.dumps(b.get("sweep"))[:600])
PY
