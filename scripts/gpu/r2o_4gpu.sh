#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2o_bench_4gpu.json 2> gpurun_out/r2o_bench_4gpu.err
python - <<'PY'
import json
b = json.loads([l for l in open("gpurun_out/r2o_bench_4gpu.json") if l.startswith("{")][-1])
print("N=4:", b["value"], "ms/step", b["ms_per_step"], "speedup vs 2074.4:", 2074.39 / b["ms_per_step"], "norm", b["plan"]["sigma_norm"])
print(json.dumps(b.get("sweep"))[:1500])
PY
tail -3 gpurun_out/r2o_bench_4gpu.err
