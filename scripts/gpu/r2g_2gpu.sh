#!/bin/bash
# 2 x B200: the sharded sweep test (NCCL) and the bench line at N = 2 (sigma builds with LPT owner groups + the whole-sweep metric)
mkdir -p gpurun_out
python -m pytest tests/test_dmrg_multigpu.py -m gpu -q -x 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g_bench_2gpu.json 2> gpurun_out/r2g_bench_2gpu.err
tail -c 3000 gpurun_out/r2g_bench_2gpu.json; tail -5 gpurun_out/r2g_bench_2gpu.err
