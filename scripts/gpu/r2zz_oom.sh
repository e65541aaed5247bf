#!/bin/bash
# round 2, session 3, the last GPU seconds: out-of-memory / spill paths of the sweep driver with the plan prefetch and the new pool policy
mkdir -p gpurun_out
timeout 44 python -m pytest tests/test_dmrg_gpu.py -x -q -m gpu -k "out_of_memory or with_spill" > gpurun_out/r2zz_tests.log 2>&1; tail -3 gpurun_out/r2zz_tests.log
