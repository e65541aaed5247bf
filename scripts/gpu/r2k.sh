#!/bin/bash
mkdir -p gpurun_out
B2_TIMING=1 timeout 300 python scripts/update_only.py 2>&1 | grep -E "b2_update_run|^update|b2_update_create" | tail -5
timeout 900 python -m pytest tests/test_update.py tests/test_large_vs_reference_gpu.py tests/test_davidson_rc_gpu.py tests/test_dmrg_gpu.py tests/test_twodm.py tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -4
