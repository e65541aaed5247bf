#!/bin/bash
# the remaining reference tests on the drop-in library: 6-8 (CASSCF), 10, 11 (3-/4-RDM), 13 — each under its own timeout
mkdir -p gpurun_out
cd /tmp
for n in 6 7 8 10 11; do
  exe=/root/repo/dropin/_build/test$n
  [ -x $exe ] || continue
  t0=$(date +%s)
  CHEMPS2_B200_VERBOSE=1 OMP_NUM_THREADS=16 OPENBLAS_NUM_THREADS=1 timeout 170 $exe > /root/repo/gpurun_out/r2q_test$n.out 2> /root/repo/gpurun_out/r2q_test$n.err
  rc=$?
  echo "test$n rc $rc $(( $(date +%s) - t0 )) s | $(grep -E 'Did test' /root/repo/gpurun_out/r2q_test$n.out | tail -1) | $(grep 'drop-in:' /root/repo/gpurun_out/r2q_test$n.err | tail -1 | cut -c1-160)"
  tail -c 20000 /root/repo/gpurun_out/r2q_test$n.out > /root/repo/gpurun_out/r2q_test$n.tail; rm -f /root/repo/gpurun_out/r2q_test$n.out
done
