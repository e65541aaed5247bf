#!/bin/bash
# The regimes next to the default bench shape (one JSON line each, CPU reference measured in the same run at the same size):
#   flat (first-sweep) sector tables, the D2h N2 shape, 60 orbitals, 18 orbitals at D=3000, D=6000; then the whole-sweep metric with the
#   unmodified reference's DMRG::Solve on the host cores of the same box (--sweep-ref).
mkdir -p gpurun_out
run() {  # name, args...
  local name=$1; shift
  timeout 900 python bench.py --steps 3 --warmup 3 --no-sweep "$@" > gpurun_out/r2h_$name.json 2> gpurun_out/r2h_$name.err
  echo "$name rc $?"; python - <<PY
import json
try:
    l = json.load(open("gpurun_out/r2h_$name.json"))
    cb, u = l.get("cpu_baseline", {}), l.get("operator_update", {})
    print("  ", l["config"]["workload"][:60], "| ms/build", round(l["ms_per_step"], 2), "| TF/s alg", round(l["roofline"]["achieved"], 2), "frac", round(l["roofline"]["frac"], 3),
          "exec frac", l["roofline"]["executed_frac"] and round(l["roofline"]["executed_frac"], 3), "| cpu builds/s", cb.get("value"), "| parity", l.get("parity_vs_reference", {}).get("max_rel_err"),
          "| update ms", u.get("ms_per_update"), "ref s", u.get("cpu_reference", {}).get("seconds_per_update"))
except Exception as e:
    print("  failed:", e)
PY
}
run synth40_flat --workload synth40 --dist flat
run n2_flat --workload n2 --dist flat
run n2_gauss --workload n2 --dist gauss
run tetracene_gauss --workload tetracene --dist gauss
run synth60_gauss --workload synth60 --dist gauss
run synth40_d6000 --workload synth40 --D 6000 --dist gauss --no-update
timeout 1500 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-update --sweep-ref > gpurun_out/r2h_sweepref.json 2> gpurun_out/r2h_sweepref.err
python -c "
import json
l=json.load(open('gpurun_out/r2h_sweepref.json')); print(json.dumps(l['sweep'])[:3000])"
