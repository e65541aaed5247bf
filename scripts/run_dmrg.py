"""Runs a whole DMRG calculation (own random MPS, PreSolve, scheduled two-site sweeps) on the GPU through the C ABI and prints the
per-half-sweep energies and the time spent per phase.  python scripts/run_dmrg.py tetracene_ppp 100:2,300:2,600:3 [--rtol 1e-5]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chemps2_b200 import api, workloads  # noqa: E402


def run(name, schedule, rtol=1e-5, noise=0.0, seed=12345, device=0, spill=False, log=print):
    w = workloads.get(name)
    ctx = w.context(device)
    ctx.bk_init(schedule[0][0])
    d = api.DMRG(ctx)
    d.set_spill(spill)
    d.random_mps(seed)
    L = w.L
    t0 = time.time()
    for i in range(L - 2):
        d.update(i, True)                       # DMRG::PreSolve (DMRG.cpp:257-266)
    log(f"presolve {time.time() - t0:.2f} s")
    out, change, seen = [], False, (0, 0)
    for entry in schedule:
        D, nsweeps = entry[0], entry[1]
        if len(entry) > 2:
            noise = entry[2]
        for _ in range(nsweeps):
            for to_right in (False, True):
                d.timers(reset=True)
                t0 = time.time()
                e, dw = d.sweep(to_right, rtol, noise, D, change)
                dt = time.time() - t0
                tm = d.timers()
                hits, misses = d.plan_cache_stats()
                tm["plan_hits"], tm["plan_misses"] = hits - seen[0], misses - seen[1]
                seen = (hits, misses)
                out.append(dict(D=D, to_right=to_right, energy=e, max_discarded=dw, seconds=dt, **tm))
                log(f"D={D:5d} {'->' if to_right else '<-'} E = {e:.10f}  w = {dw:.2e}  {dt:7.2f} s  (plan {tm['plan_s']:.2f} solve {tm['solve_s']:.2f} "
                    f"split {tm['split_s']:.2f} update {tm['update_s']:.2f}; {tm['n_matvec']} sigma builds; plans re-used {tm['plan_hits']}, built {tm['plan_misses']})")
                change = True
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("schedule", help="D:sweeps[:noise],D:sweeps[:noise],...")
    ap.add_argument("--rtol", type=float, default=1e-5)
    ap.add_argument("--noise", type=float, default=0.0)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    sched = [tuple([int(x.split(":")[0]), int(x.split(":")[1])] + [float(v) for v in x.split(":")[2:3]]) for x in a.schedule.split(",")]
    res = run(a.workload, sched, a.rtol, a.noise, a.seed)
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)
