"""Checksums of the compiled device work lists (sigma plan + both passes of the update plan) of several workloads, built on the CPU
(planning-only contexts).  Used to prove that a change of the host-side plan builder leaves every list bit-identical:
   B2_PLAN_THREADS=8 python scripts/plan_hash_cpu.py > before.txt ; <change> ; ... > after.txt ; diff before.txt after.txt
(the segmentation of a plan depends on the number of planner threads, so compare runs with the same B2_PLAN_THREADS)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from chemps2_b200 import api, workloads  # noqa: E402
from chemps2_b200._lib import Worklists, check, lib  # noqa: E402
from cpu_check import worklist_digest as digest  # noqa: E402

CASES = [("tiny", 40, "flat", None), ("n2_ccpvdz", 300, "gauss", 5), ("n2_ccpvdz", 600, "gauss", 13), ("n2_ccpvdz", 400, "flat", 13),
         ("synth40", 600, "gauss", 19), ("synth40", 400, "flat", 10), ("tetracene", 800, "gauss", 8), ("n2_ccpvdz", 1000, "gauss", 20)]
if len(sys.argv) > 1:
    CASES = CASES[:int(sys.argv[1])]
for name, D, dist, site in CASES:
    w = workloads.get(name, D=D)
    site = w.site if site is None else site
    ctx = w.context(-1)
    w.apply_distribution(ctx, dist)
    left = api.OpSet(ctx, site, True) if site > 0 else None
    right = api.OpSet(ctx, site + 2, False) if site < w.L - 2 else None
    heff = api.Heff(ctx, site, left, right)
    wl = Worklists()
    check(lib.b2_heff_worklists(heff.h, C.byref(wl)))
    st = heff.stats()
    line = f"{name} D={D} {dist} site={site}: sigma {digest(wl)} terms {st['terms']:.0f} tiles {st['tiles']:.0f} flops_exec {st['flops_exec']:.6e}"
    for mr in (True, False):
        idx = site if mr else site + 1
        old = left if mr else right
        new = api.OpSet(ctx, idx + 1 if mr else idx, mr)
        upd = api.Update(ctx, idx, mr, old, new)
        for p in (0, 1):
            uw = Worklists()
            check(lib.b2_update_worklists(upd.h, p, C.byref(uw)))
            line += f" upd{'R' if mr else 'L'}{p} {digest(uw)}"
    print(line, flush=True)
