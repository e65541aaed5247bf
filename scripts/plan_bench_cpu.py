"""Host-side plan building timed on the CPU (no GPU needed): sigma plan + update plan at one site pair of a named workload.
usage: python scripts/plan_bench_cpu.py n2_ccpvdz 2000 gauss [site]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chemps2_b200 import api, workloads

name = sys.argv[1] if len(sys.argv) > 1 else "n2_ccpvdz"
D = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
dist = sys.argv[3] if len(sys.argv) > 3 else "gauss"
w = workloads.get(name, D=D)
site = int(sys.argv[4]) if len(sys.argv) > 4 else w.site
ctx = w.context(-1)
w.apply_distribution(ctx, dist)
t0 = time.time()
left = api.OpSet(ctx, site, True) if site > 0 else None
right = api.OpSet(ctx, site + 2, False) if site < w.L - 2 else None
t1 = time.time()
for rep in range(3):
    t2 = time.time()
    heff = api.Heff(ctx, site, left, right)
    t3 = time.time()
    st = heff.stats()
    print(f"heff create {t3 - t2:.3f} s  terms {st['terms']:.0f} tiles {st['tiles']:.0f} stage1 {st['stage1']:.0f} waves {st['waves']:.0f} flops_exec {st['flops_exec']:.3e} wl bytes {st['worklist_bytes']:.3e}", flush=True)
    heff.close()
new = api.OpSet(ctx, site + 1, True)
for rep in range(3):
    t4 = time.time()
    upd = api.Update(ctx, site, True, left, new)
    t5 = time.time()
    print(f"update create {t5 - t4:.3f} s", flush=True)
    del upd
print(f"opsets {t1 - t0:.3f} s")
