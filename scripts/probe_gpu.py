"""Measures the FP64 roofline denominators on the box (register-resident DMMA / DFMA loops) and times a few sigma builds."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from chemps2_b200 import api, workloads
from chemps2_b200._lib import check, lib

ctx = api.Context(0)
res = {}
for mode, name in ((1, "dmma_tflops"), (0, "dfma_tflops")):
    v = C.c_double()
    check(lib.b2_probe_fp64(ctx.h, mode, C.byref(v)))
    res[name] = v.value
print(json.dumps(res))
import argparse
ap = argparse.ArgumentParser()
ap.add_argument("--cases", default="tiny:150,n2:120,n2:500,synth40:400,tetracene:1000")
ap.add_argument("--work-budget", type=float, default=0)
ap.add_argument("--chunk-k", type=float, default=0)
args = ap.parse_args()
for case in args.cases.split(","):
    name, D = case.split(":")[0], int(case.split(":")[1])
    w = workloads.get(name, D=D)
    c = w.context(0)
    if args.work_budget:
        c.set_option("work_budget", args.work_budget)
    if args.chunk_k:
        c.set_option("chunk_k", args.chunk_k)
    sets = [api.OpSet(c, w.site, True), api.OpSet(c, w.site + 2, False)]
    for s in sets:
        s.fill_hash(5, 1.0)
    t0 = time.time()
    h = api.Heff(c, w.site, *sets)
    tplan = time.time() - t0
    vin = api.hash_fill(h.n, 5)
    for _ in range(3):
        out = h.apply(vin)
    st = h.stats()
    ks = h.kernel_seconds()
    print(json.dumps(dict(workload=w.describe(), veclength=int(h.n), plan_s=round(tplan, 3), kernel_ms=ks * 1e3, terms=st["terms"],
                          gflops_ref=st["flops_ref"] / 1e9, gflops_exec=st["flops_exec"] / 1e9, waves=st["waves"], launches=st["launches"], ctas=st["tiles"],
                          work_mb=st["work_doubles"] * 8e-6, part_mb=st["part_doubles"] * 8e-6, lists_mb=st["worklist_bytes"] / 1e6, tflops_achieved=st["flops_ref"] / ks / 1e12, norm=float(np.linalg.norm(out)))))
