#!/usr/bin/env python
"""Where do the issued DMMAs of a sigma plan go?  Reads the compiled work lists (b2_heff_worklists) of a workload on a planning-only
context and models k_tiles exactly: per tile the sub-tiles every warp computes (predicate-free path when >= 3/4 of a warp's sub-tiles
are inside), K padded to the 16-column chunk.  Prints useful vs issued FLOPs per stage and tile class."""
import ctypes as C
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chemps2_b200 import api, workloads  # noqa: E402
from chemps2_b200._lib import Worklists, check, lib  # noqa: E402

ITEM = np.dtype([("xoff", "<i8"), ("yoff", "<i8"), ("alpha", "<f8"), ("ldx", "<i4"), ("ldy", "<i4"), ("k", "<i4"), ("xs", "u1"), ("ys", "u1"), ("flags", "u1"), ("pad", "u1")])
TILE = np.dtype([("coff", "<i8"), ("ldc", "<i4"), ("m0", "<i4"), ("n0", "<i4"), ("mrem", "<i4"), ("nrem", "<i4"), ("cm0", "<i4"), ("cn0", "<i4"),
                 ("item_begin", "<i4"), ("item_end", "<i4"), ("cspace", "u1"), ("accumulate", "u1"), ("pad", "u1", 2)])
CLASSES = [(64, 64, 2, 2), (32, 32, 2, 2), (16, 16, 1, 1), (8, 8, 1, 1)]


def view(ptr, n, dt):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dt)
    buf = (C.c_char * (n * dt.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dt)


def model(tiles, items, cls, kc=16):
    TM, TN, WM, WN = CLASSES[cls]
    WTM, WTN = TM // WM, TN // WN
    MI, NI = WTM // 8, WTN // 8
    if len(tiles) == 0:
        return 0.0, 0.0, 0
    gemm = (items["flags"] & 4) == 0
    k = np.where(gemm, items["k"], 0).astype(np.int64)
    kp = np.where(k >= 0, (k + kc - 1) // kc * kc, 0)
    # tail chunk with fewer than 4 valid columns executes only those steps... (steps S < kvalid): model exactly
    tail = k % kc
    kp = np.where((tail > 0) & (tail < 4), k - tail + tail * 4, kp)   # steps executed x 4 columns each
    ck = np.concatenate([[0], np.cumsum(k)])
    ckp = np.concatenate([[0], np.cumsum(kp)])
    ksum = ck[tiles["item_end"]] - ck[tiles["item_begin"]]
    kpsum = ckp[tiles["item_end"]] - ckp[tiles["item_begin"]]
    mrem, nrem = tiles["mrem"].astype(np.int64), tiles["nrem"].astype(np.int64)
    useful = 2.0 * (mrem * nrem * ksum).sum()
    sub = np.zeros(len(tiles), dtype=np.int64)
    for wm in range(WM):
        for wn in range(WN):
            mi = np.clip((mrem - wm * WTM + 7) >> 3, 0, MI)
            ni = np.clip((nrem - wn * WTN + 7) >> 3, 0, NI)
            full = mi * ni * 4 >= MI * NI * 3
            sub += np.where(full, MI * NI, mi * ni)
    issued = 2.0 * (sub * 64 * kpsum).sum()
    return useful, issued, len(tiles)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "synth40"
    D = int(sys.argv[2]) if len(sys.argv) > 2 else None
    dist = sys.argv[3] if len(sys.argv) > 3 else "gauss"
    w = workloads.get(name, D=D)
    ctx = w.context(-1)
    for a in sys.argv[4:]:
        k, v = a.split("=")
        ctx.set_option(k, float(v))
    w.apply_distribution(ctx, dist)
    left, right = api.OpSet(ctx, w.site, True), api.OpSet(ctx, w.site + 2, False)
    heff = api.Heff(ctx, w.site, left, right)
    st = heff.stats()
    wl = Worklists()
    check(lib.b2_heff_worklists(heff.h, C.byref(wl)))
    i1, i2 = view(wl.items1, wl.n_items1, ITEM), view(wl.items2, wl.n_items2, ITEM)
    tot_u = tot_i = 0.0
    print(f"{w.describe()} dist={dist}: flops_ref {st['flops_ref']:.4e} flops_exec {st['flops_exec']:.4e} waves {st['waves']:.0f}")
    for stage, items, tl, nt in ((1, i1, wl.tiles1, wl.n_tiles1), (2, i2, wl.tiles2, wl.n_tiles2)):
        for c in range(4):
            tiles = view(tl[c], nt[c], TILE)
            u, i, n = model(tiles, items, c)
            tot_u += u
            tot_i += i
            if n:
                print(f"  stage {stage} class {CLASSES[c][0]:2d}: {n:9d} CTAs  useful {u:.4e}  issued {i:.4e}  useful/issued {u / max(i, 1):.3f}  mean mrem {tiles['mrem'].mean():.1f} nrem {tiles['nrem'].mean():.1f}")
    print(f"  total useful {tot_u:.4e} issued {tot_i:.4e} ratio {tot_u / tot_i:.3f}")


if __name__ == "__main__":
    main()
