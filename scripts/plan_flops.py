"""FLOPs per (wave, stage, tile class) of a compiled sigma plan — to relate ncu launch durations to executed FLOPs."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chemps2_b200 import api, workloads
from chemps2_b200._lib import Worklists, check, lib

item_dt = np.dtype([("xoff", "<i8"), ("yoff", "<i8"), ("alpha", "<f8"), ("ldx", "<i4"), ("ldy", "<i4"), ("k", "<i4"), ("xs", "u1"), ("ys", "u1"), ("flags", "u1"), ("pad", "u1")])
tile_dt = np.dtype([("coff", "<i8"), ("ldc", "<i4"), ("m0", "<i4"), ("n0", "<i4"), ("mrem", "<i4"), ("nrem", "<i4"), ("cm0", "<i4"), ("cn0", "<i4"),
                    ("ib", "<i4"), ("ie", "<i4"), ("cspace", "u1"), ("acc", "u1"), ("pad", "u1", 2)])
wave_dt = np.dtype([("t1b", "<i4", 4), ("t1e", "<i4", 4), ("t2b", "<i4", 4), ("t2e", "<i4", 4), ("rb", "<i4"), ("re", "<i4")])
assert item_dt.itemsize == 40 and tile_dt.itemsize == 48

def arr(ptr, n, dt):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dt)
    return np.frombuffer((C.c_char * (n * dt.itemsize)).from_address(ptr), dtype=dt)

name, D, dist = sys.argv[1], int(sys.argv[2]), sys.argv[3]
w = workloads.get(name, D=D)
ctx = w.context(-1)
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    ctx.set_option(k, float(v))
w.apply_distribution(ctx, dist)
left, right = api.OpSet(ctx, w.site, True), api.OpSet(ctx, w.site + 2, False)
h = api.Heff(ctx, w.site, left, right)
wl = Worklists()
check(lib.b2_heff_worklists(h.h, C.byref(wl)))
it1, it2 = arr(wl.items1, wl.n_items1, item_dt), arr(wl.items2, wl.n_items2, item_dt)
waves = arr(wl.waves, wl.n_waves, wave_dt)
ksum1 = np.concatenate([[0], np.cumsum(it1["k"].astype(np.int64))])
ksum2 = np.concatenate([[0], np.cumsum(it2["k"].astype(np.int64))])
tot = np.zeros((2, 4)); ncta = np.zeros((2, 4)); pad = np.zeros((2, 4))
per_wave = []
for c in range(4):
    for st, (tp, n, ks) in enumerate(((wl.tiles1[c], wl.n_tiles1[c], ksum1), (wl.tiles2[c], wl.n_tiles2[c], ksum2))):
        t = arr(tp, n, tile_dt)
        if len(t) == 0:
            continue
        K = ks[t["ie"]] - ks[t["ib"]]
        fl = 2.0 * t["mrem"] * t["nrem"] * K
        edge = [64, 32, 16, 8][c]
        flp = 2.0 * (np.ceil(t["mrem"] / 8) * 8) * (np.ceil(t["nrem"] / 8) * 8) * K
        tot[st, c] = fl.sum(); ncta[st, c] = len(t); pad[st, c] = flp.sum()
        if c == 0:
            cs = np.concatenate([[0], np.cumsum(fl)])
            b, e = (waves["t1b"][:, 0], waves["t1e"][:, 0]) if st == 0 else (waves["t2b"][:, 0], waves["t2e"][:, 0])
            per_wave.append((cs[e] - cs[b], e - b))
st = h.stats()
print("flops_ref %.3e flops_exec %.3e waves %d" % (st["flops_ref"], st["flops_exec"], st["waves"]))
for s in range(2):
    for c in range(4):
        print(f"stage {s+1} class {[64,32,16,8][c]:2d}: ctas {int(ncta[s,c]):9d} gflop {tot[s,c]/1e9:12.2f} share {tot[s,c]/tot.sum():.3f} mma-padded/exact {pad[s,c]/max(tot[s,c],1):.3f}")
for s in range(2):
    fl, n = per_wave[s]
    sel = n > 0
    print(f"class-64 stage {s+1} per wave: median gflop {np.median(fl[sel])/1e9:.2f} median ctas {np.median(n[sel]):.0f} min ctas {n[sel].min()} max ctas {n[sel].max()}")
