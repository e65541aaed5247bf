"""Helpers shared by the tests: run an exported SigmaPlan through the CPU checker in oracle/ (test infrastructure)."""
import ctypes as C
import os

import numpy as np

from chemps2_b200 import api
from chemps2_b200._lib import FlatPresum, FlatTerm, c_dp
from chemps2_b200.fixtures import split_ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_olib = None


def oracle_lib():
    global _olib
    if _olib is None:
        _olib = C.CDLL(os.path.join(ROOT, "oracle", "libb2oracle.so"))
        _olib.b2o_presum.argtypes = [C.POINTER(FlatPresum), C.c_int64, c_dp, c_dp, c_dp, C.c_int64]
        _olib.b2o_apply.argtypes = [C.POINTER(FlatTerm), C.c_int64, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                    c_dp, c_dp, c_dp, c_dp, c_dp, C.c_int, C.c_int]
    return _olib


def _dp(a):
    return a.ctypes.data_as(c_dp)


def build_case(fx, tag, device=-1, world=1, rank=0, options=None):
    """-> (ctx, left, right, heff) for fixture section `tag`"""
    ctx = api.context_from_fixture(fx, tag, device)
    for name, value in (options or {}).items():
        ctx.set_option(name, value)
    site = int(fx[tag + "/hdr"][0])
    L = ctx.L
    left = right = None
    if site > 0:
        b, mr, ops = split_ops(fx, tag + "/left")
        assert b == site and mr
        left = api.OpSet(ctx, b, True)
        left.upload_all(ops)
    if site < L - 2:
        b, mr, ops = split_ops(fx, tag + "/right")
        assert b == site + 2 and not mr
        right = api.OpSet(ctx, b, False)
        right.upload_all(ops)
    heff = api.Heff(ctx, site, left, right, world, rank)
    return ctx, left, right, heff


def cpu_apply(ctx, left, right, heff, vec, world=1, rank=0):
    """sigma via the plain-C checker (oracle/plan_exec.c) from the exported plan"""
    o = oracle_lib()
    terms, nt, parts, npp, psize = heff.export()
    site_labels, offs = ctx.sobject_table(int(_site_of(heff)))
    nk = len(site_labels)
    la = left.host_arena() if left else np.zeros(1)
    ra = right.host_arena() if right else np.zeros(1)
    presum = np.zeros(max(psize, 1))
    o.b2o_presum(parts, npp, _dp(la), _dp(ra), _dp(presum), psize)
    rows = np.zeros(nk, dtype=np.int32)
    cols = np.zeros(nk, dtype=np.int32)
    for k, lab in enumerate(site_labels):
        rows[k] = ctx.dim(heff._site, int(lab[0]), int(lab[1]), int(lab[2]))
        cols[k] = ctx.dim(heff._site + 2, int(lab[6]), int(lab[7]), int(lab[8]))
    vin = np.ascontiguousarray(vec, dtype=np.float64)
    vout = np.zeros_like(vin)
    boffs = np.ascontiguousarray(offs[:-1])
    o.b2o_apply(terms, nt, nk, boffs.ctypes.data_as(C.POINTER(C.c_int64)), rows.ctypes.data_as(C.POINTER(C.c_int32)),
                cols.ctypes.data_as(C.POINTER(C.c_int32)), _dp(la), _dp(ra), _dp(presum), _dp(vin), _dp(vout), rank, world)
    return vout


def _site_of(heff):
    return heff._site


def emulate_worklists(ctx, left, right, heff, vec):
    """sigma via the work-list emulator (oracle/worklist_emul.cpp): executes the compiled device lists on the CPU"""
    from chemps2_b200._lib import Worklists, check, lib
    o = oracle_lib()
    o.b2o_run_worklists.argtypes = [C.POINTER(Worklists), c_dp, c_dp, c_dp, c_dp, c_dp, C.c_int64]
    wl = Worklists()
    check(lib.b2_heff_worklists(heff.h, C.byref(wl)))
    terms, nt, parts, npp, psize = heff.export()
    la = left.host_arena() if left else np.zeros(1)
    ra = right.host_arena() if right else np.zeros(1)
    presum = np.zeros(max(psize, 1))
    o.b2o_presum(parts, npp, _dp(la), _dp(ra), _dp(presum), psize)
    vin = np.ascontiguousarray(vec, dtype=np.float64)
    vout = np.zeros_like(vin)
    o.b2o_run_worklists(C.byref(wl), _dp(la), _dp(ra), _dp(presum), _dp(vin), _dp(vout), vin.size)
    return vout


def emulate_diag(ctx, left, right, heff):
    """diagonal of H_eff via the CPU emulation of the diagonal lists (oracle/worklist_emul.cpp b2o_run_diag)"""
    from chemps2_b200._lib import check, lib
    o = oracle_lib()
    vp = C.c_void_p
    o.b2o_run_diag.argtypes = [vp, vp, C.c_int64, c_dp, c_dp, c_dp, c_dp, C.c_int64]
    items, tiles, ni, nt = vp(), vp(), C.c_int64(), C.c_int64()
    check(lib.b2_heff_diag_lists(heff.h, C.byref(items), C.byref(ni), C.byref(tiles), C.byref(nt)))
    terms, nterms, parts, npp, psize = heff.export()
    la = left.host_arena() if left else np.zeros(1)
    ra = right.host_arena() if right else np.zeros(1)
    presum = np.zeros(max(psize, 1))
    o.b2o_presum(parts, npp, _dp(la), _dp(ra), _dp(presum), psize)
    out = np.zeros(max(heff.n, 1))
    o.b2o_run_diag(items, tiles, nt.value, _dp(la), _dp(ra), _dp(presum), _dp(out), heff.n)
    return out[:heff.n], ni.value


def build_update_case(fx, which, device=-1, options=None, world=1, rank=0):
    """which = 'UR' (moving right after the solve at siteB) or 'UL' (moving left after the solve at siteA).
    -> (ctx, old_set, new_set, update, t_storage, expected [(kind, i, j, data)])"""
    tag = "B" if which == "UR" else "A"
    site = int(fx[tag + "/hdr"][0])
    mr = which == "UR"
    ctx = api.context_from_fixture(fx, which, device)
    for name, value in (options or {}).items():
        ctx.set_option(name, value)
    L = ctx.L
    index = site if mr else site + 1
    old = None
    key = tag + ("/left" if mr else "/right")
    if key + "/hdr" in fx:
        b, omr, ops = split_ops(fx, key)
        assert omr == mr and b == (index if mr else index + 1)
        old = api.OpSet(ctx, b, mr)
        old.upload_all(ops)
    nb, nmr, expected = split_ops(fx, which + "/new")
    assert nmr == mr and nb == (index + 1 if mr else index)
    new = api.OpSet(ctx, nb, mr)
    upd = api.Update(ctx, index, mr, old, new, world, rank)
    return ctx, old, new, upd, fx[f"{which}/mps/{index}"], expected


def emulate_update(old, new, upd, t_storage, passes=(0, 1), arena=None):
    """runs the given passes of the compiled update work lists on the CPU (oracle/worklist_emul.cpp) -> new arena (numpy);
    `arena`: state to continue from (the all-reduced pass-0 result of a sharded update)"""
    from chemps2_b200._lib import Worklists, check, lib
    o = oracle_lib()
    o.b2o_run_update_pass.argtypes = [C.POINTER(Worklists), c_dp, c_dp, c_dp, c_dp]
    npp = lib.b2_update_num_presum_parts(upd.h)
    parts = (FlatPresum * max(npp, 1))()
    check(lib.b2_update_export_presums(upd.h, parts))
    psize = lib.b2_update_presum_size(upd.h)
    oa = old.host_arena() if old else np.zeros(1)
    presum = np.zeros(max(psize, 1))
    o.b2o_presum(parts, npp, _dp(oa), _dp(oa), _dp(presum), psize)
    arena = np.zeros(max(lib.b2_opset_arena_size(new.h), 1)) if arena is None else np.ascontiguousarray(arena, dtype=np.float64).copy()
    t = np.ascontiguousarray(t_storage, dtype=np.float64)
    for p in passes:
        wl = Worklists()
        check(lib.b2_update_worklists(upd.h, p, C.byref(wl)))
        o.b2o_run_update_pass(C.byref(wl), _dp(oa), _dp(t), _dp(presum), _dp(arena))
        if p == 1:   # whole-operator mixing after the transposed copies of pass 1 (device: k_mix_flat)
            nf = lib.b2_update_num_mix_flat(upd.h)
            flat = (FlatPresum * max(nf, 1))()
            check(lib.b2_update_export_mix_flat(upd.h, flat))
            add = np.zeros_like(arena)
            for i in range(nf):
                f = flat[i]
                src = presum if f.space == 3 else arena
                add[f.dst_off:f.dst_off + f.size] += f.coef * src[f.src_off:f.src_off + f.size]
            arena += add
    return arena


def op_slices(new):
    """[(kind, i, j, offset, size)] of an OpSet arena"""
    from chemps2_b200._lib import lib
    out = []
    # offsets follow the 16-double alignment of OpSet::add
    off = 0
    for idx in range(len(new)):
        k, i, j, size = new.info(idx)
        out.append((k, i, j, off, size))
        off += (size + 15) // 16 * 16
    return out


def worklist_digest(wl):
    """sha256 over every compiled device list of a Worklists export (items, tiles per class, reduce jobs, waves, sizes)"""
    import hashlib
    sizes = dict(item=40, tile=48, reduce=56, wave=4 * (4 * 4 + 2))   # sizeof GemmItem / Tile / ReduceJob / Wave (b2_device.h, b2_compile.h)
    h = hashlib.sha256()

    def add(ptr, n, size):
        if n and ptr:
            h.update(C.string_at(ptr, int(n) * size))
    add(wl.items1, wl.n_items1, sizes["item"])
    add(wl.items2, wl.n_items2, sizes["item"])
    for c in range(4):
        add(wl.tiles1[c], wl.n_tiles1[c], sizes["tile"])
        add(wl.tiles2[c], wl.n_tiles2[c], sizes["tile"])
    add(wl.reduces, wl.n_reduces, sizes["reduce"])
    add(wl.waves, wl.n_waves, sizes["wave"])
    h.update(repr((wl.n_items1, wl.n_items2, list(wl.n_tiles1), list(wl.n_tiles2), wl.n_reduces, wl.n_waves, wl.work_size, wl.part_size)).encode())
    return h.hexdigest()[:24]
