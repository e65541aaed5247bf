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
