"""Sobject::Split (Sobject.cpp:260-622) through its own entry point.  No block-level output of the reference's Split is comparable
directly (singular vectors are fixed only up to signs / rotations inside degenerate spaces), so Split is pinned as the inverse of Join,
which IS pinned against the reference (tests/test_join.py): without truncation Join(Split(S)) == S; with truncation to D states the
discarded weight the routine reports is exactly the lost norm of the state, 1 - |Join(T_left, T_right)|^2 / |S|^2 in the (2S_R + 1)-weighted
norm of the symmetric convention (Sobject.cpp:624-636), the kept dimensions add up to D (global truncation rule, :451-486), and both sweep
directions agree."""
import ctypes as C

import numpy as np
import pytest

import cpu_check
from chemps2_b200 import api
from chemps2_b200._lib import Worklists, check, lib


def _join_cpu(ctx, site, tl, tr):
    j = api.Join(ctx, site)
    wl = Worklists()
    check(lib.b2_join_worklists(j.h, C.byref(wl)))
    o = cpu_check.oracle_lib()
    dp = C.POINTER(C.c_double)
    o.b2o_run_worklists.argtypes = [C.POINTER(Worklists), dp, dp, dp, dp, dp, C.c_int64]
    tl, tr = np.ascontiguousarray(tl, dtype=np.float64), np.ascontiguousarray(tr, dtype=np.float64)
    out, dummy = np.zeros(j.n), np.zeros(1)
    f = lambda a: a.ctypes.data_as(dp)   # noqa: E731
    o.b2o_run_worklists(C.byref(wl), f(tl), f(tr), f(dummy), f(dummy), f(out), out.size)
    return out


def _weights(ctx, site, n):
    labels, offs = ctx.sobject_table(site)
    w = np.zeros(n)
    for k in range(len(labels)):
        w[offs[k]:offs[k + 1]] = labels[k][7] + 1.0          # 2 S_R + 1
    return w


def _total_dim(ctx, boundary, golden):
    L, group = int(golden["problem/hdr"][0]), int(golden["problem/hdr"][1])
    nirr = {0: 1, 5: 4, 7: 8}[group]
    return sum(ctx.dim(boundary, n, ts, ir) for n in range(0, 2 * L + 1) for ts in range(0, L + 2) for ir in range(nirr))


def _case(golden, tag):
    site = int(golden[tag + "/hdr"][0])
    return site, golden[tag + "/joined"]


@pytest.mark.parametrize("tag", ["A", "B"])
@pytest.mark.parametrize("moving_right", [True, False])
def test_split_is_the_inverse_of_join_cpu(golden, tag, moving_right):
    site, S = _case(golden, tag)
    ctx = api.context_from_fixture(golden, tag)
    tl, tr, dw = api.split(ctx, site, S, 10 ** 6, moving_right, True, svd=api.LAPACK_SVD)
    assert abs(dw) < 1e-13
    back = _join_cpu(ctx, site, tl, tr)
    assert np.abs(back - S).max() <= 1e-12 * max(1.0, np.abs(S).max())


@pytest.mark.parametrize("tag", ["A", "B"])
@pytest.mark.parametrize("D", [5, 11])
def test_split_truncation_and_discarded_weight_cpu(golden, tag, D):
    site, S = _case(golden, tag)
    res = {}
    for mr in (True, False):
        ctx = api.context_from_fixture(golden, tag)
        full = _total_dim(ctx, site + 1, golden)
        w = _weights(ctx, site, S.size)
        tl, tr, dw = api.split(ctx, site, S, D, mr, True, svd=api.LAPACK_SVD)
        kept = _total_dim(ctx, site + 1, golden)
        assert kept <= D and kept <= max(full, kept)
        back = _join_cpu(ctx, site, tl, tr)
        norm_s, norm_b = float(np.sum(w * S * S)), float(np.sum(w * back * back))
        assert abs(dw - (1.0 - norm_b / norm_s)) <= 1e-11
        assert abs(dw - float(np.sum(w * (S - back) ** 2)) / norm_s) <= 1e-11     # the truncated state is the projection of S
        assert 0.0 <= dw < 1.0
        res[mr] = (dw, kept)
    assert abs(res[True][0] - res[False][0]) <= 1e-13 and res[True][1] == res[False][1]


def test_split_without_change_keeps_the_dimensions(golden):
    site, S = _case(golden, "A")
    ctx = api.context_from_fixture(golden, "A")
    before = _total_dim(ctx, site + 1, golden)
    size_l, size_r = int(lib.b2_tensor_t_size(ctx.h, site)), int(lib.b2_tensor_t_size(ctx.h, site + 1))
    tl, tr, dw = api.split(ctx, site, S, 10 ** 6, True, False, svd=api.LAPACK_SVD)
    assert _total_dim(ctx, site + 1, golden) == before and tl.size == size_l and tr.size == size_r


def test_split_argument_checks(golden):
    site, S = _case(golden, "A")
    ctx = api.context_from_fixture(golden, "A")
    out, dw = C.c_void_p(), C.c_double()
    s = np.ascontiguousarray(S)
    sp = s.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.b2_sobject_split(ctx.h, site, sp, 0, 1, 1, None, None, C.byref(out), C.byref(dw)) == -1          # D < 1
    assert lib.b2_sobject_split(ctx.h, 99, sp, 10, 1, 1, None, None, C.byref(out), C.byref(dw)) == -1
    assert lib.b2_sobject_split(ctx.h, site, sp, 10, 1, 1, None, None, C.byref(out), C.byref(dw)) == -2          # no device, no caller SVD
    failing = api.SVD_FN(lambda *a: 7)
    assert lib.b2_sobject_split(ctx.h, site, sp, 10, 1, 1, C.cast(failing, C.c_void_p), None, C.byref(out), C.byref(dw)) == -1
    assert b"returned 7" in lib.b2_last_error()
    assert lib.b2_split_size(None, 0) == -1
    lib.b2_split_destroy(None)
