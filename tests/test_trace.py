"""Sobject::Split, TensorT::random and Sobject::addNoise against the reference's OWN outputs (rows S3 / S4 of SURVEY.md 8(a)).

tests/golden/trace_*.npz (tests/golden/make_golden.py -> oracle/ref_driver `trace`) record three noisy half sweeps of the unmodified
reference from its seeded random MPS: per site the two-site object that goes INTO Sobject::Split (after Davidson and addNoise), the
discarded weight, the re-dimensioned boundary, the new site tensors and their Join.  Singular vectors are fixed only up to a gauge, so
tensors are compared through Join(T_left, T_right), which is gauge invariant; everything else is compared directly:
   discarded weight 1e-12 (north_star: 1e-8), new virtual dimensions exactly, Join of the new tensors 1e-10, site energies 1e-9.
"""
import ctypes as C

import numpy as np
import pytest

from chemps2_b200 import api
from chemps2_b200._lib import lib
from test_split import _join_cpu


def _ctx(fx, bk_rows, device=-1):
    L, group, N, twoS, irrep = [int(x) for x in fx["problem/hdr"]]
    ctx = api.Context(device)
    ctx.set_problem(L, group, N, twoS, irrep, fx["problem/orb_irrep"], mx=fx["problem/mx"], econst=float(fx["problem/econst"][0]))
    ctx.bk_init(1)
    ctx.bk_import(bk_rows)
    return ctx


def _dims(ctx, rows):
    return [ctx.dim(int(b), int(n), int(ts), int(ir)) for b, n, ts, ir, _, _ in np.asarray(rows).reshape(-1, 6)]


def test_rand_stream_is_glibc_rand():
    """the private generator behind b2_dmrg_random_mps / the noise == srand(seed); rand(); rand(); ... of the C library the reference calls"""
    libc = C.CDLL("libc.so.6")
    for seed in (0, 1, 1234, 4321, 2 ** 31 + 5):
        libc.srand(C.c_uint(seed & 0xFFFFFFFF))
        ref = [libc.rand() for _ in range(2000)]
        out = (C.c_int * 2000)()
        assert lib.b2_rand_stream(seed, 2000, out) == 0
        assert list(out) == ref


def test_split_vs_reference_every_step_cpu(trace):
    nsteps = int(trace["trace/hdr"][2])
    checked_truncation = 0
    for k in range(nsteps):
        p = f"trace/s{k}"
        index, mr, change, D = [int(x) for x in trace[p + "/hdr"]]
        ctx = _ctx(trace, trace[p + "/bk"])
        tl, tr, dw = api.split(ctx, index, trace[p + "/S"], D, bool(mr), bool(change), svd=api.LAPACK_SVD)
        ref_dw = float(trace[p + "/res"][1])
        assert abs(dw - ref_dw) <= 1e-12, (k, dw, ref_dw)
        after = np.asarray(trace[p + "/bk_after"]).reshape(-1, 6)
        assert _dims(ctx, after) == [int(x) for x in after[:, 4]], k               # the new virtual dimensions, sector by sector
        assert tl.size == trace[p + "/tl"].size and tr.size == trace[p + "/tr"].size
        back = _join_cpu(ctx, index, tl, tr)
        ref = trace[p + "/joined_after"]
        assert np.abs(back - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max()), k
        checked_truncation += ref_dw > 0
    assert checked_truncation >= 4


@pytest.mark.gpu
def test_random_mps_equals_reference_gpu(trace):
    """srand(seed) + TensorT::random + left-normalisation (DMRG::setupBookkeeperAndMPS): the very tensors of the reference"""
    D, seed = int(trace["trace/hdr"][0]), int(trace["trace/hdr"][1])
    ctx = _ctx(trace, trace["trace/bk0"], device=0)
    d = api.DMRG(ctx)
    d.random_mps(seed)
    for s in range(ctx.L):
        ref = trace[f"trace/mps0/{s}"]
        got = d.get_mps(s)
        assert got.size == ref.size
        assert np.abs(got - ref).max() <= 1e-12, s


@pytest.mark.gpu
def test_noisy_sweeps_follow_the_reference_step_by_step_gpu(trace):
    """every solve_site of three noisy half sweeps against the reference's record: site energy 1e-9 at EVERY step; discarded weight 1e-8,
    new dimensions exactly and Join of the new tensors 1e-8 at every step where the Davidson eigenvector comes out with the reference's
    overall sign.  (The sign of an eigenvector is arbitrary — the reference's is whatever dsyev_ returns for the projected matrix — and
    the noise Sobject::addNoise adds element-wise from the shared rand() stream is not invariant under S -> -S: with the opposite sign
    the perturbed state is S - n instead of S + n, an equally valid but different realisation.  Split on the reference's exact input is
    compared at every step in test_split_vs_reference_every_step_cpu.)  After each step the reference's own new tensors are adopted
    (LAPACK's SVD gauge is not reproducible either), so every step starts from the reference's exact state."""
    D, seed = int(trace["trace/hdr"][0]), int(trace["trace/hdr"][1])
    rtol, noise = [float(x) for x in trace["trace/pars"]]
    ctx = _ctx(trace, trace["trace/bk0"], device=0)
    d = api.DMRG(ctx)
    d.random_mps(seed)
    d.presolve()
    nsteps = int(trace["trace/hdr"][2])
    same_sign = truncating = 0
    for k in range(nsteps):
        p = f"trace/s{k}"
        index, mr, change, Dk = [int(x) for x in trace[p + "/hdr"]]
        e, dw, _ = d.solve_site(index, rtol, noise, Dk, bool(mr), bool(change))
        ref_e, ref_dw = [float(x) for x in trace[p + "/res"]]
        assert abs(e - ref_e) <= 1e-9, (k, e, ref_e)
        after = np.asarray(trace[p + "/bk_after"]).reshape(-1, 6)
        ref = trace[p + "/joined_after"]
        sizes_equal = _dims(ctx, after) == [int(x) for x in after[:, 4]]
        got = api.Join(ctx, index).run(d.get_mps(index), d.get_mps(index + 1)) if sizes_equal else None
        if got is not None and float(np.dot(got, ref)) > 0.0:
            same_sign += 1
            truncating += ref_dw > 0
            assert abs(dw - ref_dw) <= 1e-8, (k, dw, ref_dw)
            assert np.abs(got - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max()), k      # identical noise, identical truncation
        else:
            assert dw <= 10.0 * ref_dw + 1e-6, (k, dw, ref_dw)                             # the other noise realisation: same scale
        for b, n, ts, ir, cur, _ in after:                                                # adopt the reference's state
            ctx_dim = ctx.dim(int(b), int(n), int(ts), int(ir))
            if ctx_dim != int(cur):
                from chemps2_b200._lib import check
                check(lib.b2_bk_set_dim(ctx.h, int(b), int(n), int(ts), int(ir), int(cur)))
        d.set_mps(index, trace[p + "/tl"])
        d.set_mps(index + 1, trace[p + "/tr"])
        d.update(index if mr else index + 1, bool(mr))
    assert same_sign >= nsteps // 5, (same_sign, nsteps)
