"""Site contribution to the 2-RDM: b2_twodm_* against TwoDM::FillSite of the unmodified reference
(TwoDM.cpp:445-628, doD1..doD24 :642-1592; golden keys UL/twodm_A, UL/twodm_B = the arrays after ONE FillSite call).

CPU: the plan's work lists (effective operators M = T^T [L_g] T) run through the work-list emulator in oracle/, the inner products
<M, stored operator> and diagram 1 are taken with numpy, and b2_twodm_scatter places them like FillSite -> pins every sector loop,
Wigner factor and index combination of the plan without a GPU.  GPU: b2_twodm_fill_site through the C ABI."""
import ctypes as C

import numpy as np
import pytest

import cpu_check
from chemps2_b200 import api, fixtures
from chemps2_b200._lib import Worklists, c_dp, c_ip, c_lp, check, lib, vp

TOL = 1e-12


def _case(golden, device=-1):
    site = int(golden["A/hdr"][0])
    ctx = api.context_from_fixture(golden, "UL", device)
    left = right = None
    if "A/left/hdr" in golden:
        b, mr, ops = fixtures.split_ops(golden, "A/left")
        assert b == site and mr
        left = api.OpSet(ctx, b, True)
        left.upload_all(ops)
    b, mr, ops = fixtures.split_ops(golden, "UL/new")
    assert b == site + 1 and not mr
    right = api.OpSet(ctx, b, False)
    right.upload_all(ops)
    return ctx, site, left, right, golden[f"UL/mps/{site}"]


def _dp(a):
    return a.ctypes.data_as(c_dp)


def test_fill_site_plan_cpu(golden):
    ctx, site, left, right, t = _case(golden)
    p = vp()
    check(lib.b2_twodm_create(ctx.h, site, left.h if left else None, right.h, C.byref(p)))
    try:
        o = cpu_check.oracle_lib()
        o.b2o_run_update_pass.argtypes = [C.POINTER(Worklists), c_dp, c_dp, c_dp, c_dp]
        wl = Worklists()
        check(lib.b2_twodm_worklists(p, C.byref(wl)))
        la = left.host_arena() if left else np.zeros(1)
        tt = np.ascontiguousarray(t, dtype=np.float64)
        M = np.zeros(max(lib.b2_twodm_m_size(p), 1))
        o.b2o_run_update_pass(C.byref(wl), _dp(la), _dp(tt), _dp(np.zeros(1)), _dp(M))
        grams, keep = [], []
        for g in range(lib.b2_twodm_num_groups(p)):
            ls, off, stride, size, nm, npart = C.c_int(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int(), C.c_int()
            check(lib.b2_twodm_group_info(p, g, C.byref(ls), C.byref(off), C.byref(stride), C.byref(size), C.byref(nm), C.byref(npart), None, 0))
            partners = np.zeros(max(npart.value, 1), dtype=np.int32)
            check(lib.b2_twodm_group_info(p, g, None, None, None, None, None, None, partners.ctypes.data_as(c_ip), npart.value))
            side = left if ls.value else right
            G = np.zeros((npart.value, nm.value))                       # [partner][member] = member + members * partner, flattened
            for c in range(npart.value):
                data = side.download(int(partners[c]))
                for m in range(nm.value):
                    a = off.value + m * stride.value
                    G[c, m] = float(M[a:a + size.value] @ data[:size.value]) if size.value else 0.0
            keep.append(np.ascontiguousarray(G.ravel()))
            grams.append(_dp(keep[-1]))
        nk = lib.b2_twodm_d1_scale(p, _dp(np.zeros(1)), 0)
        scale = np.zeros(max(nk, 1))
        lib.b2_twodm_d1_scale(p, _dp(scale), nk)
        # diagram 1 from the T blocks (the scale vector is per block in TensorT storage order)
        d1 = _d1(ctx, int(golden["problem/hdr"][1]), [int(x) for x in golden["problem/orb_irrep"]], site, tt, scale[:nk])
        L = ctx.L
        A, B = np.zeros(L ** 4), np.zeros(L ** 4)
        arr = (c_dp * max(len(grams), 1))(*grams)
        check(lib.b2_twodm_scatter(p, arr, d1, _dp(A), _dp(B)))
    finally:
        lib.b2_twodm_destroy(p)
    refA, refB = golden["UL/twodm_A"], golden["UL/twodm_B"]
    assert np.abs(refA).max() > 0.0                                     # the case exercises something
    assert np.abs(A - refA).max() <= TOL * max(1.0, np.abs(refA).max())
    assert np.abs(B - refB).max() <= TOL * max(1.0, np.abs(refB).max())


def _d1(ctx, group, orb_irrep, site, t, scale):
    """sum_blocks scale_k |T_k|^2 with the TensorT block sizes of site `site` (TensorT.cpp:38-104 enumeration order)"""
    from chemps2_b200._lib import lib as L_
    sizes = []
    # walk the blocks exactly like the layout: NL, 2SL, IL, NR, 2SR (IR follows)
    group_nirr = {0: 1, 1: 2, 2: 2, 3: 2, 4: 4, 5: 4, 6: 4, 7: 8}[group]
    irr_site = orb_irrep[site]
    b = site
    for NL in range(L_.b2_bk_nmin(ctx.h, b), L_.b2_bk_nmax(ctx.h, b) + 1):
        for TwoSL in range(L_.b2_bk_twosmin(ctx.h, b, NL), L_.b2_bk_twosmax(ctx.h, b, NL) + 1, 2):
            for IL in range(group_nirr):
                dl = ctx.dim(b, NL, TwoSL, IL)
                if dl <= 0:
                    continue
                for NR in range(NL, NL + 3):
                    TwoJ = 1 if NR == NL + 1 else 0
                    for TwoSR in range(TwoSL - TwoJ, TwoSL + TwoJ + 1, 2):
                        if TwoSR < 0:
                            continue
                        IR = IL ^ irr_site if NR == NL + 1 else IL
                        dr = ctx.dim(b + 1, NR, TwoSR, IR)
                        if dr > 0:
                            sizes.append(dl * dr)
    assert len(sizes) == len(scale) and sum(sizes) == t.size
    pos, tot = 0, 0.0
    for s, f in zip(sizes, scale):
        tot += f * float(t[pos:pos + s] @ t[pos:pos + s])
        pos += s
    return tot


@pytest.mark.gpu
def test_fill_site_gpu(golden):
    ctx, site, left, right, t = _case(golden, device=0)
    A, B = api.twodm_fill_site(ctx, site, t, left, right)
    refA, refB = golden["UL/twodm_A"], golden["UL/twodm_B"]
    assert np.abs(A - refA).max() <= TOL * max(1.0, np.abs(refA).max())
    assert np.abs(B - refB).max() <= TOL * max(1.0, np.abs(refB).max())
    A2, B2 = api.twodm_fill_site(ctx, site, t, left, right)
    assert np.array_equal(A, A2) and np.array_equal(B, B2)                # deterministic


@pytest.mark.gpu
def test_full_2rdm_vs_reference(golden):
    """b2_dmrg_calc_2rdm on the reference's final MPS (golden corr/mps, corr/bk) against the 2-RDM DMRG::calc2DMandCorrelations of the
    reference computed from the same state (golden twodm/A, twodm/B): north_star tolerance 1e-8 on every element (gauge moves differ,
    so only rounding separates the two); trace = N(N-1) and the energy 0.5 * sum A * gMxElement + Econst equal the reference's."""
    ctx = api.context_from_fixture(golden, "corr", device=0)
    L = ctx.L
    d = api.DMRG(ctx)
    for s in range(L):
        d.set_mps(s, golden[f"corr/mps/{s}"])
    A, B = d.calc_2rdm()
    refA = golden["twodm/A"].reshape((L, L, L, L), order="F")
    refB = golden["twodm/B"].reshape((L, L, L, L), order="F")
    assert np.abs(A - refA).max() < 1e-8
    assert np.abs(B - refB).max() < 1e-8
    N = int(golden["problem/hdr"][2])
    trace = float(np.einsum("ijij->", A))
    assert abs(trace - N * (N - 1)) < 1e-8 and abs(trace - golden["twodm/trace_energy"][0]) < 1e-8
    mx = golden["problem/mx"].reshape((L, L, L, L), order="F")
    energy = float(golden["problem/econst"][0]) + 0.5 * float(np.sum(A * mx))
    assert abs(energy - golden["twodm/trace_energy"][1]) < 1e-8
    assert abs(energy - golden["energies"][-1]) < 1e-6          # and it is the energy of the last sweep step


@pytest.mark.gpu
def test_correlations_vs_reference(golden):
    """b2_dmrg_calc_correlations (after b2_dmrg_calc_2rdm) on the reference's final MPS against the Correlations object of the reference
    for the same state (golden corrfun/*): spin, density, spin-flip and singlet-diradical correlation functions and the two-orbital
    mutual information (Correlations.cpp:69-103, 212-560; G/Y/Z/K/M tensors of DMRGoperators3RDM.cpp:415-479)."""
    ctx = api.context_from_fixture(golden, "corr", device=0)
    L = ctx.L
    d = api.DMRG(ctx)
    for s in range(L):
        d.set_mps(s, golden[f"corr/mps/{s}"])
    A, B = d.calc_2rdm()
    tables = d.calc_correlations(A, B)
    for key in ("Cspin", "Cdens", "Cspinflip", "Cdirad", "MutInfo"):
        ref = golden["corrfun/" + key].reshape((L, L), order="F")
        assert np.abs(tables[key] - ref).max() < 1e-8, key
    assert np.abs(tables["MutInfo"]).max() > 1e-6
