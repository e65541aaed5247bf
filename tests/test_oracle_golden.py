"""CPU tests: the host logic (sector tables, Wigner symbols, the SigmaPlan) and the plain-C checker in oracle/ are pinned
against golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py -> Heff::makeHeff)."""
import os

import numpy as np
import pytest

import cpu_check
from chemps2_b200 import api, fixtures
from chemps2_b200._lib import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_wigner_table():
    """b2::wigner6j/9j (host, plan-build time) vs CheMPS2::Wigner (Wigner.cpp:294-368)"""
    import ctypes as C
    fx = fixtures.load(os.path.join(ROOT, "tests", "golden", "wigner.npz"))
    lib.b2_wigner6j.restype = C.c_double
    lib.b2_wigner9j.restype = C.c_double
    a6 = fx["w6j/args"].reshape(-1, 6)
    got = np.array([lib.b2_wigner6j(*[int(x) for x in r]) for r in a6])
    assert np.abs(got - fx["w6j/vals"]).max() < 1e-13
    assert np.count_nonzero(fx["w6j/vals"]) > 50
    a9 = fx["w9j/args"].reshape(-1, 9)
    got = np.array([lib.b2_wigner9j(*[int(x) for x in r]) for r in a9])
    assert np.abs(got - fx["w9j/vals"]).max() < 1e-13
    assert np.count_nonzero(fx["w9j/vals"]) > 5


def test_bookkeeper_fci_dims(golden):
    """Bookkeeper::init reproduces SyBookkeeper's sector ranges and FCI dimensions (SyBookkeeper.cpp:76-233)"""
    L, group, N, twoS, irrep = [int(x) for x in golden["problem/hdr"]]
    ctx = api.Context(-1)
    ctx.set_problem(L, group, N, twoS, irrep, golden["problem/orb_irrep"], mx=golden["problem/mx"])
    ctx.bk_init(7)
    rows = golden["A/bk"].reshape(-1, 6)
    for b, n, ts, ir, _, fci in rows:
        assert ctx.fcidim(int(b), int(n), int(ts), int(ir)) == fci
    # sector ranges: every row the reference enumerates exists, and nothing outside does
    per_b = {}
    for b, n, ts, ir, _, _ in rows:
        per_b.setdefault(int(b), set()).add((int(n), int(ts)))
    for b, s in per_b.items():
        assert lib.b2_bk_nmin(ctx.h, b) == min(n for n, _ in s) and lib.b2_bk_nmax(ctx.h, b) == max(n for n, _ in s)
        for n in set(n for n, _ in s):
            assert lib.b2_bk_twosmin(ctx.h, b, n) == min(t for m, t in s if m == n)
            assert lib.b2_bk_twosmax(ctx.h, b, n) == max(t for m, t in s if m == n)


def test_mx_elem_from_integrals(golden):
    """Problem::build == Problem::construct_mxelem (Problem.cpp:363-384): one-body part folded into the two-body table"""
    if "problem/direct_mx" in golden and int(golden["problem/direct_mx"][0]):
        pytest.skip("the table of this fixture was written with Problem::setMxElement (tests/test12.cpp.in), not folded from (T, V)")
    L, group, N, twoS, irrep = [int(x) for x in golden["problem/hdr"]]
    ctx = api.Context(-1)
    ctx.set_problem(L, group, N, twoS, irrep, golden["problem/orb_irrep"], tmat=golden["problem/tmat"], vmat=golden["problem/vmat"])
    got = ctx.mx_elem()
    assert np.abs(got - golden["problem/mx"]).max() < 1e-14


@pytest.mark.parametrize("tag", ["A", "B"])
def test_layout_sizes(golden, tag):
    """packed sizes of TensorT / Sobject / every operator equal the reference's kappa2index totals"""
    ctx = api.context_from_fixture(golden, tag)
    for s in range(ctx.L):
        assert lib.b2_tensor_t_size(ctx.h, s) == golden[f"{tag}/mps/{s}"].size
    site, n, nk = [int(x) for x in golden[tag + "/hdr"]]
    assert lib.b2_sobject_size(ctx.h, site) == n
    assert lib.b2_sobject_nkappa(ctx.h, site) == nk
    for side in ("left", "right"):
        if tag + "/" + side + "/hdr" not in golden:
            continue
        b, mr, ops = fixtures.split_ops(golden, tag + "/" + side)
        st = api.OpSet(ctx, b, mr)
        assert len(st) == len(ops)
        for kind, i, j, data in ops:
            idx = st.find(kind, i, j)
            assert idx >= 0 and st.info(idx)[3] == data.size


@pytest.mark.parametrize("tag", ["A", "B"])
def test_sigma_plan_vs_reference(golden, tag):
    """SigmaPlan (all diagram families) executed by the plain-C checker == Heff::makeHeff output of the reference"""
    ctx, left, right, heff = cpu_check.build_case(golden, tag)
    for a, b in (("vec_in", "vec_out"), ("rnd_in", "rnd_out")):
        out = cpu_check.cpu_apply(ctx, left, right, heff, golden[f"{tag}/{a}"])
        ref = golden[f"{tag}/{b}"]
        assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    st = heff.stats()
    assert st["terms"] > 0 and st["flops_ref"] > 0


@pytest.mark.parametrize("world", [2, 8])
def test_sigma_plan_owner_sharding(golden, world):
    """partial sigma vectors of the owner shards (MPIchemps2.h:158-231 maps) sum to the full sigma"""
    ctx, left, right, heff = cpu_check.build_case(golden, "A", world=world, rank=0)
    vin = golden["A/rnd_in"]
    tot = np.zeros_like(vin)
    for r in range(world):
        tot += cpu_check.cpu_apply(ctx, left, right, heff, vin, world=world, rank=r)
    ref = golden["A/rnd_out"]
    assert np.abs(tot - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("options", [{}, {"work_budget": 2048, "chunk_k": 64}, {"work_budget": 50000, "chunk_k": 16},
                                     {"parallel_min_terms": 16}, {"parallel_min_terms": 16, "work_budget": 16384, "chunk_k": 32}],
                         ids=["default", "tiny-waves", "tiny-chunks", "parallel-plan", "parallel-plan-tiny-waves"])
@pytest.mark.parametrize("tag", ["A", "B"])
def test_compiled_worklists_vs_reference(golden, tag, options):
    """the device work lists (waves, split-K chunks, reduces, shared intermediates) executed by the CPU emulator
    reproduce Heff::makeHeff — checks the scheduler of b2_heff.cpp without a GPU"""
    ctx, left, right, heff = cpu_check.build_case(golden, tag, options=options)
    vin, ref = golden[f"{tag}/rnd_in"], golden[f"{tag}/rnd_out"]
    out = cpu_check.emulate_worklists(ctx, left, right, heff, vin)
    assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    st = heff.stats()
    if options.get("work_budget", 1e9) < 10000 and st["terms"] > 500:
        assert st["waves"] > 1


@pytest.mark.parametrize("tag", ["A", "B"])
def test_diagonal_lists_vs_reference(golden, tag):
    """the diagonal lists of the plan (what k_diag executes) evaluated on the CPU reproduce Heff::fillHeffDiag (Heff.cpp:250-315)"""
    ctx, left, right, heff = cpu_check.build_case(golden, tag)
    out, nitems = cpu_check.emulate_diag(ctx, left, right, heff)
    ref = golden[tag + "/diag"]
    assert nitems > 0
    assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


def test_no_device_is_loud(golden):
    """planning-only context: compute entry points fail with B2_ERR_NO_DEVICE instead of falling back to the CPU"""
    ctx, left, right, heff = cpu_check.build_case(golden, "A")
    with pytest.raises(api.B2Error):
        heff.apply(golden["A/vec_in"])


@pytest.mark.parametrize("n", [1, 2, 3, 7, 32])
def test_small_symmetric_eig(n):
    """host Rayleigh-Ritz step of the device Davidson (stand-in for dsyev_, Davidson.cpp:276)"""
    from chemps2_b200._lib import c_dp, check
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n))
    a = a + a.T
    if n == 7:
        a[3, :] = a[2, :]; a[:, 3] = a[:, 2]; a[3, 3] = a[2, 2]   # degenerate pair
    ev, vec = np.zeros(n), np.zeros(n * n)
    check(lib.b2_small_symmetric_eig(n, a.ravel(order="F").ctypes.data_as(c_dp), ev.ctypes.data_as(c_dp), vec.ctypes.data_as(c_dp)))
    v = vec.reshape(n, n, order="F")
    assert np.abs(ev - np.linalg.eigvalsh(a)).max() < 1e-12 * max(1.0, np.abs(a).max())
    assert np.abs(v.T @ v - np.eye(n)).max() < 1e-12
    assert np.abs(a @ v - v * ev).max() < 1e-11 * max(1.0, np.abs(a).max())
