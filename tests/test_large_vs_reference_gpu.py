"""The GPU sigma build and Heff diagonal against the UNMODIFIED reference at sizes where the large tile classes (64 / 32), the multi-chunk K
pipeline, split-K and several waves are all in play — the golden fixtures stop at D = 64.  The reference runs on the host cores of the
same box (oracle/_ref/ref_driver `synth`: Heff::makeHeff + fillHeffDiag on hash-filled operators, the same fill b2_opset_fill_hash
produces); it is test infrastructure and travels prebuilt."""
import numpy as np
import pytest

from chemps2_b200 import api, workloads
from oracle import refrun

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not built")]

CASES = [
    ("synth40", 600, "gauss", None),     # config 5 shape, converged-like sector model: blocks of 30-60 rows, 64/32 tile classes, split-K
    ("synth40", 900, "flat", None),      # the reference's flat first-sweep distribution: thousands of tiny blocks (8/16 tile classes)
    ("n2", 500, "flat", None),           # config 2 shape: D2h, a few large sectors + a long tail of tiny ones
    ("n2", 700, "gauss", None),
    ("synth60", 500, "gauss", None),     # config 4 shape (60 orbitals: the O(L^2) pair sums dominate)
    ("tetracene", 1500, "gauss", 3),      # config 3 shape, away from the middle of the chain (left operators fewer than right ones)
]


@pytest.mark.parametrize("name,D,dist,site", CASES, ids=[f"{c[0]}-D{c[1]}-{c[2]}" for c in CASES])
def test_sigma_and_diag_vs_reference_large(name, D, dist, site):
    w = workloads.get(name, D=D, site=site)
    ctx = w.context(0)
    dims = w.apply_distribution(ctx, dist)
    left, right = api.OpSet(ctx, w.site, True), api.OpSet(ctx, w.site + 2, False)
    left.fill_hash(7, 1.0)
    right.fill_hash(7, 1.0)
    heff = api.Heff(ctx, w.site, left, right)
    ref = refrun.run_reference_synth(w, 7, reps=1, dims=dims)
    assert ref["veclength"] == heff.n
    out = heff.apply(api.hash_fill(heff.n, 7))
    scale = np.abs(ref["vec_out"]).max()
    assert np.abs(out - ref["vec_out"]).max() <= 1e-12 * scale
    dscale = np.abs(ref["diag"]).max()
    assert np.abs(heff.diag() - ref["diag"]).max() <= 1e-12 * dscale
    st = heff.stats()
    assert st["flops_ref"] > 5e8


UPDATE_CASES = [
    ("synth40", 400, "gauss", 19, True),     # config 5 shape, moving right over the middle site
    ("synth40", 400, "gauss", 20, False),    # ... and moving left
    ("n2", 500, "flat", 13, True),           # D2h: many tiny blocks
    ("tetracene", 800, "gauss", 5, False),
    ("synth60", 300, "gauss", 0, True),      # the chain edge: TensorL::create / makenew paths only (no old operators)
]


@pytest.mark.parametrize("name,D,dist,site,mr", UPDATE_CASES, ids=[f"{c[0]}-D{c[1]}-{c[2]}-site{c[3]}-{'right' if c[4] else 'left'}" for c in UPDATE_CASES])
def test_operator_update_vs_reference_large(name, D, dist, site, mr):
    """b2_update_run == DMRG::updateMovingRight / updateMovingLeft of the unmodified reference (every operator kind of the new boundary) at
    sizes where the large tile classes, split-K and the second (A/B/C/D mixing) pass with thousands of terms are in play.  Per operator
    the sum, the sum of squares and the dot product with a hash vector are compared (1e-11 relative to the operator's norm)."""
    w = workloads.get(name, D=D, site=site)
    ctx = w.context(0)
    dims = w.apply_distribution(ctx, dist)
    L = w.L
    b_old, b_new = (site, site + 1) if mr else (site + 1, site)
    have_old = (site > 0) if mr else (site < L - 1)
    old = api.OpSet(ctx, b_old, mr) if have_old else None
    if old is not None:
        old.fill_hash(9, 1.0)
    new = api.OpSet(ctx, b_new, mr)
    upd = api.Update(ctx, site, mr, old, new)
    from chemps2_b200._lib import lib
    nt = lib.b2_tensor_t_size(ctx.h, site)
    upd.run(api.hash_fill(nt, 9, key=refrun.op_key(4, 0, -1, -1), amp=0.1))
    ref = refrun.run_reference_update(w, 9, moving_right=mr, site=site, dims=dims)
    assert len(ref["ops"]) == len(new)
    checked = 0
    for kind, i, j, size, rsum, rsq, rdot in ref["ops"]:
        idx = new.find(kind, i, j)
        assert idx >= 0 and new.info(idx)[3] == size, (api.KIND_NAMES[kind], i, j)
        if size == 0:
            continue
        got = new.download(idx)
        h = api.hash_fill(size, 9 + 17, key=refrun.op_key(5, kind, i, j))
        norm = max(np.sqrt(rsq), 1e-300)
        tol = 1e-11 * norm * np.sqrt(size)
        assert abs(got.sum() - rsum) <= tol and abs(float(np.dot(got, got)) - rsq) <= 1e-11 * max(rsq, 1e-300) * 10 and abs(float(np.dot(got, h)) - rdot) <= tol, \
            (api.KIND_NAMES[kind], i, j, got.sum(), rsum, float(np.dot(got, h)), rdot)
        checked += 1
    assert checked > 10
