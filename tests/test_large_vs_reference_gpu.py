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
