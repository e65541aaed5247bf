"""GPU side of the per-object Sobject entries (b2_join_run, b2_sobject_split with the batched device SVD); the CPU side is in
tests/test_join.py and tests/test_split.py.  (Sorted last on purpose: these entries were added after the last GPU session of the round.)"""
import glob
import os

import numpy as np
import pytest

from chemps2_b200 import api, fixtures
from test_join import _inputs
from test_split import _case, _total_dim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# own (function-scoped) parametrisation over the fixture files instead of the session-scoped `golden` fixture: pytest groups tests by
# session-scoped parameters, which would interleave these tests with the rest of the suite
PATHS = sorted(p for p in glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))
               if not p.endswith("wigner.npz") and not os.path.basename(p).startswith(("problem_", "trace_")))
IDS = [os.path.basename(p)[:-4] for p in PATHS]


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["A", "B"])
@pytest.mark.parametrize("path", PATHS, ids=IDS)
def test_join_gpu(path, tag):
    """GPU through the C ABI (host buffers in and out)"""
    golden = fixtures.load(path)
    site, tl, tr, ref = _inputs(golden, tag)
    ctx = api.context_from_fixture(golden, tag, device=0)
    out = api.Join(ctx, site).run(tl, tr)
    assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["A", "B"])
@pytest.mark.parametrize("path", PATHS, ids=IDS)
def test_split_device_svd_gpu(path, tag):
    """GPU: the batched device SVD inside Split gives the same truncation (dimensions, discarded weight) as LAPACK, and Join (GPU) of the
    result reproduces S when nothing is truncated"""
    golden = fixtures.load(path)
    site, S = _case(golden, tag)
    ctx = api.context_from_fixture(golden, tag, device=0)
    tl, tr, dw = api.split(ctx, site, S, 10 ** 6, True, True)
    assert abs(dw) < 1e-13
    back = api.Join(ctx, site).run(tl, tr)
    assert np.abs(back - S).max() <= 1e-11 * max(1.0, np.abs(S).max())
    got = {}
    for name, svd in (("device", None), ("lapack", api.LAPACK_SVD)):
        c = api.context_from_fixture(golden, tag, device=0)
        _, _, dwt = api.split(c, site, S, 7, False, True, svd=svd)
        got[name] = (dwt, _total_dim(c, site + 1, golden))
    # the discarded weight depends only on the singular values; the kept total may differ between two SVDs only through ties at round-off
    # level (Sobject.cpp:468-476 keeps values strictly above the (D+1)-th one), so it is bounded, not compared
    assert abs(got["device"][0] - got["lapack"][0]) <= 1e-12
    assert 0 < got["device"][1] <= 7 and 0 < got["lapack"][1] <= 7


@pytest.mark.gpu
def test_pairing_model_known_answer_python():
    """the reference's tests/test12.cpp.in (reduced BCS model, L = 8, folded table written with Problem::setMxElement, not 8-fold symmetric):
    own random start, the two-instruction scheme of the test (D = 100 with noise prefactor 0.5, then D = 1000); known answer
    -25.5134137600604 pinned there to 1e-8"""
    fx = fixtures.load(os.path.join(ROOT, "tests", "golden", "pairing8.npz"))
    L, group, N, twoS, irrep = [int(x) for x in fx["problem/hdr"]]
    ctx = api.Context(0)
    ctx.set_problem(L, group, N, twoS, irrep, fx["problem/orb_irrep"], mx=fx["problem/mx"], econst=float(fx["problem/econst"][0]))
    ctx.bk_init(100)
    d = api.DMRG(ctx)
    d.random_mps(2024)
    e = d.solve([(100, 1e-10, 10, 0.5, 1e-5), (1000, 1e-10, 10, 0.0, 1e-5)])
    assert abs(e - (-25.5134137600604)) < 1e-8, e


@pytest.mark.gpu
def test_pairing_model_known_answer_cpp_caller():
    """the same test written against the C++ mirror exactly like the reference's source (tests/cpp/dmrg_caller.cpp `pairing`): the model is
    set with Problem::setMxElement after the DMRG object exists, PreSolve picks the new table up"""
    import json
    import subprocess
    caller = os.path.join(ROOT, "tests", "cpp", "_bin", "dmrg_caller")
    res = subprocess.run([caller, "pairing"], capture_output=True, text=True, timeout=900)
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("B2JSON ")]
    assert res.returncode == 0 and line, res.stdout[-1500:] + res.stderr[-1500:]
    out = json.loads(line[-1][len("B2JSON "):])
    assert abs(out["energy"] - (-25.5134137600604)) < 1e-8
    assert abs(out["rdm_energy"] - out["energy"]) < 1e-7 and abs(out["trace"] - 8 * 7) < 1e-8
    assert abs(out["pairs"] - 4.0) < 1e-4          # seniority zero: the 8 electrons sit in 4 pairs


@pytest.mark.gpu
def test_momentum_space_hubbard_known_answer():
    """the reference's tests/test9.cpp.in: 3 x 3 Hubbard model with periodic boundaries (U = 5, t = 1, 9 electrons, doublet) in MOMENTUM space -
    a dense folded table with only 4-fold permutation symmetry.  The test demands the site-basis energy; the unmodified reference run here
    (oracle/ref_driver.cpp energies --hubbard2d 3 5.0 -1.0 [--momentum]) gives -6.578839268776 (site) and -6.578839268783 (momentum)."""
    fx = fixtures.load(os.path.join(ROOT, "tests", "golden", "hubbard3x3_momentum.npz"))
    L, group, N, twoS, irrep = [int(x) for x in fx["problem/hdr"]]
    ctx = api.Context(0)
    ctx.set_problem(L, group, N, twoS, irrep, fx["problem/orb_irrep"], mx=fx["problem/mx"], econst=float(fx["problem/econst"][0]))
    ctx.bk_init(500)
    d = api.DMRG(ctx)
    d.random_mps(909)
    e = d.solve([(500, 1e-10, 3, 0.05, 1e-5), (1000, 1e-10, 10, 0.0, 1e-5)])
    assert abs(e - (-6.57883926878)) < 1e-8, e
