"""MPS checkpoints in the reference's schema (DMRG::saveMPS / loadDIM / loadMPS, DMRGmpsio.cpp:30-131: /Convergence/Converged_yn,
/VirtDim_<b>_<N>_<2S>_<I>/Value, /MPS_<site>/Values): the unmodified reference (oracle/_ref/ref_driver, HDF5 calls bridged by
env_shims/hdf5.h) and the GPU library resume EACH OTHER's states, and the sweep that follows gives the same site energies (1e-9 Eh)."""
import os
import re
import subprocess

import numpy as np
import pytest

from chemps2_b200 import api, fixtures
from oracle import refrun

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not built")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = 30


def _problem_file(fx, path):
    """binary problem file of `ref_driver --problem` (chemps2_b200/workloads.py write_problem_file) from a golden fixture"""
    with open(path, "wb") as f:
        np.asarray(fx["problem/hdr"], dtype="<i4").tofile(f)
        np.asarray(fx["problem/orb_irrep"], dtype="<i4").tofile(f)
        np.asarray(fx["problem/econst"], dtype="<f8").tofile(f)
        np.asarray(fx["problem/tmat"], dtype="<f8").tofile(f)
        np.asarray(fx["problem/vmat"], dtype="<f8").tofile(f)


def _reference_sweeps(pfile, workdir, nsweeps, seed=99):
    """runs (or resumes, when workdir holds CheMPS2_MPS0.h5) the reference for `nsweeps` sweeps -> per-site energies in the order printed"""
    env = dict(os.environ, OMP_NUM_THREADS="4", OPENBLAS_NUM_THREADS="1")
    res = subprocess.run([refrun.REF_DRIVER, "energies", "--problem", pfile, "--schedule", f"{D}:1e-14:{nsweeps}:0.0:1e-8", "--seed", str(seed),
                          "--chkpt-dir", workdir], capture_output=True, text=True, env=env, check=True)
    loaded = "Loaded MPS" in res.stdout
    return [float(x) for x in re.findall(r"Energy at sites \(\d+, \d+\) is (-?[\d.]+)", res.stdout)], loaded


def _gpu_driver(fx):
    L, group, N, twoS, irrep = [int(x) for x in fx["problem/hdr"]]
    ctx = api.Context(0)
    ctx.set_problem(L, group, N, twoS, irrep, fx["problem/orb_irrep"], mx=fx["problem/mx"], econst=float(fx["problem/econst"][0]))
    ctx.bk_init(D)
    return ctx, api.DMRG(ctx)


def _gpu_sweep_site_energies(ctx, d, first_left_fixed):
    """one left + one right sweep by hand, like DMRG::Solve right after a checkpoint was loaded: site energies in the reference's order"""
    L, out = ctx.L, []
    for index in range(L - 2, 0, -1):
        e, _, _ = d.solve_site(index, 1e-8, 0.0, D, False, not first_left_fixed)
        d.update(index + 1, False)
        out.append(e)
    for index in range(0, L - 2):
        e, _, _ = d.solve_site(index, 1e-8, 0.0, D, True, True)
        d.update(index, True)
        out.append(e)
    return out


@pytest.fixture(scope="module")
def h2o():
    return fixtures.load(os.path.join(ROOT, "tests", "golden", "h2o_631g.npz"))


def test_gpu_resumes_a_reference_checkpoint(h2o, tmp_path):
    pfile = str(tmp_path / "problem.bin")
    _problem_file(h2o, pfile)
    _, loaded = _reference_sweeps(pfile, str(tmp_path), 2)              # writes CheMPS2_MPS0.h5 after every sweep
    assert not loaded and os.path.exists(tmp_path / "CheMPS2_MPS0.h5")
    ctx, d = _gpu_driver(h2o)
    assert d.load_mps(str(tmp_path / "CheMPS2_MPS0.h5")) is False       # the reference stores "not converged" while sweeping
    d.presolve()
    got = _gpu_sweep_site_energies(ctx, d, first_left_fixed=True)
    ref, loaded = _reference_sweeps(pfile, str(tmp_path), 1)            # the reference resumes its own checkpoint for one more sweep
    assert loaded and len(ref) == len(got)
    assert np.abs(np.array(got) - np.array(ref)).max() < 1e-9


def test_reference_resumes_a_gpu_checkpoint(h2o, tmp_path):
    import shutil
    pfile = str(tmp_path / "problem.bin")
    _problem_file(h2o, pfile)
    ctx, d = _gpu_driver(h2o)
    d.random_mps(7)
    d.solve([(D, 1e-14, 2, 0.0, 1e-8)])
    d.save_mps(str(tmp_path / "gpu_state.h5"), converged=False)
    shutil.copy(tmp_path / "gpu_state.h5", tmp_path / "CheMPS2_MPS0.h5")
    ref, loaded = _reference_sweeps(pfile, str(tmp_path), 1)            # the DMRG constructor finds and loads the GPU library's checkpoint
    assert loaded                                                       # (and rewrites the file after its own sweep)
    ctx2, d2 = _gpu_driver(h2o)
    d2.load_mps(str(tmp_path / "gpu_state.h5"))
    d2.presolve()
    got = _gpu_sweep_site_energies(ctx2, d2, first_left_fixed=True)
    assert len(ref) == len(got)
    assert np.abs(np.array(got) - np.array(ref)).max() < 1e-9


def test_config2_sweep_from_the_reference_checkpoint():
    """BASELINE config 2 (N2/cc-pVDZ, D2h, reordered) from a SHARED state: the unmodified reference wrote tests/golden/chk_n2_ccpvdz_D100.h5 after its
    first sweep at D=100 and then resumed from it for one more sweep (tests/golden/chk_n2_ccpvdz_D100.json); the GPU library loads the same
    checkpoint and must reproduce every site energy of that sweep to 1e-9 Eh and the largest discarded weight to 1e-8."""
    import json
    from chemps2_b200 import workloads
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "chk_n2_ccpvdz_D100.json")))
    w = workloads.get("n2_ccpvdz", D=ref["D"])
    ctx = w.context(0)
    d = api.DMRG(ctx)
    d.load_mps(os.path.join(ROOT, "tests", "golden", "chk_n2_ccpvdz_D100.h5"))
    d.presolve()
    L, got, dws = ctx.L, [], [0.0, 0.0]
    for index in range(L - 2, 0, -1):
        e, dw, _ = d.solve_site(index, ref["rtol"], 0.0, ref["D"], False, False)     # first sweep after a load: fixed dimensions (DMRG.cpp:270)
        d.update(index + 1, False)
        got.append(e); dws[0] = max(dws[0], dw)
    for index in range(0, L - 2):
        e, dw, _ = d.solve_site(index, ref["rtol"], 0.0, ref["D"], True, True)
        d.update(index, True)
        got.append(e); dws[1] = max(dws[1], dw)
    assert ref["sites"] == list(range(L - 2, 0, -1)) + list(range(0, L - 2))
    assert np.abs(np.array(got) - np.array(ref["energies"])).max() < 1e-9
    assert abs(dws[0] - ref["max_discarded"][0]) < 1e-8 and abs(dws[1] - ref["max_discarded"][1]) < 1e-8
