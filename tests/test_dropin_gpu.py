"""The DROP-IN library: the reference's OWN tests and its `chemps2` binary, built from the unmodified reference sources and linked against
dropin/_build/libchemps2.so.3, in which Heff::SolveDAVIDSON / makeHeff / fillHeffDiag and DMRG::updateMovingRight / updateMovingLeft are
the CUDA library's (dropin/chemps2_b200_shim.cpp over the C ABI; dropin/build_dropin.sh).  The test programs are the reference's
tests/testN.cpp.in verbatim (only the data path substituted): they check DMRG against FCI (tests 1-4, 9), literal known answers
(tests 5 and 12), 2-RDM energies, DMRG-CASSCF (tests 6, 8: the reference's CASSCF driver on top of the GPU sweep) and — tests 10, 11 —
3-RDM / 4-RDM contractions; they return 0 on success.  The shim's exit report
proves that the sigma builds and operator updates of each run went through the GPU."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "dropin", "_build")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(os.path.join(BUILD, "libchemps2.so.3")), reason="dropin/_build not built (needs /root/reference at build time)")]


def _run(cmd, timeout):
    env = dict(os.environ, CHEMPS2_B200_VERBOSE="1", OMP_NUM_THREADS=str(min(os.cpu_count() or 1, 16)), OPENBLAS_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd="/tmp")
    m = re.search(r"chemps2_b200 drop-in: (\d+) Davidson solves \((\d+) sigma builds\).*?(\d+) operator updates", res.stderr)
    return res, m


# the reference's tests that exercise the sweep path: DMRG vs FCI (1-4, 9), known answers (5, 12), DMRG-CASSCF (6, 8), 3-RDM / 4-RDM
# contractions (10, 11).  test7 (CASSCF with the FCI solver: no DMRG object) and test13 run with B2_DROPIN_ALL=1.
FAST = [1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 12]
SLOW = [7, 13]


@pytest.mark.parametrize("n", FAST + (SLOW if os.environ.get("B2_DROPIN_ALL") else []))
def test_reference_test_passes_on_the_dropin(n):
    exe = os.path.join(BUILD, f"test{n}")
    if not os.path.exists(exe):
        pytest.skip(f"test{n} not built")
    res, m = _run([exe], 1500)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    assert f"Did test {n} succeed : yes" in res.stdout
    assert m is not None, res.stderr[-800:]
    solves, sigma, updates = (int(x) for x in m.groups())
    if n not in SLOW:
        assert solves > 0 and sigma >= solves and updates > 0      # the hot path really ran through the CUDA library


def test_chemps2_binary_on_the_dropin():
    """`chemps2 --file=tests/test2.input` (H2O/6-31G DMRG-CI, schedule D = 240 ... 30 with noise): the unmodified reference prints
    -76.1212850381306 as the minimum energy of the run (oracle/_ref/chemps2 on the host); noise is seeded from the clock, the converged
    energy does not depend on it"""
    res, m = _run([os.path.join(BUILD, "chemps2"), "--file=" + os.path.join(BUILD, "tests", "test2.input")], 1500)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    energies = [float(x) for x in re.findall(r"Minimum energy encountered during all instructions = (-?[\d.]+)", res.stdout)]
    assert energies and abs(min(energies) - (-76.1212850381306)) < 1e-8, energies[-3:]
    assert m is not None and int(m.group(1)) > 0
