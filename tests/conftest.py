import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the product library and the CPU checker are built in-tree; build them if this checkout has not been built yet
    built = [os.path.join(ROOT, "chemps2_b200", "libchemps2_b200.so"), os.path.join(ROOT, "oracle", "libb2oracle.so"),
             os.path.join(ROOT, "tests", "cpp", "_bin", "dmrg_caller")]
    if not all(os.path.exists(p) for p in built):
        subprocess.run(["make", "-j8"], cwd=ROOT, check=True, stdout=subprocess.DEVNULL)


GOLDEN = sorted(p for p in glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))
                if not p.endswith("wigner.npz") and not os.path.basename(p).startswith("problem_") and not os.path.basename(p).startswith("trace_"))

TRACES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "trace_*.npz")))


@pytest.fixture(scope="session", params=TRACES, ids=[os.path.basename(p)[:-4] for p in TRACES])
def trace(request):
    """step-by-step record of noisy sweeps of the unmodified reference from its seeded random MPS (ref_driver `trace`)"""
    from chemps2_b200 import fixtures
    return fixtures.load(request.param)


@pytest.fixture(scope="session", params=GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def golden(request):
    from chemps2_b200 import fixtures
    return fixtures.load(request.param)


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """cases added after the last GPU session of the round (the test12 pairing-model, test9 momentum-space and N2+ doublet fixtures, the per-object Sobject entries) run last, so
    that under `-x` a surprise there cannot hide the results of the long-proven tests"""
    late = [it for it in items if "pairing8" in it.nodeid or "hubbard3x3" in it.nodeid or "cation_doublet" in it.nodeid or "test_zz_" in it.nodeid]
    if late:
        ids = {id(it) for it in late}
        items[:] = [it for it in items if id(it) not in ids] + late
