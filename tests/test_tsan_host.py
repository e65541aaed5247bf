"""ThreadSanitizer runs of the host-side concurrency of the library (opt-in: B2_RUN_TSAN=1, about two minutes):
  pool_stress  — several callers of parallel_run at once + nested use (b2_core.cpp worker pool: several open jobs)
  cache_stress — the host block cache under concurrent acquire / release with a small budget (overflow path)
  plan_race    — what the sweep driver's plan prefetch does: the sigma plan of the next site on a helper thread while the (sharded)
                 update plan is built on the calling thread, planning-only contexts through the C ABI
Every host translation unit is compiled with -fsanitize=thread; the CUDA objects of the in-tree build are linked as they are."""
import glob
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "chemps2_b200", "csrc")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")

pytestmark = pytest.mark.skipif(os.environ.get("B2_RUN_TSAN") != "1", reason="opt-in: B2_RUN_TSAN=1")


def _run(cmd, **kw):
    res = subprocess.run(cmd, capture_output=True, text=True, **kw)
    assert res.returncode == 0, (cmd, res.stdout[-3000:], res.stderr[-3000:])
    return res


def test_host_concurrency_under_tsan(tmp_path):
    flags = ["-O1", "-g", "-std=c++17", "-fsanitize=thread", f"-I{SRC}", f"-I{ROOT}/include", f"-I{CUDA}/include"]
    core = os.path.join(SRC, "b2_core.cpp")
    for name in ("pool_stress", "cache_stress"):
        exe = str(tmp_path / name)
        _run(["g++", *flags, os.path.join(ROOT, "tests", "cpp", "tsan", name + ".cpp"), core, "-o", exe, "-lpthread"])
        out = _run([exe], timeout=900)
        assert "ok" in out.stdout and "ThreadSanitizer" not in out.stderr
    objs = []
    procs = []
    for cpp in sorted(glob.glob(os.path.join(SRC, "*.cpp"))):
        obj = str(tmp_path / (os.path.basename(cpp)[:-4] + ".o"))
        objs.append(obj)
        procs.append(subprocess.Popen(["g++", *flags, "-fPIC", "-c", cpp, "-o", obj]))
    assert all(p.wait() == 0 for p in procs)
    cu_objs = [os.path.join(SRC, n) for n in ("b2_kernels.o", "b2_blas1.o", "b2_svd.o")]
    assert all(os.path.exists(o) for o in cu_objs), "run make first"
    exe = str(tmp_path / "plan_race")
    _run(["g++", *flags, os.path.join(ROOT, "tests", "cpp", "tsan", "plan_race.cpp"), *objs, *cu_objs, "-o", exe, f"-L{CUDA}/lib64", "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    out = _run([exe, "100"], env=dict(os.environ, B2_PLAN_THREADS="6"), timeout=1800)
    assert out.stdout.strip().endswith("ok") and "ThreadSanitizer" not in out.stderr
