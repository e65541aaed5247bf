"""Host-side plan building on several threads (b2_sigma_plan.cpp fragments stitched in block order, b2_compile.cpp segments merged
wave by wave) must give the same plan as the sequential build: identical term list (order included) and, for the compiled work
lists, the same arithmetic (checked through the work-list emulator against the reference's sigma vector)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import sys, os
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import cpu_check
from chemps2_b200 import fixtures
fx = fixtures.load(os.path.join({root!r}, "tests", "golden", "h2o_631g.npz"))
out = []
for tag in ("A", "B"):
    ctx, left, right, heff = cpu_check.build_case(fx, tag, options={{"parallel_min_terms": 16}})
    terms, nt, parts, npp, psize = heff.export()
    st = heff.stats()
    sig = cpu_check.emulate_worklists(ctx, left, right, heff, fx[tag + "/rnd_in"])
    err = float(np.abs(sig - fx[tag + "/rnd_out"]).max() / max(1.0, np.abs(fx[tag + "/rnd_out"]).max()))
    out.append((int(nt), st["flops_ref"], st["terms"], err))
print("B2PLAN", out)
"""


def _run(threads):
    env = dict(os.environ, B2_PLAN_THREADS=str(threads))
    res = subprocess.run([sys.executable, "-c", WORKER.format(root=ROOT)], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("B2PLAN")][-1]
    return eval(line[len("B2PLAN"):])


def test_parallel_plan_equals_sequential_plan():
    seq, par = _run(1), _run(5)
    for (n1, f1, t1, e1), (n2, f2, t2, e2) in zip(seq, par):
        assert n1 == n2 and t1 == t2                      # same number of terms
        assert abs(f1 - f2) <= 1e-9 * f1                  # same algorithmic FLOP count (summed per thread, so only rounding may differ)
        assert e1 < 1e-12 and e2 < 1e-12                  # both reproduce Heff::makeHeff through the emulated work lists


POOL_WORKER = r"""
import sys, os, threading
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import cpu_check
from chemps2_b200 import fixtures
fx = fixtures.load(os.path.join({root!r}, "tests", "golden", "h2o_631g.npz"))

def one(tag):
    ctx, left, right, heff = cpu_check.build_case(fx, tag, options={{"parallel_min_terms": 16}})
    sig = cpu_check.emulate_worklists(ctx, left, right, heff, fx[tag + "/rnd_in"])
    return float(np.abs(sig - fx[tag + "/rnd_out"]).max() / max(1.0, np.abs(fx[tag + "/rnd_out"]).max()))

errs = [one("A")]                                   # creates the parked workers
res = {{}}
ths = [threading.Thread(target=lambda t=t: res.__setitem__(t, one("AB"[t % 2]))) for t in range(4)]
for th in ths: th.start()                           # concurrent builds: one owns the pool, the others run their pieces inline
for th in ths: th.join()
errs += [res[t] for t in range(4)]
pid = os.fork()                                     # the child inherits the pool object but not its threads
if pid == 0:
    ok = one("B") < 1e-12
    os._exit(0 if ok else 3)
_, status = os.waitpid(pid, 0)
errs.append(one("B"))                               # and the parent's pool still works
print("B2POOL", errs, os.WEXITSTATUS(status))
"""


def test_worker_pool_concurrent_callers_and_fork():
    """the parked host workers (b2_core.cpp parallel_run) are shared by every caller: concurrent builds fall back to inline pieces,
    a forked child starts its own workers; every plan still reproduces the reference's sigma vector"""
    env = dict(os.environ, B2_PLAN_THREADS="4")
    res = subprocess.run([sys.executable, "-c", POOL_WORKER.format(root=ROOT)], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("B2POOL")][-1]
    payload = line[len("B2POOL"):].strip()
    errs, child = eval("(" + payload.rsplit("]", 1)[0] + "]," + payload.rsplit("]", 1)[1] + ")")
    assert child == 0
    assert len(errs) == 6 and max(errs) < 1e-12, errs


UPDATE_WORKER = r"""
import sys, os
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import cpu_check
from chemps2_b200 import fixtures
out = []
for name in ("h2o_631g", "n2_sto3g_quintet_b1u"):
    fx = fixtures.load(os.path.join({root!r}, "tests", "golden", name + ".npz"))
    for which in ("UR", "UL"):
        ctx, old, new, upd, t, expected = cpu_check.build_update_case(fx, which)
        arena = cpu_check.emulate_update(old, new, upd, t)
        st = upd.stats()
        out.append((int(st["terms"]), int(st["mix_terms"]), int(st["presums"]), st["flops_ref"], float(np.abs(arena).sum())))
print("B2UPD", out)
"""


def test_parallel_update_plan_equals_sequential_plan():
    """b2_update_plan.cpp enumerates the new operators on several host threads; stitched in operator order (pre-sum arena offsets shifted)
    the plan is the sequential one: same term / mixing-term / pre-sum counts and the emulated arena agrees to the last bit of its abs-sum
    up to summation order"""
    def run(threads):
        env = dict(os.environ, B2_PLAN_THREADS=str(threads))
        res = subprocess.run([sys.executable, "-c", UPDATE_WORKER.format(root=ROOT)], capture_output=True, text=True, env=env, timeout=600)
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
        return eval([ln for ln in res.stdout.splitlines() if ln.startswith("B2UPD")][-1][len("B2UPD"):])
    seq, par = run(1), run(6)
    assert len(seq) == 4
    for (t1, m1, p1, f1, s1), (t2, m2, p2, f2, s2) in zip(seq, par):
        assert (t1, m1, p1) == (t2, m2, p2)
        assert abs(f1 - f2) <= 1e-9 * f1
        assert abs(s1 - s2) <= 1e-10 * max(1.0, abs(s1))


PREFETCH_WORKER = r"""
import sys, os, threading
import ctypes as C
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import cpu_check
from chemps2_b200 import api, workloads
from chemps2_b200._lib import Worklists, check, lib

w = workloads.get("n2_ccpvdz", D=300)
ctx = w.context(-1)
w.apply_distribution(ctx, "gauss")
ctx.set_option("parallel_min_terms", 2000)

def sigma_digest(site, left, right):
    heff = api.Heff(ctx, site, left, right)
    wl = Worklists()
    check(lib.b2_heff_worklists(heff.h, C.byref(wl)))
    return cpu_check.worklist_digest(wl)

def update_digest(index, old, new):
    upd = api.Update(ctx, index, True, old, new)
    out = []
    for p in (0, 1):
        wl = Worklists()
        check(lib.b2_update_worklists(upd.h, p, C.byref(wl)))
        out.append(cpu_check.worklist_digest(wl))
    return out

res = []
for s in (8, 12, 16):
    old = api.OpSet(ctx, s, True)
    fresh = api.OpSet(ctx, s + 1, True)
    right = api.OpSet(ctx, s + 3, False)
    seq = (sigma_digest(s + 1, fresh, right), update_digest(s, old, fresh))
    box = {{}}
    th = threading.Thread(target=lambda: box.__setitem__("sigma", sigma_digest(s + 1, fresh, right)))   # ctypes drops the GIL: real concurrency
    th.start()
    upd = update_digest(s, old, fresh)
    th.join()
    res.append(seq == (box["sigma"], upd))
print("B2PREFETCH", res)
"""


def test_next_sigma_plan_concurrent_with_update_plan():
    """what the sweep driver's plan prefetch does on the host (b2_capi_dmrg.cpp dmrg_prefetch_start): the sigma plan of the NEXT site pair
    is built on a helper thread while the update plan that produces its left operator set is built on the calling thread; both share
    the worker pool (b2_core.cpp: several open jobs) and read the same bookkeeper / operator-set layouts.  Every compiled list must be
    bit-identical to the one built alone."""
    env = dict(os.environ, B2_PLAN_THREADS="6")
    res = subprocess.run([sys.executable, "-c", PREFETCH_WORKER.format(root=ROOT)], capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("B2PREFETCH")][-1]
    assert eval(line[len("B2PREFETCH"):]) == [True, True, True]


RECYCLE_WORKER = r"""
import sys, os
import ctypes as C
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import cpu_check
from chemps2_b200 import api, workloads
from chemps2_b200._lib import Worklists, check, lib

def digests(name, D, dist, site):
    w = workloads.get(name, D=D)
    ctx = w.context(-1)
    w.apply_distribution(ctx, dist)
    left = api.OpSet(ctx, site, True)
    right = api.OpSet(ctx, site + 2, False)
    heff = api.Heff(ctx, site, left, right)
    wl = Worklists()
    check(lib.b2_heff_worklists(heff.h, C.byref(wl)))
    out = [cpu_check.worklist_digest(wl)]
    new = api.OpSet(ctx, site + 1, True)
    upd = api.Update(ctx, site, True, left, new)
    for p in (0, 1):
        uw = Worklists()
        check(lib.b2_update_worklists(upd.h, p, C.byref(uw)))
        out.append(cpu_check.worklist_digest(uw))
    return out

a1 = digests("n2_ccpvdz", 500, "gauss", 12)
b = digests("synth40", 300, "gauss", 19)          # other sizes, other contents: the recycled blocks now hold ITS lists
a2 = digests("n2_ccpvdz", 500, "gauss", 12)
print("B2RECYCLE", a1 == a2, a1 != b)
"""


def test_plans_do_not_depend_on_recycled_memory():
    """work lists and term arrays live in blocks recycled through the host block cache (b2_core.cpp host_block_acquire) and are sized
    without a zero fill (NoInitAlloc): a plan built on blocks that still hold ANOTHER plan's lists must come out bit-identical to the
    one built on fresh memory — every field of every list entry is written by the builder"""
    for threads in ("1", "6"):
        env = dict(os.environ, B2_PLAN_THREADS=threads)
        res = subprocess.run([sys.executable, "-c", RECYCLE_WORKER.format(root=ROOT)], capture_output=True, text=True, env=env, timeout=900)
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
        line = [ln for ln in res.stdout.splitlines() if ln.startswith("B2RECYCLE")][-1]
        assert line.split()[1:] == ["True", "True"], line
