"""Sobject::Join (Sobject.cpp:212-258) through its own entry point: the two-site object built from the site tensors of the reference's MPS
must equal the S storage the reference's Join produced (golden key <tag>/joined, program convention)."""
import ctypes as C

import numpy as np
import pytest

import cpu_check
from chemps2_b200 import api
from chemps2_b200._lib import Worklists, check, lib


def _inputs(golden, tag):
    site = int(golden[tag + "/hdr"][0])
    return site, golden[f"{tag}/mps/{site}"], golden[f"{tag}/mps/{site + 1}"], golden[tag + "/joined"]


@pytest.mark.parametrize("tag", ["A", "B"])
def test_join_worklists_vs_reference_cpu(golden, tag):
    """CPU: the compiled Join work lists executed by the emulator in oracle/"""
    site, tl, tr, ref = _inputs(golden, tag)
    ctx = api.context_from_fixture(golden, tag)
    j = api.Join(ctx, site)
    assert j.n == ref.size
    wl = Worklists()
    check(lib.b2_join_worklists(j.h, C.byref(wl)))
    o = cpu_check.oracle_lib()
    dp = C.POINTER(C.c_double)
    o.b2o_run_worklists.argtypes = [C.POINTER(Worklists), dp, dp, dp, dp, dp, C.c_int64]
    tl = np.ascontiguousarray(tl, dtype=np.float64); tr = np.ascontiguousarray(tr, dtype=np.float64)
    dummy = np.zeros(1)
    out = np.zeros(ref.size)
    as_dp = lambda a: a.ctypes.data_as(dp)
    o.b2o_run_worklists(C.byref(wl), as_dp(tl), as_dp(tr), as_dp(dummy), as_dp(dummy), as_dp(out), out.size)
    assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    with pytest.raises(api.B2Error):          # planning-only context: no compute
        j.run(tl, tr)


def test_join_argument_checks(golden):
    ctx = api.context_from_fixture(golden, "A")
    L = int(golden["problem/hdr"][0])
    out = C.c_void_p()
    assert lib.b2_join_create(ctx.h, L - 1, C.byref(out)) == -1
    assert lib.b2_join_create(ctx.h, -1, C.byref(out)) == -1
    assert lib.b2_join_worklists(None, None) == -1 and lib.b2_join_run(None, None, None, None) == -1
    lib.b2_join_destroy(None)
