"""Error behaviour of the C ABI (SURVEY.md 8(b) "errors"): the reference asserts or returns NULL / -1 sentinels on impossible input
(asserts are compiled out in its default Release build, i.e. undefined behaviour); every entry point here returns a code and leaves a
message for b2_last_error(), range queries return the reference's sentinels, and nothing crashes.  All on planning-only contexts
(device = -1): no GPU is touched, and every compute entry refuses to run instead of falling back to the CPU."""
import ctypes as C
import os

import numpy as np
import pytest

import cpu_check
from chemps2_b200 import api, fixtures
from chemps2_b200._lib import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ERR_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_STATE = -1, -2, -3, -4


def _ip(a):
    return np.ascontiguousarray(a, dtype=np.int32).ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.fixture()
def h2o():
    return fixtures.load(os.path.join(ROOT, "tests", "golden", "h2o_631g.npz"))


def _ctx():
    h = C.c_void_p()
    assert lib.b2_ctx_create(-1, C.byref(h)) == 0
    return h


def test_context_device_out_of_range():
    h = C.c_void_p()
    assert lib.b2_ctx_create(4096, C.byref(h)) == ERR_NO_DEVICE
    assert b"not available" in lib.b2_last_error()
    assert lib.b2_ctx_create(-1, None) == ERR_ARG


@pytest.mark.parametrize("bad, word", [
    (dict(group=9), b"group"), (dict(irrep=4), b"irrep"), (dict(N=1), b"N must be"), (dict(twoS=-2), b"TwoS"),
    (dict(N=40), b"N > 2*L"), (dict(twoS=1), b"% 2"), (dict(twoS=14, N=12), b"TwoS > L"), (dict(orb=7), b"orbital irrep"),
])
def test_problem_consistency_checks(h2o, bad, word):
    """Problem::checkConsistency (Problem.cpp:386-430) + Hamiltonian's irrep range asserts, as error codes"""
    L, group, N, twoS, irrep = [int(x) for x in h2o["problem/hdr"]]
    irr = np.array(h2o["problem/orb_irrep"], dtype=np.int32)
    if "orb" in bad:
        irr = irr.copy(); irr[3] = bad["orb"]
    a = dict(group=group, N=N, twoS=twoS, irrep=irrep); a.update({k: v for k, v in bad.items() if k != "orb"})
    ctx = _ctx()
    mx = np.ascontiguousarray(h2o["problem/mx"])
    rc = lib.b2_problem_set(ctx, L, a["group"], a["N"], a["twoS"], a["irrep"], _ip(irr), _dp(mx), 0.0)
    assert rc == ERR_ARG and word in lib.b2_last_error(), lib.b2_last_error()
    assert lib.b2_bk_init(ctx, 10) == ERR_STATE          # still no problem
    lib.b2_ctx_destroy(ctx)


def test_unreachable_target_sector(h2o):
    """SyBookkeeper::IsPossible: all orbitals A1 but a B1 target cannot be reached"""
    L, group, N, twoS, _ = [int(x) for x in h2o["problem/hdr"]]
    ctx = _ctx()
    mx = np.ascontiguousarray(h2o["problem/mx"])
    assert lib.b2_problem_set(ctx, L, group, N, twoS, 2, _ip(np.zeros(L, dtype=np.int32)), _dp(mx), 0.0) == 0
    assert lib.b2_bk_init(ctx, 10) == ERR_ARG and b"not reachable" in lib.b2_last_error()
    lib.b2_ctx_destroy(ctx)


def test_bookkeeper_sentinels(h2o):
    """outside the tables: dimension 0 (SyBookkeeper::gCurrentDim, SyBookkeeper.cpp:271-280), empty ranges; no bookkeeper: the same"""
    ctx = api.context_from_fixture(h2o, "A")
    L = int(h2o["problem/hdr"][0])
    assert ctx.dim(3, 99, 0, 0) == 0 and ctx.dim(-1, 0, 0, 0) == 0 and ctx.dim(L + 1, 0, 0, 0) == 0 and ctx.dim(3, 2, 1, 0) == 0
    assert ctx.fcidim(3, 2, 0, 9) == 0
    assert lib.b2_bk_nmin(ctx.h, L + 5) > lib.b2_bk_nmax(ctx.h, L + 5)                    # empty range
    assert lib.b2_bk_twosmin(ctx.h, 3, 99) > lib.b2_bk_twosmax(ctx.h, 3, 99)
    assert lib.b2_bk_set_dim(ctx.h, 3, 99, 0, 0, 5) == 0 and ctx.dim(3, 99, 0, 0) == 0     # ignored, like SyBookkeeper::SetDim
    assert lib.b2_bk_set_dim(ctx.h, 3, 2, 0, 0, -1) == ERR_ARG
    assert lib.b2_tensor_t_size(ctx.h, L) == -1 and lib.b2_sobject_size(ctx.h, L - 1) == -1 and lib.b2_sobject_nkappa(ctx.h, -1) == -1
    assert lib.b2_sobject_table(ctx.h, 0, None, None) == ERR_ARG
    empty = _ctx()
    assert lib.b2_bk_dim(empty, 0, 0, 0, 0) == 0 and lib.b2_bk_nmax(empty, 0) < lib.b2_bk_nmin(empty, 0)
    assert lib.b2_tensor_t_size(empty, 0) == -1
    assert lib.b2_bk_set_dim(empty, 0, 0, 0, 0, 1) == ERR_STATE
    lib.b2_ctx_destroy(empty)


def test_operator_set_and_plan_argument_checks(h2o):
    ctx, left, right, heff = cpu_check.build_case(h2o, "A")
    L = int(h2o["problem/hdr"][0])
    site = int(h2o["A/hdr"][0])
    out = C.c_void_p()
    assert lib.b2_opset_create(ctx.h, 0, 1, C.byref(out)) == ERR_ARG               # no operators live on the outer boundaries
    assert lib.b2_opset_create(ctx.h, L, 0, C.byref(out)) == ERR_ARG
    assert left.find(api.K_L if hasattr(api, "K_L") else 0, 99, 99) == -1          # TensorOperator-style sentinel
    assert lib.b2_opset_info(left.h, 10 ** 6, None, None, None, None) == ERR_ARG
    assert lib.b2_opset_upload(left.h, 0, None) == ERR_ARG and lib.b2_opset_download(left.h, -1, None) == ERR_ARG
    # a sigma plan wants the moving-right set of boundary `site` and the moving-left set of boundary `site + 2`
    assert lib.b2_heff_create(ctx.h, site, right.h, left.h, 1, 0, C.byref(out)) == ERR_ARG and b"left operator set" in lib.b2_last_error()
    assert lib.b2_heff_create(ctx.h, L - 1, left.h, right.h, 1, 0, C.byref(out)) == ERR_ARG
    assert lib.b2_heff_create(ctx.h, site, left.h, right.h, 2, 2, C.byref(out)) == ERR_ARG   # rank >= world
    assert lib.b2_heff_create(ctx.h, site, left.h, right.h, 0, 0, C.byref(out)) == ERR_ARG
    # the update of boundary index -> index + 1 wants the old set at `index` and the new one at `index + 1`, same direction
    assert lib.b2_update_create(ctx.h, site, 1, left.h, left.h, C.byref(out)) != 0
    assert lib.b2_update_create(ctx.h, site, 1, right.h, None, C.byref(out)) != 0


def test_compute_entries_refuse_without_device(h2o):
    """no CPU fallback anywhere: every entry that would launch a kernel fails with B2_ERR_NO_DEVICE on a planning-only context"""
    ctx, left, right, heff = cpu_check.build_case(h2o, "A")
    n = heff.n
    v = np.zeros(n)
    e, nm = C.c_double(), C.c_int()
    assert lib.b2_heff_apply(heff.h, _dp(v), _dp(v)) == ERR_NO_DEVICE
    assert lib.b2_heff_diag(heff.h, _dp(v)) == ERR_NO_DEVICE
    assert lib.b2_heff_solve(heff.h, _dp(v), 1e-5, C.byref(e), C.byref(nm)) == ERR_NO_DEVICE
    out = C.c_void_p()
    assert lib.b2_dmrg_create(ctx.h, C.byref(out)) == ERR_NO_DEVICE and b"no CPU fallback" in lib.b2_last_error()
    assert lib.b2_opset_offload(left.h) == ERR_NO_DEVICE
    assert lib.b2_opset_offload_file(left.h, b"/tmp/b2_never_written.bin") == ERR_NO_DEVICE
    dav = C.c_void_p()
    assert lib.b2_davidson_create(ctx.h, n, 32, 3, 1e-5, 1e-12, C.byref(dav)) == ERR_NO_DEVICE and b"no CPU fallback" in lib.b2_last_error()
    t = C.c_double()
    assert lib.b2_probe_fp64(ctx.h, 1, C.byref(t)) == ERR_NO_DEVICE
    assert lib.b2_ctx_set_stream(ctx.h, None) == ERR_NO_DEVICE
    m, nn = np.array([2], dtype=np.int32), np.array([2], dtype=np.int32)
    assert lib.b2_svd_batch(ctx.h, 1, _ip(m), _ip(nn), None, None, None, None) != 0


def test_null_handles_do_not_crash():
    assert lib.b2_heff_veclength(None) <= 0
    assert lib.b2_opset_count(None) == 0 and lib.b2_opset_find(None, 0, 0, 0) == -1 and lib.b2_opset_resident(None) == 0
    assert lib.b2_dmrg_mps_size(None, 0) == -1 and lib.b2_dmrg_num_lower_states(None) == 0
    assert lib.b2_dmrg_sweep_info(None, None) == ERR_ARG and lib.b2_dmrg_presolve(None) == ERR_ARG
    assert lib.b2_ctx_device(None) == -1
    assert lib.b2_opset_offload_file(None, None) == ERR_ARG and lib.b2_dmrg_set_spill_dir(None, None) == ERR_ARG and lib.b2_dmrg_srand(None, 1) == ERR_ARG
    assert lib.b2_davidson_fetch(None, None, None, None) == ERR_ARG and lib.b2_davidson_num_multiplications(None) == 0
    assert lib.b2_update_num_mix_flat(None) == 0 and lib.b2_rand_stream(1, -1, None) == ERR_ARG
    assert lib.b2_dmrg_set_plan_prefetch(None, 1) == ERR_ARG and lib.b2_dmrg_plan_prefetched(None) == 0
    lib.b2_davidson_destroy(None)
    lib.b2_ctx_destroy(None); lib.b2_opset_destroy(None); lib.b2_heff_destroy(None); lib.b2_update_destroy(None); lib.b2_dmrg_destroy(None); lib.b2_twodm_destroy(None)
