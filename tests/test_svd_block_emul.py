"""The block-Jacobi SVD of the device (k_jacobi_block, b2_svd.cu: pairs of 8-column blocks, 16 x 16 Gram matrix, one cyclic Jacobi sweep
in shared memory, rotation applied to W and V) emulated phase by phase on the CPU (oracle/svd_block_emul.cpp) against LAPACK: the
round-robin block schedule reaches every column pair, the convergence rule (Hestenes criterion on the fresh Gram matrix) stops at
LAPACK's singular values, and the factors reconstruct the matrix with orthonormal vectors — for column counts below, at and above the
block size, odd / even numbers of blocks, rank-deficient, graded and zero matrices."""
import ctypes as C

import numpy as np
import pytest

from cpu_check import oracle_lib


def _svd_block(a):
    o = oracle_lib()
    dp = C.POINTER(C.c_double)
    o.b2o_svd_block.argtypes = [C.c_int, C.c_int, dp, dp]
    o.b2o_svd_block.restype = C.c_int
    R, Cc = a.shape
    w = np.asfortranarray(a, dtype=np.float64).copy(order="F")
    v = np.zeros((Cc, Cc), order="F")
    sweeps = o.b2o_svd_block(R, Cc, w.ctypes.data_as(dp), v.ctypes.data_as(dp))
    return w, v, sweeps


def _cases():
    rng = np.random.default_rng(11)
    out = [(f"{r}x{c}", rng.standard_normal((r, c))) for r, c in
           [(1, 1), (9, 1), (5, 2), (7, 7), (8, 8), (30, 9), (16, 16), (40, 17), (33, 33), (129, 77), (100, 100), (260, 130)]]
    out.append(("rank6of45", rng.standard_normal((60, 6)) @ rng.standard_normal((6, 45))))
    out.append(("graded80", rng.standard_normal((80, 80)) @ np.diag(10.0 ** -np.linspace(0, 14, 80)) @ rng.standard_normal((80, 80))))
    out.append(("zeros", np.zeros((6, 4))))
    return out


@pytest.mark.parametrize("name,a", _cases(), ids=[n for n, _ in _cases()])
def test_block_jacobi_emulation_vs_lapack(name, a):
    w, v, sweeps = _svd_block(a)
    assert sweeps < 60                                                          # converged, not capped
    s = np.linalg.norm(w, axis=0)
    ref = np.linalg.svd(a, compute_uv=False)
    scale = max(1.0, ref[0]) if ref.size else 1.0
    assert np.abs(np.sort(s)[::-1] - ref).max() <= 1e-12 * scale                # LAPACK's singular values
    assert np.abs(w @ v.T - a).max() <= 1e-12 * scale                           # A = W V^T
    assert np.abs(v.T @ v - np.eye(a.shape[1])).max() <= 1e-12                  # V orthogonal
    nz = s > 1e-10 * scale
    if nz.any():
        u = w[:, nz] / s[nz]
        assert np.abs(u.T @ u - np.eye(nz.sum())).max() <= 1e-11                # columns of W orthogonal
