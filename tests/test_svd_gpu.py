"""Batched device SVD (b2_svd_batch, b2_svd.cu) = the decomposition step of Sobject::Split (dgesdd_ at Sobject.cpp:412-419):
singular values against LAPACK (numpy), factors through reconstruction and orthonormality; tall, wide, square, rank-deficient,
single-row/column and odd-sized matrices in ONE batch (they progress together on the device)."""
import numpy as np
import pytest

from chemps2_b200 import api

pytestmark = pytest.mark.gpu


def test_svd_batch_vs_lapack():
    rng = np.random.default_rng(5)
    shapes = [(1, 1), (1, 7), (9, 1), (5, 5), (33, 17), (17, 33), (64, 64), (129, 77), (40, 200), (257, 255)]
    mats = [rng.standard_normal(s) for s in shapes]
    lowrank = rng.standard_normal((60, 6)) @ rng.standard_normal((6, 45))          # rank 6 of 45
    graded = rng.standard_normal((80, 80)) @ np.diag(10.0 ** -np.linspace(0, 14, 80)) @ rng.standard_normal((80, 80))
    mats += [lowrank, graded, np.zeros((6, 4))]
    ctx = api.Context(0)
    for (u, s, vt), a in zip(api.svd_batch(ctx, mats), mats):
        ref = np.linalg.svd(a, compute_uv=False)
        scale = max(1.0, ref[0]) if ref.size else 1.0
        assert s.size < 2 or np.all(np.diff(s) <= 1e-300 + 1e-14 * scale)           # decreasing
        assert np.abs(s - ref).max() <= 1e-12 * scale                              # LAPACK's singular values
        assert np.abs((u * s) @ vt - a).max() <= 1e-12 * scale                     # a = U diag(s) V^T
        nz = s > 1e-10 * scale                                                     # vectors of non-negligible singular values
        if nz.any():
            assert np.abs(u[:, nz].T @ u[:, nz] - np.eye(nz.sum())).max() <= 1e-11
            assert np.abs(vt[nz] @ vt[nz].T - np.eye(nz.sum())).max() <= 1e-11


def test_svd_batch_deterministic():
    rng = np.random.default_rng(9)
    mats = [rng.standard_normal((70, 50)), rng.standard_normal((31, 90))]
    ctx = api.Context(0)
    r1, r2 = api.svd_batch(ctx, mats), api.svd_batch(ctx, mats)
    for (u1, s1, v1), (u2, s2, v2) in zip(r1, r2):
        assert np.array_equal(u1, u2) and np.array_equal(s1, s2) and np.array_equal(v1, v2)
