"""Excited-state projector: Heff::makeHeff / fillHeffDiag with nLower = 2 (HeffDiagrams1.cpp:65-85,
HeffDiagonal.cpp:621-640).  The golden vectors come from the unmodified reference; the numpy restatement in oracle/ is pinned
against them on the CPU, the CUDA path (b2_heff_set_excitations) against both on the GPU."""
import os
import sys

import numpy as np
import pytest

import cpu_check

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import excitations as oracle_exc  # noqa: E402

TOL = 1e-12


def _close(a, b, tol=TOL):
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


@pytest.mark.parametrize("tag", ["A", "B"])
def test_oracle_restatement_vs_reference(golden, tag):
    vs = [golden[f"{tag}/exc_v0"], golden[f"{tag}/exc_v1"]]
    assert _close(oracle_exc.add_excitations(golden[f"{tag}/rnd_out"], golden[f"{tag}/rnd_in"], vs), golden[f"{tag}/exc_out"])
    assert _close(oracle_exc.add_diagonal_excitations(golden[f"{tag}/diag"], vs), golden[f"{tag}/exc_diag"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["A", "B"])
def test_excitations_gpu_vs_reference(golden, tag):
    ctx, left, right, heff = cpu_check.build_case(golden, tag, device=0)
    vs = [golden[f"{tag}/exc_v0"], golden[f"{tag}/exc_v1"]]
    heff.set_excitations(vs)
    assert _close(heff.apply(golden[f"{tag}/rnd_in"]), golden[f"{tag}/exc_out"])
    assert _close(heff.diag(), golden[f"{tag}/exc_diag"])
    heff.set_excitations([])                       # switched off again: the plain sigma build
    assert _close(heff.apply(golden[f"{tag}/rnd_in"]), golden[f"{tag}/rnd_out"])


@pytest.mark.gpu
def test_excitations_owner_shards_sum(golden):
    """state s is applied by GPU s % world: the shard sums still give the reference result"""
    ref = golden["A/exc_out"]
    tot = np.zeros_like(ref)
    vs = [golden["A/exc_v0"], golden["A/exc_v1"]]
    for r in range(2):
        ctx, left, right, heff = cpu_check.build_case(golden, "A", device=0, world=2, rank=r)
        heff.set_excitations(vs)
        tot += heff.apply(golden["A/rnd_in"])
    assert _close(tot, ref)
