"""Operator-update parity: every renormalized operator produced by the UpdatePlan (L, S0/S1, F0/F1, A/B/C/D, Q, X; both sweep
directions) against the tensors the reference's DMRG::updateMovingRight / updateMovingLeft produced from the same inputs."""
import numpy as np
import pytest

import cpu_check
from chemps2_b200 import api


def _compare(arena_of, new, expected):
    sl = {(k, i, j): (off, size) for k, i, j, off, size in cpu_check.op_slices(new)}
    assert len(sl) == len(expected)
    seen = set()
    for kind, i, j, data in expected:
        off, size = sl[(kind, i, j)]
        assert size == data.size
        if size == 0:
            continue
        got = arena_of(kind, i, j, off, size)
        scale = max(1.0, np.abs(data).max())
        assert np.abs(got - data).max() <= 1e-12 * scale, (api.KIND_NAMES[kind], i, j)
        seen.add(api.KIND_NAMES[kind])
    return seen


@pytest.mark.parametrize("options", [{}, {"work_budget": 4096, "chunk_k": 32}], ids=["default", "tiny-waves"])
@pytest.mark.parametrize("which", ["UR", "UL"])
def test_update_worklists_vs_reference_cpu(golden, which, options):
    """CPU: the compiled update work lists executed by the emulator in oracle/ (no GPU needed to pin the plan)"""
    ctx, old, new, upd, t, expected = cpu_check.build_update_case(golden, which, options=options)
    arena = cpu_check.emulate_update(old, new, upd, t)
    seen = _compare(lambda k, i, j, off, size: arena[off:off + size], new, expected)
    assert {"S0", "F0", "F1", "C", "X"} <= seen
    st = upd.stats()
    assert st["terms"] > 0 and st["flops_ref"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["UR", "UL"])
def test_update_vs_reference_gpu(golden, which):
    """GPU, through the C ABI: b2_update_run == DMRG::updateMovingRight/Left"""
    ctx, old, new, upd, t, expected = cpu_check.build_update_case(golden, which, device=0)
    upd.run(t)
    idx = {}
    for n in range(len(new)):
        k, i, j, _ = new.info(n)
        idx[(k, i, j)] = n
    seen = _compare(lambda k, i, j, off, size: new.download(idx[(k, i, j)]), new, expected)
    assert {"S0", "F0", "F1", "C", "X"} <= seen
    # deterministic: a second run gives bit-identical operators
    first = new.download(idx[(10, -1, -1)]).copy()
    upd.run(t)
    assert np.array_equal(first, new.download(idx[(10, -1, -1)]))
