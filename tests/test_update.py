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


@pytest.mark.parametrize("options", [{}, {"work_budget": 4096, "chunk_k": 32}, {"parallel_min_terms": 16}, {"parallel_min_terms": 16, "work_budget": 8192, "chunk_k": 32}],
                         ids=["default", "tiny-waves", "parallel-plan", "parallel-plan-tiny-waves"])
@pytest.mark.parametrize("which", ["UR", "UL"])
def test_update_worklists_vs_reference_cpu(golden, which, options):
    """CPU: the compiled update work lists executed by the emulator in oracle/ (no GPU needed to pin the plan)"""
    ctx, old, new, upd, t, expected = cpu_check.build_update_case(golden, which, options=options)
    arena = cpu_check.emulate_update(old, new, upd, t)
    seen = _compare(lambda k, i, j, off, size: arena[off:off + size], new, expected)
    assert {"S0", "F0", "F1", "C", "X"} <= seen
    st = upd.stats()
    assert st["terms"] > 0 and st["flops_ref"] > 0


@pytest.mark.parametrize("world", [2, 5])
@pytest.mark.parametrize("which", ["UR", "UL"])
def test_sharded_update_sums_to_reference_cpu(golden, which, world):
    """multi-GPU update on the CPU emulator: every rank runs pass 0 for the operators assigned to it, the arenas are summed
    (the all-reduce), then the mixing pass runs in full -> the reference's operators"""
    total, flops = None, []
    for r in range(world):
        ctx, old, new, upd, t, expected = cpu_check.build_update_case(golden, which, world=world, rank=r)
        part = cpu_check.emulate_update(old, new, upd, t, passes=(0,))
        total = part if total is None else total + part
        flops.append(upd.stats()["terms"])
    arena = cpu_check.emulate_update(old, new, upd, t, passes=(1,), arena=total)
    _compare(lambda k, i, j, off, size: arena[off:off + size], new, expected)
    ctx1, old1, new1, upd1, _, _ = cpu_check.build_update_case(golden, which)
    assert sum(flops) == upd1.stats()["terms"]           # every term computed exactly once
    if world == 2:
        assert min(flops) > 0.2 * max(flops)                # both ranks own real work


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["UR", "UL"])
def test_sharded_update_gpu_single_process(golden, which):
    """GPU, world = 2 emulated in one process: rank 1 runs first and its all-reduce callback snapshots its pass-0 arena; rank 0's
    callback adds that snapshot -> rank 0 ends with the reference's operators (b2_update_create_sharded / _set_allreduce)"""
    import ctypes as C
    import torch
    snap = {}

    def make(cb):
        ar = api.AllReduce.__new__(api.AllReduce)
        ar.cfn = api.ALLREDUCE_FN(cb)
        return ar

    def view(ptr, n):
        return torch.as_tensor(api.AllReduce._Dev(ptr, n), device="cuda")

    def cb1(user, ptr, n, stream):
        torch.cuda.synchronize()
        snap["x"] = view(ptr, n).clone()
        return 0

    def cb0(user, ptr, n, stream):
        torch.cuda.synchronize()
        view(ptr, n).add_(snap["x"])
        torch.cuda.synchronize()
        return 0

    ctx, old, new1, upd1, t, expected = cpu_check.build_update_case(golden, which, device=0, world=2, rank=1)
    upd1.set_allreduce(make(cb1))
    upd1.run(t)
    ctx0, old0, new0, upd0, t, expected = cpu_check.build_update_case(golden, which, device=0, world=2, rank=0)
    upd0.set_allreduce(make(cb0))
    upd0.run(t)
    idx = {}
    for n in range(len(new0)):
        k, i, j, _ = new0.info(n)
        idx[(k, i, j)] = n
    _compare(lambda k, i, j, off, size: new0.download(idx[(k, i, j)]), new0, expected)
    # without a callback a sharded update must refuse to run instead of producing partial operators
    ctx2, old2, new2, upd2, t, _ = cpu_check.build_update_case(golden, which, device=0, world=2, rank=0)
    with pytest.raises(Exception):
        upd2.run(t)


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
