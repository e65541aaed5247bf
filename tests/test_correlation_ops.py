"""G / Y / Z / K / M helper tensors of the two-orbital correlation functions: the chain of updates
b2_update_create on correlation operator sets against DMRG::update_correlations_tensors (DMRGoperators3RDM.cpp:415-479) of the
reference, i.e. TensorGYZ::construct / TensorKM::construct (TensorGYZ.cpp:42-98, TensorKM.cpp:42-95) for the newest site and
TensorOperator::update without Jordan-Wigner phase (TensorOperator.cpp:163-405) for the older ones, at EVERY boundary of the chain."""
import numpy as np
import pytest

import cpu_check
from chemps2_b200 import api, fixtures


def _expected(golden, b):
    nb, mr, ops = fixtures.split_ops(golden, f"corr/b{b}")
    assert nb == b and mr
    return ops


def _check(new, arena_of, expected):
    sl = {(k, i, j): (off, size) for k, i, j, off, size in cpu_check.op_slices(new)}
    assert len(sl) == len(expected)
    kinds = set()
    for kind, i, j, data in expected:
        off, size = sl[(kind, i, j)]
        assert size == data.size
        if size == 0:
            continue
        got = arena_of(kind, i, j, off, size)
        assert np.abs(got - data).max() <= 1e-12 * max(1.0, np.abs(data).max()), (api.KIND_NAMES[kind], i, j)
        kinds.add(api.KIND_NAMES[kind])
    return kinds


def test_correlation_tensors_chain_cpu(golden):
    """CPU: compiled work lists through the emulator in oracle/; the old set of every step is the REFERENCE's previous table"""
    ctx = api.context_from_fixture(golden, "corr")
    seen = set()
    for b in range(1, ctx.L):
        old = None
        if b > 1:
            old = api.OpSet(ctx, b - 1, True, correlation=True)
            old.upload_all(_expected(golden, b - 1))
        new = api.OpSet(ctx, b, True, correlation=True)
        upd = api.Update(ctx, b - 1, True, old, new)
        arena = cpu_check.emulate_update(old, new, upd, golden[f"corr/mps/{b - 1}"])
        seen |= _check(new, lambda k, i, j, off, size: arena[off:off + size], _expected(golden, b))
    assert seen == {"G", "Y", "Z", "K", "M"}


@pytest.mark.gpu
def test_correlation_tensors_chain_gpu(golden):
    """GPU through the C ABI, chained on OUR OWN previous tables (errors would accumulate along the chain)"""
    ctx = api.context_from_fixture(golden, "corr", device=0)
    old = None
    for b in range(1, ctx.L):
        new = api.OpSet(ctx, b, True, correlation=True)
        api.Update(ctx, b - 1, True, old, new).run(golden[f"corr/mps/{b - 1}"])
        idx = {}
        for n in range(len(new)):
            k, i, j, _ = new.info(n)
            idx[(k, i, j)] = n
        _check(new, lambda k, i, j, off, size: new.download(idx[(k, i, j)]), _expected(golden, b))
        old = new
