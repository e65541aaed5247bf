"""Multi-GPU sweep (needs >= 2 GPUs; skipped on the 1-GPU box): two ranks over NCCL share one DMRG calculation — sigma terms
sharded by the reference's ownership maps, operator updates sharded by operator, MPS / Davidson / Split replicated — and must
reproduce the single-GPU sweep energies (which are pinned to the reference in test_dmrg_gpu.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, {root!r})
from chemps2_b200 import api, fixtures
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
fx = fixtures.load(os.path.join({root!r}, "tests", "golden", "n2_sto3g_singlet.npz"))
L, group, N, twoS, irrep = [int(x) for x in fx["problem/hdr"]]

def run(world_, rank_, ar, own_stream=False):
    ctx = api.Context(local)
    if not own_stream:   # own_stream: the library keeps the stream it created; the all-reduce callback must run ON that stream
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_problem(L, group, N, twoS, irrep, fx["problem/orb_irrep"], mx=fx["problem/mx"], econst=float(fx["problem/econst"][0]))
    D = 64
    ctx.bk_init(D)
    d = api.DMRG(ctx)
    if world_ > 1:
        d.set_world(world_, rank_, ar)
    d.random_mps(77)
    for i in range(L - 2):
        d.update(i, True)
    out = []
    change = False
    for it in range(3):
        el, dl = d.sweep(False, 1e-8, 0.0, D, change)
        change = True
        er, dr = d.sweep(True, 1e-8, 0.0, D, change)
        out += [el, er, dl, dr]
    return np.array(out)

ar = api.AllReduce()
multi = run(world, rank, ar)
multi_own = run(world, rank, ar, own_stream=True)
single = run(1, 0, None)
err = max(float(np.abs(multi - single).max()), float(np.abs(multi_own - single).max()))
t = torch.tensor([err], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("B2MG max_diff", float(t.item()), "energy", multi[-3], "allreduce_calls", ar.calls)
dist.destroy_process_group()
"""


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_two_gpu_sweep_matches_single_gpu(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29631", str(script)], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("B2MG")][-1].split()
    assert float(line[2]) < 1e-9, line     # energies and discarded weights agree with the single-GPU sweep
    assert int(line[-1]) > 10              # the all-reduce path really ran
