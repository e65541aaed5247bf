"""N > 1 path on the CPU: two processes (gloo), each holding the owner share of the sigma terms the reference's ownership maps
(MPIchemps2.h:158-231, mpi_size -> world) give to its rank; the partial sigma vectors are summed with an all-reduce exactly like
bench.py does with NCCL.  The per-rank arithmetic runs through the work-list emulator in oracle/ (test infrastructure)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import cpu_check
from chemps2_b200 import fixtures
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
fx = fixtures.load(os.path.join({root!r}, "tests", "golden", "h2o_631g.npz"))
worst = 0.0
for tag in ("A", "B"):
    ctx, left, right, heff = cpu_check.build_case(fx, tag, world=world, rank=rank)
    part = cpu_check.emulate_worklists(ctx, left, right, heff, fx[tag + "/rnd_in"])
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t)
    ref = fx[tag + "/rnd_out"]
    worst = max(worst, float(np.abs(t.numpy() - ref).max() / max(1.0, np.abs(ref).max())))
    frac = heff.stats()["flops_exec"]
    fr = torch.tensor([frac], dtype=torch.float64)
    dist.all_reduce(fr)
    if rank == 0:
        print("B2DIST", tag, "share_of_rank0", frac / float(fr.item()))
# operator update sharded by operator (b2_update_create_sharded): pass 0 per rank, all-reduce of the arenas, mixing pass in full
for which in ("UR", "UL"):
    ctx, old, new, upd, tt, expected = cpu_check.build_update_case(fx, which, world=world, rank=rank)
    part = torch.from_numpy(cpu_check.emulate_update(old, new, upd, tt, passes=(0,)))
    dist.all_reduce(part)
    arena = cpu_check.emulate_update(old, new, upd, tt, passes=(1,), arena=part.numpy())
    sl = {{(k, i, j): (off, size) for k, i, j, off, size in cpu_check.op_slices(new)}}
    for kind, i, j, data in expected:
        off, size = sl[(kind, i, j)]
        if size:
            worst = max(worst, float(np.abs(arena[off:off + size] - data).max() / max(1.0, np.abs(data).max())))
if rank == 0:
    print("B2DIST worst", worst)
dist.destroy_process_group()
"""


def test_owner_sharded_sigma_allreduce_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29617", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("B2DIST")]
    worst = float([ln for ln in lines if "worst" in ln][0].split()[-1])
    assert worst < 1e-12
    shares = [float(ln.split()[-1]) for ln in lines if "share_of_rank0" in ln]
    assert all(0.05 < s < 0.95 for s in shares), shares   # both ranks own real work
