"""Times only the operator update at the bench shape (for ncu launch lists): python scripts/update_only.py [D]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chemps2_b200 import api, workloads
from chemps2_b200._lib import lib
D = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
w = workloads.get("synth40", D=D)
ctx = w.context(0)
w.apply_distribution(ctx, "gauss")
old = api.OpSet(ctx, w.site, True)
old.fill_hash(7, 1.0)
new = api.OpSet(ctx, w.site + 1, True)
upd = api.Update(ctx, w.site, True, old, new)
t = torch.from_numpy(api.hash_fill(lib.b2_tensor_t_size(ctx.h, w.site), 55) * 0.1).cuda()
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    upd.run_device(t.data_ptr())
    torch.cuda.synchronize(); print("update", rep, time.time() - t0, upd.stats())
