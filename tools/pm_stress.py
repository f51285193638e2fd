"""Dev tool: repeats single-direction PatchMatch runs (BASELINE config 5 geometry) and checks run-to-run identity."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
from oracle import synth
pkg = g.load_package(); dev = torch.device("cuda:0"); ctx = pkg.Context(0)
Cn, H, W = 256, 128, 128
a, b = synth.pm_sweep_volumes(Cn, H, W)
ta, tb = ctx.norm(torch.from_numpy(a).to(dev)), ctx.norm(torch.from_numpy(b).to(dev))
ctx.synchronize()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
bad = 0
for iters in (7, 3, 10):
    ref = None
    for r in range(reps):
        ann = torch.empty(H * W, dtype=torch.int32, device=dev)
        annd = torch.zeros(H * W, dtype=torch.float32, device=dev)
        ctx.init_ann(ann, H, W, H, W)
        ctx.count_evals(bool(r & 1))
        ctx.patchmatch_single(ta, tb, ann, annd, pkg.make_params(Cn, H, W, H, W, iters=iters, rs_max=8))
        ctx.synchronize()
        out = ann.cpu().numpy().copy()
        if ref is None:
            ref = out
            ident = np.array_equal(out.view(np.uint32), ((np.arange(H * W) // W) << 12 | (np.arange(H * W) % W)).astype(np.uint32))
            print(f"iters {iters}: first run identity={ident}")
        elif not np.array_equal(out, ref):
            bad += 1
            ident = np.array_equal(out.view(np.uint32), ((np.arange(H * W) // W) << 12 | (np.arange(H * W) % W)).astype(np.uint32))
            print(f"iters {iters} rep {r}: {(out != ref).sum()} entries differ from the first run; identity={ident}")
print("mismatching runs:", bad)
