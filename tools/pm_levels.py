"""Dev tool: time nct_patchmatch_bidir at the five level shapes of a 700^2 pair (CUDA events on the
launching stream) and print achieved algorithmic GB/s (evaluated candidates x 9 x C x 4 B)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402
from oracle import synth  # noqa: E402

pkg = g.load_package()
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
ctx = pkg.Context(0, stream)
side = int(sys.argv[1]) if len(sys.argv) > 1 else 700
sizes = pkg.level_sizes(side)[::-1]
chans = [512, 512, 256, 128, 64]
ranges = [side // 16, side // 32, side // 64, 32, 32]
res = []
for lvl, (n, Cn, rs) in enumerate(zip(sizes, chans, ranges)):
    a = torch.from_numpy(synth.feature_volume(21 + lvl, n, n, Cn, smooth=max(2, n // 22))).to(dev)
    b = torch.from_numpy(synth.feature_volume(31 + lvl, n, n, Cn, smooth=max(2, n // 22))).to(dev)
    with torch.cuda.stream(stream):
        na, nb = ctx.norm(a), ctx.norm(b)
        ann = torch.empty(n * n, dtype=torch.int32, device=dev)
        bnn = torch.empty(n * n, dtype=torch.int32, device=dev)
        annd = torch.empty(n * n, dtype=torch.float32, device=dev)
        bnnd = torch.empty(n * n, dtype=torch.float32, device=dev)
        p = pkg.make_params(Cn, n, n, n, n, iters=10, rs_max=rs)
        ctx.count_evals(True)
        ctx.init_ann(ann, n, n, n, n); ctx.init_ann(bnn, n, n, n, n)
        ctx.patchmatch_bidir(na, nb, ann, annd, bnn, bnnd, p)
        ev, ev_ref = ctx.patchmatch_stats()
        ctx.count_evals(False)
        times = []
        for r in range(4):
            ctx.init_ann(ann, n, n, n, n); ctx.init_ann(bnn, n, n, n, n)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.patchmatch_bidir(na, nb, ann, annd, bnn, bnnd, p)
            e1.record(stream)
            stream.synchronize()
            times.append(e0.elapsed_time(e1))
    ms = min(times[1:])
    gb = ev * 9 * Cn * 4 / 1e9
    res.append(dict(level=lvl, n=n, C=Cn, ms=round(ms, 3), evals=ev, evals_ref=ev_ref, alg_GB=round(gb, 2),
                    GBps=round(gb / ms * 1e3, 1), mean_annd=float(annd.mean())))
    print(json.dumps(res[-1]), flush=True)
print(json.dumps(dict(total_ms=round(sum(r["ms"] for r in res), 2), total_GB=round(sum(r["alg_GB"] for r in res), 1))))
