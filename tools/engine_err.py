"""Dev tool: per-level error of the tensor-core conv engines against the FP32 CUDA-core engine."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
from oracle import synth
pkg = g.load_package(); dev = torch.device("cuda:0")
w = synth.vgg19_weights(19)
img = torch.from_numpy(synth.pair(7, 256, 256)[0]).to(dev)
feats = {}
for eng in (0, 1, 2):
    c = pkg.Context(0); c.load_vgg19_weights(w); c.set_vgg_engine(eng)
    f = c.predict(img, 0); c.synchronize(); feats[eng] = [t.cpu().numpy() for t in f]; c.close()
for eng in (1, 2):
    for l in (4, 3, 2, 1, 0):
        r, x = feats[0][l].astype(np.float64), feats[eng][l].astype(np.float64)
        d = x - r
        print(json.dumps(dict(engine=eng, level=l, max_of_range=float(np.abs(d).max() / np.abs(r).max()),
                              rms_rel=float(np.sqrt((d ** 2).mean()) / np.sqrt((r ** 2).mean())),
                              mean_signed_rel=float(d.mean() / np.abs(r).mean()))))
