"""Dev tool: time the VGG-19 trunk forward (full and truncated) at 700^2 with CUDA events."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
from oracle import synth
pkg = g.load_package(); dev = torch.device("cuda:0"); stream = torch.cuda.Stream(); ctx = pkg.Context(0, stream)
ctx.load_vgg19_weights(synth.vgg19_weights(19))
ctx.set_vgg_engine(int(os.environ.get("NCT_VGG_ENGINE", "0")))
side = int(sys.argv[1]) if len(sys.argv) > 1 else 700
img = torch.from_numpy(synth.pair(0, side, side)[0]).to(dev)
GF = {0: 355.5, 1: 346.3, 2: 110.1 + 0, 3: 55.9, 4: 1.7}
with torch.cuda.stream(stream):
    for deepest in (0, 1, 2, 3, 4):
        feats = ctx.predict(img, deepest)
        ts = []
        for r in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); ctx.predict(img, deepest, out=feats); e1.record(stream); stream.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(json.dumps(dict(deepest_level=deepest, ms=round(min(ts[1:]), 3))), flush=True)
