"""Dev tool for ncu: runs N pairs of the full pipeline at side x side (no timing)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
from oracle import synth
side = int(sys.argv[1]) if len(sys.argv) > 1 else 700
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pkg = g.load_package(); dev = torch.device("cuda:0"); ctx = pkg.Context(0)
ctx.load_vgg19_weights(synth.vgg19_weights(19))
ctx.set_vgg_engine(int(os.environ.get("NCT_ENGINE", "3")))
c, s = synth.pair(0, side, side)
tc, ts = torch.from_numpy(c).to(dev), torch.from_numpy(s).to(dev)
for i in range(n):
    out = ctx.transfer_pair_dev(tc, ts)
ctx.synchronize()
print("done", int(out.sum()))
