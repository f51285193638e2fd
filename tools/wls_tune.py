"""Dev tool: WLS MG-PCG iteration counts / stage time on one 700^2 pair for the current NCT_MG_* environment."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
from oracle import synth
pkg = g.load_package(); dev = torch.device("cuda:0"); ctx = pkg.Context(0)
ctx.load_vgg19_weights(synth.vgg19_weights(19))
c, s = synth.pair(0, 700, 700)
tc, ts = torch.from_numpy(c).to(dev), torch.from_numpy(s).to(dev)
tol = float(os.environ.get("WLS_TOL", "1e-10"))
cfg = ctx.default_config(wls_rel_tol=tol)
ctx.transfer_pair_dev(tc, ts, cfg); ctx.synchronize()
ctx.profile(True)
ctx.transfer_pair_dev(tc, ts, cfg)
rep = ctx.profile_report()
print(json.dumps(dict(alpha=os.environ.get("NCT_MG_ALPHA"), omega=os.environ.get("NCT_MG_OMEGA"), tol=tol, wls_ms=round(rep["wls"][0], 1))))
