"""Dev tool: (1) how does tcgen05 kind::tf32 convert FP32 operands (truncate or round-to-nearest)?
(2) effect of the TF32 conv engine on the final image of a 700^2 pair vs the FP32 engine."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
from oracle import synth, pipeline
pkg = g.load_package(); dev = torch.device("cuda:0")
# ---- (1) conversion probe: conv1_1 (CUDA cores) emits the constant c on channel 0; conv1_2 and conv2_1 (tensor cores)
# pass channel 0 through with weight 1 -> level-3 feature channel 0 = tf32-converted c
def probe(c, wval=1.0):
    w = {k: (np.zeros_like(v[0]), np.zeros_like(v[1])) for k, v in synth.vgg19_weights(19).items()}
    w["conv1_1"][1][0] = c
    w["conv1_2"][0][0, 0, 1, 1] = 1.0
    w["conv2_1"][0][0, 0, 1, 1] = wval
    ctx = pkg.Context(0); ctx.load_vgg19_weights(w); ctx.set_vgg_engine(1)
    img = torch.zeros((32, 32, 3), dtype=torch.uint8, device=dev)
    # preprocessing subtracts the mean, but conv1_1 weights are zero: output = bias = c
    f = ctx.predict(img, 3); ctx.synchronize()
    v = float(f[3][8, 8, 0]); ctx.close(); return v
one = np.float32(1.0)
for frac in (0.25, 0.5, 0.75):
    c = float(np.float32(1.0 + frac * 2.0 ** -10))
    print(json.dumps(dict(test="operand A", c=c, out=probe(c), trunc=1.0, rn=float(1.0 + (2.0 ** -10 if frac >= 0.5 else 0)))))
for frac in (0.25, 0.75):
    wv = float(np.float32(1.0 + frac * 2.0 ** -10))
    print(json.dumps(dict(test="operand B", w=wv, out=probe(1.0, wv))))
# ---- (2) pipeline: engine 1 vs engine 0
wts = synth.vgg19_weights(19)
outs = []
for eng in (0, 1, 2):
    ctx = pkg.Context(0); ctx.load_vgg19_weights(wts); ctx.set_vgg_engine(eng)
    c, s = synth.pair(0, 700, 700)
    outs.append(ctx.transfer_pair(c, s)); ctx.close()
for k, name in ((1, "tf32"), (2, "3xtf32")):
    print(json.dumps(dict(test=f"pipeline 700^2 {name} vs fp32 engine", psnr=pipeline.psnr(outs[0], outs[k]),
                          mean_abs=float(np.abs(outs[0].astype(int) - outs[k].astype(int)).mean()))))
