#!/usr/bin/env python
"""GPU diagnostic of the exact fixed-point conv engine (conv_i8.cu) against oracle/vgg.py::q_conv3x3_relu: raw INT32
accumulators and outputs, per tile configuration (NCT_I8_BN / NCT_I8_KB).  Run on a B200: `python tools/q_diag.py`."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import __graft_entry__ as g
from oracle import vgg

pkg = g.load_package()
ctx = pkg.Context(0)
dev = torch.device("cuda:0")
rng = np.random.default_rng(3)


def case(H, W, cin, cout, bn, kb, probe=False):
    os.environ["NCT_I8_BN"], os.environ["NCT_I8_KB"] = str(bn), str(kb)
    if probe:  # channel-pairing probe: x[p, c] = c + 1, w[o, c, centre] = (c == o % cin)
        x = np.tile(np.arange(1, cin + 1, dtype=np.float32), (H, W, 1))
        w = np.zeros((cout, cin, 3, 3), np.float32)
        for o in range(cout):
            w[o, o % cin, 1, 1] = 1.0
        b = np.zeros(cout, np.float32)
    else:
        x = np.abs(rng.standard_normal((H, W, cin))).astype(np.float32) * rng.uniform(0.2, 3.0)
        x[rng.random((H, W, cin)) < 0.3] = 0
        w = (rng.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / (9 * cin))).astype(np.float32)
        b = (0.1 * rng.standard_normal(cout)).astype(np.float32)
    ref, racc = vgg.q_conv3x3_relu(x, w, b, return_acc=True)
    xt = torch.from_numpy(x).to(dev)
    torch.cuda.synchronize()
    t0 = time.time()
    out, acc = ctx.conv3x3_fixedpoint(xt, w, b, debug_acc=True)
    ctx.synchronize()
    out, acc = out.cpu().numpy(), acc.cpu().numpy().reshape(4, H, W, cout)
    msg = f"H{H} W{W} cin{cin} cout{cout} BN{bn} KB{kb}{' probe' if probe else ''}:"
    ok = True
    for d in range(4):
        bad = acc[d].astype(np.int64) != racc[d]
        msg += f" acc{d} {int(bad.sum())}/{bad.size}"
        if bad.any():
            ok = False
            idx = np.argwhere(bad)[:4]
            msg += " e.g. " + "; ".join(f"(y{y},x{x_},o{o}) gpu {acc[d][y, x_, o]} ref {racc[d][y, x_, o]}" for y, x_, o in idx)
    nbad = int((out.view(np.uint32) != ref.view(np.uint32)).sum())
    msg += f" | out {nbad}/{out.size} differ, max abs {np.abs(out - ref).max():.3e}"
    print(("OK   " if ok and nbad == 0 else "FAIL ") + msg, flush=True)
    if probe and not ok:
        print("   pairing row o=0..15 of acc0/32 at centre pixel:", (acc[0][H // 2, W // 2, :16] // 32).tolist(), flush=True)
    return ok and nbad == 0


allok = True
for bn, kb in [(128, 128), (64, 128), (128, 64), (64, 64)]:
    cin = 128
    r = case(24, 40, cin, 128, bn, kb)
    if not r:
        case(16, 16, cin, 128, bn, kb, probe=True)
    allok &= r
r = case(37, 53, 64, 64, 64, 64)      # conv1_2 geometry: Cin = 64 forces KB = 64
if not r:
    case(16, 16, 64, 64, 64, 64, probe=True)
allok &= r
allok &= case(19, 21, 64, 128, 128, 64)
allok &= case(45, 29, 256, 256, 128, 128)
allok &= case(11, 13, 512, 512, 128, 128)
allok &= case(11, 13, 512, 512, 128, 64)
print("ALL OK" if allok else "SOME FAILED")
# timing of the conv1_2 / conv3_2 / conv4_2 geometries at 700^2 through the full trunk is in bench.py; quick layer timing here
for (H, W, cin, cout) in [(700, 700, 64, 64), (350, 350, 128, 128), (175, 175, 256, 256), (88, 88, 512, 512)]:
    for bn, kb in [(128, 128), (64, 128), (128, 64), (64, 64)]:
        if cin == 64 and kb == 128 or cout == 64 and bn == 128:
            continue
        os.environ["NCT_I8_BN"], os.environ["NCT_I8_KB"] = str(bn), str(kb)
        x = torch.rand((H, W, cin), device=dev)
        w = (rng.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / (9 * cin))).astype(np.float32)
        b = np.zeros(cout, np.float32)
        torch.cuda.synchronize()
        ctx.conv3x3_fixedpoint(x, w, b)
        prof0 = ctx.launch_count
        ctx.profile(True) if hasattr(ctx, "profile") else None
        t0 = time.time()
        for _ in range(3):
            ctx.conv3x3_fixedpoint(x, w, b)
        ctx.synchronize()
        print(f"layer {H}x{W} {cin}->{cout} BN{bn} KB{kb}: {(time.time() - t0) / 3 * 1e3:.2f} ms wall per call (incl. host weight prep + upload)", flush=True)
ctx.close()
