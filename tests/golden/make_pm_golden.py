"""Generates tests/golden/pm_golden.npz from the oracle (run once; the file is committed so a
later change to oracle/pm_oracle.c that alters results is caught)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402
from oracle import synth  # noqa: E402

out = {}
cases = [(64, 20, 24, 22, 19, 3, 6), (128, 16, 16, 16, 16, 10, 4), (256, 12, 14, 13, 12, 2, 32), (512, 11, 11, 11, 11, 4, 2),
         (32, 10, 12, 12, 10, 2, 4), (16, 10, 10, 10, 10, 2, 4)]
for i, (Cn, ah, aw, bh, bw, iters, rs) in enumerate(cases):
    a = oracle.l2norm_hwc(synth.feature_volume(11, ah, aw, Cn))
    b = oracle.l2norm_hwc(synth.feature_volume(12, bh, bw, Cn))
    ann, annd, st = oracle.patchmatch(a, b, oracle.nnf_init(ah, aw, bh, bw), oracle.make_params(Cn, ah, aw, bh, bw, iters=iters, rs_max=rs))
    out[f"case{i}_cfg"] = np.array([Cn, ah, aw, bh, bw, iters, rs], np.int32)
    out[f"case{i}_ann"] = ann
    out[f"case{i}_annd"] = annd
out["xorwow_seed0"] = oracle.xorwow_raw(0, 8)
out["xorwow_seed699"] = oracle.xorwow_raw(699, 8)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pm_golden.npz"), **out)
print("written", len(out), "arrays")
