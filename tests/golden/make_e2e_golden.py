"""Generates tests/golden/e2e_golden.npz: the whole pipeline on two small synthetic pairs through the canonical oracle
(oracle/conv_oracle.c features, deterministic PatchMatch, canonical-order CG, direct WLS solve).  The file is committed:
a later change to any oracle part that alters the result is caught on the CPU (tests/test_oracle_pipeline.py), and the
GPU product is compared with the same bytes (tests/test_gpu_pipeline.py) without running the oracle.

Run once from the repo root:  python tests/golden/make_e2e_golden.py"""
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pipeline, synth, vgg  # noqa: E402

CASES = [(31, 96, 96, 96, 96), (32, 88, 112, 104, 80)]  # (pair seed, content h, w, style h, w)


def run(seed, ch, cw, sh, sw, weights):
    cnt, stl = synth.pair(seed, ch, cw, sh, sw)
    levels = {}

    def on_level(l, d):
        levels[l] = dict(ann=zlib.crc32(np.ascontiguousarray(d["ann"]).tobytes()), bnn=zlib.crc32(np.ascontiguousarray(d["bnn"]).tobytes()),
                         sml=zlib.crc32(np.ascontiguousarray(d["sml"]).tobytes()), knn=zlib.crc32(np.ascontiguousarray(d["knn_id"]).tobytes()),
                         img=zlib.crc32(np.ascontiguousarray(d["result"]).tobytes()), cg_iters=list(d["cg_iters"]))

    out = pipeline.transfer_pair(cnt, stl, None, features_fn=lambda img, deepest: vgg.features_canonical(img, weights, deepest),
                                 cg_mode="canonical", on_level=on_level)
    return out, levels


if __name__ == "__main__":
    w = synth.vgg19_weights(19)
    data = {"cases": np.array(CASES, np.int32)}
    for i, c in enumerate(CASES):
        out, levels = run(*c, w)
        data[f"case{i}_out"] = out
        data[f"case{i}_crc"] = np.array([[levels[l][k] for k in ("ann", "bnn", "sml", "knn", "img")] for l in range(5)], np.uint32)
        data[f"case{i}_cg_iters"] = np.array([levels[l]["cg_iters"] for l in range(5)], np.int32)
        print(f"case {i}: {c} -> image crc {zlib.crc32(out.tobytes()):08x}")
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "e2e_golden.npz"), **data)
