"""Generates tests/golden/fullsize_golden.npz: BASELINE.json's headline configuration through the oracle, once, offline.

  * `e2e700_*`  configs[1]: one 700 x 700 synthetic pair (synth.pair(0, 700, 700), the pair bench.py's first context runs),
    full L = 5 -> 1 pyramid, BDS 2.0.  Oracle = fixed-point features (oracle/vgg.py::features_fixedpoint, the defined
    arithmetic of the product's default tensor-core engine), deterministic PatchMatch, canonical-order CG, DIRECT WLS
    solve (scipy splu).  Stored: the final image and per-level CRC32s of both NNFs, the BDS colour vote, the k-NN ids
    and the level's result image, so a GPU mismatch can be localised to a level and stage.
  * `pm700_*`   the finest PatchMatch level on its own: 700 x 700 x 64, both directions, 10 iterations, rs_max = 32
    (NCT/main.cu:77-83), on synth.feature_volume volumes: CRC32s of NNF and distances + every 97th entry.

Minutes of CPU time; the file is committed so the GPU tests compare against bytes, not against a live oracle run.
Run from the repo root:   python tests/golden/make_fullsize_golden.py [e2e|pm|all]"""
import os
import sys
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402
from oracle import pipeline, synth, vgg  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "fullsize_golden.npz")
SIDE = 700


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def e2e(data):
    w = synth.vgg19_weights(19)
    cnt, stl = synth.pair(0, SIDE, SIDE)
    rows, iters = {}, {}

    def on_level(l, d):
        rows[l] = [crc(d["ann"]), crc(d["bnn"]), crc(d["sml"]), crc(d["knn_id"]), crc(d["result"])]
        iters[l] = list(d["cg_iters"])
        print(f"  level {l} done at {time.time() - t0:.0f} s: {['%08x' % v for v in rows[l]]}", flush=True)

    t0 = time.time()
    out = pipeline.transfer_pair(cnt, stl, None, features_fn=lambda img, deepest: vgg.features_fixedpoint(img, w, deepest),
                                 cg_mode="canonical", on_level=on_level)
    data["e2e700_cfg"] = np.array([0, SIDE, SIDE, SIDE, SIDE], np.int32)
    data["e2e700_out"] = out
    data["e2e700_crc"] = np.array([rows[l] for l in range(5)], np.uint32)
    data["e2e700_cg_iters"] = np.array([iters[l] for l in range(5)], np.int32)
    print(f"e2e 700: image crc {crc(out):08x}, {time.time() - t0:.0f} s", flush=True)


def pm(data):
    Cn, rs, iters = 64, 32, 10
    t0 = time.time()
    a = oracle.l2norm_hwc(synth.feature_volume(21, SIDE, SIDE, Cn))
    b = oracle.l2norm_hwc(synth.feature_volume(22, SIDE, SIDE, Cn))
    p = oracle.make_params(Cn, SIDE, SIDE, SIDE, SIDE, iters=iters, rs_max=rs)
    ann, annd, st = oracle.patchmatch(a, b, oracle.nnf_init(SIDE, SIDE, SIDE, SIDE), p)
    bnn, bnnd, st2 = oracle.patchmatch(b, a, oracle.nnf_init(SIDE, SIDE, SIDE, SIDE), p)
    data["pm700_cfg"] = np.array([Cn, SIDE, SIDE, SIDE, SIDE, iters, rs, 21, 22], np.int32)
    data["pm700_crc"] = np.array([crc(ann), crc(annd), crc(bnn), crc(bnnd)], np.uint32)
    data["pm700_ann_s97"] = ann[::97].copy()
    data["pm700_annd_s97"] = annd[::97].copy()
    data["pm700_bnn_s97"] = bnn[::97].copy()
    data["pm700_evals"] = np.array([int(st[1]), int(st2[1])], np.int64)
    print(f"pm 700x700x64: crc {data['pm700_crc']}, {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    data = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    if what in ("pm", "all"):
        pm(data)
        np.savez_compressed(OUT, **data)
    if what in ("e2e", "all"):
        e2e(data)
        np.savez_compressed(OUT, **data)
    print("written", OUT, sorted(data))
