"""Synthetic inputs of SURVEY.md section 8(d) -- shared by tests and bench.py.

Pure numpy, no reference code involved; part of the product package so that bench.py never imports oracle/ on the product arm.
"""
from __future__ import annotations

import numpy as np


def _bilinear_up(g: np.ndarray, H: int, W: int) -> np.ndarray:
    """Bilinear resample of a coarse grid g[gh, gw, 3] to H x W (corner-aligned), float64."""
    gh, gw, _ = g.shape
    ys = np.linspace(0.0, gh - 1.0, H)
    xs = np.linspace(0.0, gw - 1.0, W)
    y0 = np.minimum(np.floor(ys).astype(np.int64), gh - 2)
    x0 = np.minimum(np.floor(xs).astype(np.int64), gw - 2)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    g00 = g[y0][:, x0]
    g01 = g[y0][:, x0 + 1]
    g10 = g[y0 + 1][:, x0]
    g11 = g[y0 + 1][:, x0 + 1]
    return (g00 * (1 - fx) + g01 * fx) * (1 - fy) + (g10 * (1 - fx) + g11 * fx) * fy


def image(seed: int, H: int, W: int) -> np.ndarray:
    """img(seed,H,W): 4 octaves of smooth noise, min-max to [0,255], uint8 BGR (H, W, 3)."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((H, W, 3), np.float64)
    for o in range(4):
        gh = -(-H // 2 ** (6 - o)) + 1
        gw = -(-W // 2 ** (6 - o)) + 1
        g = rng.random((gh, gw, 3))
        acc += 2.0 ** (-o) * _bilinear_up(g, H, W)
    lo, hi = acc.min(), acc.max()
    return np.clip(np.rint((acc - lo) / (hi - lo) * 255.0), 0, 255).astype(np.uint8)


def pair(i: int, H: int, W: int, Hs: int | None = None, Ws: int | None = None):
    """(content, style) uint8 BGR for pair index i: content seed 1000+i; style seed 2000+i,
    channels rolled by one and gamma 0.8 (different colour statistics, similar structure)."""
    Hs = Hs or H
    Ws = Ws or W
    cnt = image(1000 + i, H, W)
    stl = image(2000 + i, Hs, Ws)
    stl = np.roll(stl, 1, axis=2)
    stl = np.clip(np.rint(255.0 * (stl / 255.0) ** 0.8), 0, 255).astype(np.uint8)
    return cnt, stl


VGG19_TRUNK = [  # (name, cin, cout, pool_before) up to conv5_1, prototxt order
    ("conv1_1", 3, 64, False), ("conv1_2", 64, 64, False),
    ("conv2_1", 64, 128, True), ("conv2_2", 128, 128, False),
    ("conv3_1", 128, 256, True), ("conv3_2", 256, 256, False), ("conv3_3", 256, 256, False), ("conv3_4", 256, 256, False),
    ("conv4_1", 256, 512, True), ("conv4_2", 512, 512, False), ("conv4_3", 512, 512, False), ("conv4_4", 512, 512, False),
    ("conv5_1", 512, 512, True),
]


def vgg19_weights(seed: int = 19):
    """He-normal N(0, 2/(9*Cin)) OIHW weights, zero bias, FP32, layer order of the prototxt."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, cin, cout, _ in VGG19_TRUNK:
        w = rng.standard_normal((cout, cin, 3, 3)).astype(np.float32) * np.float32(np.sqrt(2.0 / (9.0 * cin)))
        out[name] = (w, np.zeros(cout, np.float32))
    return out


def pm_sweep_volumes(Cn: int = 256, H: int = 128, W: int = 128, shift=(7, -3)):
    """BASELINE config 5: A = |N(0,1)| seed 5; B = A circularly shifted by (+7,-3) + 0.1|N(0,1)| seed 6;
    returned un-normalised, HWC float32 (normalise with the implementation under test)."""
    a = np.abs(np.random.default_rng(5).standard_normal((H, W, Cn))).astype(np.float32)
    n = np.abs(np.random.default_rng(6).standard_normal((H, W, Cn))).astype(np.float32)
    b = np.roll(a, shift=(shift[1], shift[0]), axis=(0, 1)) + np.float32(0.1) * n
    return a, b.astype(np.float32)


def feature_volume(seed: int, H: int, W: int, Cn: int, smooth: int = 4):
    """Post-ReLU-like synthetic feature volume (HWC float32): smooth non-negative noise, so
    neighbouring pixels correlate the way conv features do."""
    rng = np.random.default_rng(seed)
    gh, gw = -(-H // smooth) + 1, -(-W // smooth) + 1
    g = rng.standard_normal((gh, gw, Cn))
    ys = np.linspace(0.0, gh - 1.0, H)
    xs = np.linspace(0.0, gw - 1.0, W)
    y0 = np.minimum(np.floor(ys).astype(np.int64), gh - 2)
    x0 = np.minimum(np.floor(xs).astype(np.int64), gw - 2)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    v = (g[y0][:, x0] * (1 - fx) + g[y0][:, x0 + 1] * fx) * (1 - fy) + (g[y0 + 1][:, x0] * (1 - fx) + g[y0 + 1][:, x0 + 1] * fx) * fy
    v = v + 0.25 * rng.standard_normal((H, W, Cn))
    return np.maximum(v, 0.0).astype(np.float32) + np.float32(1e-3)
