"""Pins the oracle's restated distance function against the REFERENCE'S OWN dist_single, compiled verbatim from
/root/reference into oracle/_ref/ by oracle/build_ref.sh (the only part of the hot path that builds from the
reference's sources here)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from oracle import synth

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def load(name):
    p = os.path.join(REF, name)
    if not os.path.exists(p):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    L = C.CDLL(p)
    fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
    L.ref_dist_single.argtypes = [fp, fp] + [C.c_int] * 10 + [C.c_float]
    L.ref_dist_single.restype = C.c_float
    L.ref_xy_to_int.restype = C.c_uint
    return L


@pytest.mark.parametrize("Cn", [3, 64, 512])
def test_restated_distance_equals_verbatim_reference(Cn):
    lib, use_fma = "libref_dist.so", 0
    L = load(lib)
    rng = np.random.default_rng(Cn)
    ah, aw, bh, bw = 7, 9, 8, 6
    a = rng.standard_normal((Cn, ah, aw)).astype(np.float32)
    b = rng.standard_normal((Cn, bh, bw)).astype(np.float32)
    for _ in range(200):
        ax, ay = int(rng.integers(0, aw)), int(rng.integers(0, ah))
        bx, by = int(rng.integers(0, bw)), int(rng.integers(0, bh))
        cutoff = float(rng.choice([2.0 ** 31, 0.0, -1.0, 5.0]))
        ref = L.ref_dist_single(a, b, Cn, ah, aw, bh, bw, ax, ay, bx, by, 3, cutoff)
        mine = oracle.dist_ref_chw(a, b, ax, ay, bx, by, cutoff=cutoff, use_fma=use_fma)
        assert np.float32(ref) == np.float32(mine) or (np.isnan(ref) and np.isnan(mine))


def test_canonical_distance_within_rounding_of_verbatim_reference():
    L = load("libref_dist.so")
    Cn, h, w = 256, 9, 9
    a = oracle.l2norm_hwc(synth.feature_volume(1, h, w, Cn))
    b = oracle.l2norm_hwc(synth.feature_volume(2, h, w, Cn))
    a_chw = np.ascontiguousarray(a.transpose(2, 0, 1))
    b_chw = np.ascontiguousarray(b.transpose(2, 0, 1))
    rng = np.random.default_rng(0)
    for _ in range(100):
        ax, ay, bx, by = (int(v) for v in rng.integers(0, 9, 4))
        ref = L.ref_dist_single(a_chw, b_chw, Cn, h, w, h, w, ax, ay, bx, by, 3, 2.0 ** 31)
        assert abs(ref - oracle.dist_canon(a, b, ax, ay, bx, by)) < 2e-6


def test_packing_matches_reference():
    L = load("libref_dist.so")
    for x, y in [(0, 0), (699, 699), (4095, 4095), (12, 3000)]:
        v = L.ref_xy_to_int(x, y)
        assert v == ((y << 12) | x) and L.ref_int_to_x(v) == x and L.ref_int_to_y(v) == y


@pytest.mark.gpu
def test_restated_distance_equals_reference_device_code(dev):
    """The reference's dist_single compiled by nvcc for sm_100a (default -fmad=true) and run on the GPU equals the
    oracle's reference-order distance with fused multiply-subtract (use_fma = 1): pins decision D2's baseline."""
    import torch

    p = os.path.join(REF, "libref_dist_dev.so")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/libref_dist_dev.so not built")
    L = C.CDLL(p)
    rng = np.random.default_rng(1)
    Cn, ah, aw, bh, bw = 128, 11, 9, 10, 12
    a = rng.standard_normal((Cn, ah, aw)).astype(np.float32)
    b = rng.standard_normal((Cn, bh, bw)).astype(np.float32)
    nq = 300
    q = np.stack([rng.integers(0, aw, nq), rng.integers(0, ah, nq), rng.integers(0, bw, nq), rng.integers(0, bh, nq)], 1).astype(np.int32)
    ta, tb, tq = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev), torch.from_numpy(q).to(dev)
    out = torch.empty(nq, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    assert L.ref_dist_device(C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()), Cn, ah, aw, bh, bw, C.c_void_p(tq.data_ptr()), nq,
                             C.c_void_p(out.data_ptr())) == 0
    got = out.cpu().numpy()
    mine = np.array([oracle.dist_ref_chw(a, b, *map(int, q[i]), use_fma=1) for i in range(nq)], np.float32)
    assert np.array_equal(got.view(np.uint32), mine.view(np.uint32))
