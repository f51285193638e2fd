"""GPU parity tests of k-means labels and the in-cluster 8-NN search (bit-exact ids, weights to 1 ulp of exp)."""
import numpy as np
import pytest

import oracle
from oracle import color, synth

pytestmark = pytest.mark.gpu


def to_dev(x, dev):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    torch.cuda.synchronize()  # libnct contexts run on their own non-blocking stream: the copy must have landed
    return t


@pytest.mark.parametrize("h,w,Cn,seed", [(44, 44, 512, 1), (63, 63, 512, 2), (16, 16, 512, 3), (33, 22, 512, 4), (20, 20, 64, 5)])
def test_kmeans_labels_bit_exact(ctx, dev, h, w, Cn, seed):
    f = oracle.l2norm_hwc(synth.feature_volume(seed, h, w, Cn, smooth=6))
    g = ctx.cluster_features(to_dev(f, dev), 10, 11)
    o, nl = oracle.kmeans_labels(f.reshape(h * w, Cn), 10, 11)
    assert nl == 10
    assert np.array_equal(g.cpu().numpy(), o)


def test_kmeans_degenerate_inputs(ctx, dev):
    f = oracle.l2norm_hwc(synth.feature_volume(1, 3, 2, 64))
    assert np.all(ctx.cluster_features(to_dev(f, dev)).cpu().numpy() == 0)           # fewer than 10 points
    dup = np.repeat(f.reshape(6, 64)[:3], 12, axis=0).reshape(6, 6, 64)
    assert np.all(ctx.cluster_features(to_dev(dup, dev)).cpu().numpy() == 0)          # fewer than 10 distinct points


@pytest.mark.parametrize("brute", [False, True])
@pytest.mark.parametrize("lw,lh,samples,h,w", [(6, 6, 1, 6, 6), (6, 6, 2, 12, 11), (11, 13, 4, 50, 43), (44, 44, 2, 88, 88), (44, 44, 4, 175, 175)])
def test_find_knns_bit_exact(ctx, dev, lw, lh, samples, h, w, brute):
    rng = np.random.default_rng(lw + h)
    # blobby label map, like a k-means segmentation
    base = rng.integers(0, 10, ((lh + 3) // 4, (lw + 3) // 4))
    labels = np.kron(base, np.ones((4, 4), np.int64))[:lh, :lw].astype(np.int32).ravel()
    cnt, _ = synth.pair(6, h, w)
    lab = color.bgr2lab_u8(cnt)
    gi, gw = ctx.find_knns(to_dev(labels, dev), lw, lh, to_dev(lab, dev), samples, brute=brute)
    ctx.synchronize()
    oi, ow = oracle.find_knns(labels, lw, lh, lab, samples)
    assert np.array_equal(gi.cpu().numpy(), oi)
    assert np.array_equal(gw.cpu().numpy(), ow)  # exp() from a host-libm table: bit-exact


def test_find_knns_full_size_level(ctx, dev):
    """finest level of a 700^2 pair: 490k queries, 44x44 label grid, 16x16-pixel cells"""
    rng = np.random.default_rng(0)
    base = rng.integers(0, 10, (11, 11))
    labels = np.kron(base, np.ones((4, 4), np.int64)).astype(np.int32).ravel()
    cnt, _ = synth.pair(7, 700, 700)
    lab = color.bgr2lab_u8(cnt)
    gi, gw = ctx.find_knns(to_dev(labels, dev), 44, 44, to_dev(lab, dev), 16)
    ctx.synchronize()
    oi, ow = oracle.find_knns(labels, 44, 44, lab, 16)
    assert np.array_equal(gi.cpu().numpy(), oi)
    assert np.array_equal(gw.cpu().numpy(), ow)  # exp() from a host-libm table: bit-exact


def test_find_knns_sparse_colours_and_tiny_clusters(ctx, dev):
    """random (sparse, far apart) colours force the shell search to its whole-cluster fallback; one label covers a
    single cell (tiny cluster path); both must still equal the oracle."""
    rng = np.random.default_rng(5)
    lw = lh = 8
    labels = rng.integers(0, 3, lw * lh).astype(np.int32)
    labels[0] = 9
    lab = rng.integers(0, 256, (32, 32, 3), dtype=np.uint8)
    gi, gw = ctx.find_knns(to_dev(labels, dev), lw, lh, to_dev(lab, dev), 4)
    ctx.synchronize()
    oi, ow = oracle.find_knns(labels, lw, lh, lab, 4)
    assert np.array_equal(gi.cpu().numpy(), oi)
    assert np.array_equal(gw.cpu().numpy(), ow)  # exp() from a host-libm table: bit-exact
