"""CPU tests of the BDS-vote oracle (oracle/bds_oracle.c)."""
import numpy as np
import pytest

import oracle
from oracle import synth


def identity_nnf(h, w):
    gx, gy = np.meshgrid(np.arange(w), np.arange(h))
    return ((gy.ravel().astype(np.uint32) << 12) | gx.ravel().astype(np.uint32))


def test_reconstruct_identity_returns_b():
    """With identity NNFs in both directions every vote for a pixel is that pixel of B."""
    _, b = synth.pair(0, 24, 20)
    ann = identity_nnf(24, 20)
    out = oracle.reconstruct_bds(b, b, ann, ann, 1.0, 2.0)
    # (v*na*wa + v*nb*wb)/(na*wa + nb*wb) can land one ulp below v before truncation (decision B4)
    assert np.all((out.astype(int) - b.astype(int) <= 0) & (b.astype(int) - out.astype(int) <= 1))


def test_reconstruct_zero_completeness_weight_is_coherence_average():
    cnt, stl = synth.pair(1, 18, 22, 20, 19)
    rng = np.random.default_rng(3)
    ann = ((rng.integers(0, 20, 18 * 22).astype(np.uint32) << 12) | rng.integers(0, 19, 18 * 22).astype(np.uint32))
    bnn = ((rng.integers(0, 18, 20 * 19).astype(np.uint32) << 12) | rng.integers(0, 22, 20 * 19).astype(np.uint32))
    out = oracle.reconstruct_bds(cnt, stl, ann, bnn, 1.0, 0.0)
    # brute-force coherence average
    x, y = oracle.unpack(ann)
    for (ax, ay) in [(0, 0), (5, 7), (21, 17), (10, 0)]:
        s = np.zeros(3, np.int64)
        n = 0
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                if 0 <= ax + dx < 22 and 0 <= ay + dy < 18:
                    q = (ay + dy) * 22 + ax + dx
                    xp, yp = x[q] - dx, y[q] - dy
                    if 0 <= xp < 19 and 0 <= yp < 20:
                        s += stl[yp, xp]
                        n += 1
        wa = 1.0 / (22 * 18)
        expect = ((s * wa) / (n * wa)).astype(np.uint8)
        assert np.array_equal(out[ay, ax], expect)


@pytest.mark.parametrize("Cn", [64, 128, 512])
def test_feature_error_modes_agree(Cn):
    """canonical (mode 0) vs reference-order (mode 1) reductions differ by FP32 rounding only (decision B3)."""
    ah, aw, bh, bw = 14, 17, 16, 13
    c = oracle.l2norm_hwc(synth.feature_volume(1, ah, aw, Cn))
    s = synth.feature_volume(2, bh, bw, Cn) * np.float32(7.0)
    rng = np.random.default_rng(5)
    ann = ((rng.integers(0, bh, ah * aw).astype(np.uint32) << 12) | rng.integers(0, bw, ah * aw).astype(np.uint32))
    bnn = ((rng.integers(0, ah, bh * bw).astype(np.uint32) << 12) | rng.integers(0, aw, bh * bw).astype(np.uint32))
    e0, v0 = oracle.bds_feature_error(c, s, ann, bnn, 1.0, 2.0, mode=0, want_vote=True)
    e1, v1 = oracle.bds_feature_error(c, s, ann, bnn, 1.0, 2.0, mode=1, want_vote=True)
    assert np.array_equal(v0, v1)
    assert np.abs(e0 - e1).max() < 1e-5
    assert e0.min() >= -1.0 - 1e-5 and e0.max() <= 1e-5  # -cosine of two non-negative vectors


def test_feature_error_identity_is_self_similarity():
    """Identity NNFs and S == C: the vote at an interior pixel is C's own feature, err = -1."""
    Cn, h, w = 64, 12, 12
    raw = synth.feature_volume(1, h, w, Cn)
    c = oracle.l2norm_hwc(raw)
    ann = identity_nnf(h, w)
    e = oracle.bds_feature_error(c, raw, ann, ann, 1.0, 2.0).reshape(h, w)
    assert np.allclose(e, -1.0, atol=1e-5)
