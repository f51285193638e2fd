"""CPU tests of the PatchMatch-stage oracle (oracle/pm_oracle.c): self-consistency, the
properties the reference semantics imply, and the committed golden vectors."""
import os

import numpy as np
import pytest

import oracle
from oracle import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_pack_unpack():
    ann = oracle.nnf_init(5, 7, 9, 11)
    x, y = oracle.unpack(ann)
    assert x.max() == 10 and y.max() == 8 and x.min() == 0 and y.min() == 0
    # NCT/GeneralizedPatchMatch.cu:540-541: bx = min(int(ax/(aw-1)*(bw-1)), bw-1)
    for ay in range(5):
        for ax in range(7):
            ex = min(int(np.float32(np.float32(ax) / np.float32(6)) * np.float32(10)), 10)
            ey = min(int(np.float32(np.float32(ay) / np.float32(4)) * np.float32(8)), 8)
            assert x[ay * 7 + ax] == ex and y[ay * 7 + ax] == ey


def test_identity_init_same_size():
    ann = oracle.nnf_init(44, 44, 44, 44)
    x, y = oracle.unpack(ann)
    gx, gy = np.meshgrid(np.arange(44), np.arange(44))
    # float rounding may land one below for a few columns; never above
    assert np.all(x <= gx.ravel()) and np.all(gx.ravel() - x <= 1)
    assert np.all(y <= gy.ravel()) and np.all(gy.ravel() - y <= 1)


def test_upsample_preserves_identity_flow():
    # an identity NNF upsampled 44 -> 88 stays (nearly) the identity: offsets scale by the ratio
    half = oracle.nnf_init(44, 44, 44, 44)
    gx, gy = np.meshgrid(np.arange(44), np.arange(44))
    half = ((gy.ravel().astype(np.uint32) << 12) | gx.ravel().astype(np.uint32))
    up = oracle.nnf_upsample(half, 44, 44, 88, 88, 88, 88)
    x, y = oracle.unpack(up)
    gx, gy = np.meshgrid(np.arange(88), np.arange(88))
    assert np.array_equal(x, gx.ravel()) and np.array_equal(y, gy.ravel())


def test_upsample_odd_sizes_in_bounds():
    rng = np.random.default_rng(0)
    ahh, awh, bhh, bwh = 44, 45, 40, 47
    half = ((rng.integers(0, bhh, ahh * awh).astype(np.uint32) << 12) | rng.integers(0, bwh, ahh * awh).astype(np.uint32))
    up = oracle.nnf_upsample(half, ahh, awh, 88, 89, 79, 93)
    x, y = oracle.unpack(up)
    assert x.min() >= 0 and x.max() <= 92 and y.min() >= 0 and y.max() <= 78


def test_l2norm_unit_and_zero_pixel():
    a = synth.feature_volume(3, 9, 7, 128)
    a[2, 3, :] = 0
    n = oracle.l2norm_hwc(a)
    nn = np.sqrt((n.astype(np.float64) ** 2).sum(-1))
    assert np.all(n[2, 3] == 0)
    nn[2, 3] = 1
    assert np.allclose(nn, 1.0, atol=1e-6)


@pytest.mark.parametrize("Cn", [16, 32, 64, 128, 256, 512])
def test_canonical_dist_close_to_reference_order(Cn):
    """|canonical-order dist - reference-order dist| is FP32 rounding only (decision D2)."""
    H, W = 9, 11
    a = oracle.l2norm_hwc(synth.feature_volume(1, H, W, Cn))
    b = oracle.l2norm_hwc(synth.feature_volume(2, H + 2, W + 1, Cn))
    a_chw = np.ascontiguousarray(a.transpose(2, 0, 1))
    b_chw = np.ascontiguousarray(b.transpose(2, 0, 1))
    rng = np.random.default_rng(1)
    for _ in range(60):
        ax, ay = int(rng.integers(0, W)), int(rng.integers(0, H))
        bx, by = int(rng.integers(0, W + 1)), int(rng.integers(0, H + 2))
        dc = oracle.dist_canon(a, b, ax, ay, bx, by)
        dr = oracle.dist_ref_chw(a_chw, b_chw, ax, ay, bx, by)
        # float64 ground truth
        s, n = 0.0, 0
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if 0 <= ay + dy < H and 0 <= ax + dx < W and 0 <= by + dy < H + 2 and 0 <= bx + dx < W + 1:
                    s -= float(np.dot(a[ay + dy, ax + dx].astype(np.float64), b[by + dy, bx + dx].astype(np.float64)))
                    n += 1
        assert abs(dc - s / n) < 2e-6 and abs(dr - s / n) < 2e-6


def test_reference_cutoff_semantics():
    a = oracle.l2norm_hwc(synth.feature_volume(1, 6, 6, 64))
    a_chw = np.ascontiguousarray(a.transpose(2, 0, 1))
    d = oracle.dist_ref_chw(a_chw, a_chw, 2, 2, 3, 3)
    assert oracle.dist_ref_chw(a_chw, a_chw, 2, 2, 3, 3, cutoff=d - 0.1) == pytest.approx(d - 0.1)
    assert oracle.dist_ref_chw(a_chw, a_chw, 2, 2, 3, 3, cutoff=d + 0.1) == d


def test_self_match_is_minus_one():
    a = oracle.l2norm_hwc(synth.feature_volume(1, 8, 8, 64))
    assert oracle.dist_canon(a, a, 4, 4, 4, 4) == pytest.approx(-1.0, abs=1e-6)
    assert oracle.dist_canon(a, a, 0, 0, 0, 0) == pytest.approx(-1.0, abs=1e-6)  # 4 valid pixels


def test_patchmatch_monotone_and_inbounds():
    H, W, Cn = 40, 36, 64
    a = oracle.l2norm_hwc(synth.feature_volume(1, H, W, Cn))
    b = oracle.l2norm_hwc(synth.feature_volume(2, H - 3, W + 5, Cn))
    init = oracle.nnf_init(H, W, H - 3, W + 5)
    prev = None
    for iters in (0, 1, 2, 5):
        ann, annd, st = oracle.patchmatch(a, b, init, oracle.make_params(Cn, H, W, H - 3, W + 5, iters=iters, rs_max=9))
        x, y = oracle.unpack(ann)
        assert x.max() < W + 5 and y.max() < H - 3
        if prev is not None:
            assert np.all(annd <= prev)  # every entry can only improve
        prev = annd
        # annd is the distance of ann
        for p in (0, 17, H * W - 1):
            assert annd[p] == np.float32(oracle.dist_canon(a, b, p % W, p // W, int(x[p]), int(y[p])))
        assert st[1] <= st[0] <= H * W * (1 + iters * (16 + 4))


def test_patchmatch_recovers_shift():
    """BASELINE config 5 geometry (small): B is A shifted by (+7,-3) plus noise; the NNF finds the shift."""
    a, b = synth.pm_sweep_volumes(64, 48, 48)
    a, b = oracle.l2norm_hwc(a), oracle.l2norm_hwc(b)
    ann, annd, _ = oracle.patchmatch(a, b, oracle.nnf_init(48, 48, 48, 48), oracle.make_params(64, 48, 48, 48, 48, iters=10, rs_max=8))
    x, y = oracle.unpack(ann)
    gx, gy = np.meshgrid(np.arange(48), np.arange(48))
    ok = (x == (gx.ravel() + 7) % 48) & (y == (gy.ravel() - 3) % 48)
    inner = ((gx.ravel() + 7 < 48) & (gy.ravel() - 3 >= 0))
    assert ok[inner].mean() > 0.97


def test_xorwow_uniform_range_and_column_dependence():
    t = oracle.xorwow_uniform_table(8, 100)
    assert t.min() > 0.0 and t.max() <= 1.0
    assert not np.array_equal(t[0], t[1])
    # the same column always yields the same stream (every row of a column shares it, :60-66)
    assert np.array_equal(t[3], oracle.xorwow_uniform_table(4, 100)[3])


def test_golden_vectors():
    """Committed outputs of the oracle on seeded inputs (tests/golden/make_pm_golden.py)."""
    path = os.path.join(GOLD, "pm_golden.npz")
    g = np.load(path)
    for key in [k[:-4] for k in g.files if k.endswith("_ann")]:
        Cn, ah, aw, bh, bw, iters, rs = [int(v) for v in g[key + "_cfg"]]
        a = oracle.l2norm_hwc(synth.feature_volume(11, ah, aw, Cn))
        b = oracle.l2norm_hwc(synth.feature_volume(12, bh, bw, Cn))
        ann, annd, st = oracle.patchmatch(a, b, oracle.nnf_init(ah, aw, bh, bw), oracle.make_params(Cn, ah, aw, bh, bw, iters=iters, rs_max=rs))
        assert np.array_equal(ann, g[key + "_ann"]), key
        assert np.array_equal(annd.view(np.uint32), g[key + "_annd"].view(np.uint32)), key
    assert np.array_equal(oracle.xorwow_raw(0, 8), g["xorwow_seed0"])
    assert np.array_equal(oracle.xorwow_raw(699, 8), g["xorwow_seed699"])


@pytest.mark.parametrize("Cn,ah,aw,bh,bw,rs", [(64, 40, 36, 37, 41, 9), (128, 31, 33, 33, 29, 6), (256, 24, 24, 24, 24, 4)])
def test_d4_unchanged_source_skip_never_changes_the_field(Cn, ah, aw, bh, bw, rs):
    """D4 only drops candidates that would be rejected again: NNF and distances are bit-identical with and without it,
    at every iteration count, while the number of evaluated candidates drops."""
    a = oracle.l2norm_hwc(synth.feature_volume(41, ah, aw, Cn, smooth=3))
    b = oracle.l2norm_hwc(synth.feature_volume(42, bh, bw, Cn, smooth=3))
    init = oracle.nnf_init(ah, aw, bh, bw)
    for iters in (1, 2, 3, 10):
        p = oracle.make_params(Cn, ah, aw, bh, bw, iters=iters, rs_max=rs)
        ann1, annd1, st1 = oracle.patchmatch(a, b, init, p, d4=True)
        ann0, annd0, st0 = oracle.patchmatch(a, b, init, p, d4=False)
        assert np.array_equal(ann1, ann0) and np.array_equal(annd1.view(np.uint32), annd0.view(np.uint32))
        assert st1[0] == st0[0] and st1[2] == st0[2] == st0[1]
        assert st1[1] <= st0[1]
        if iters == 1:
            assert st1[1] == st0[1]  # nothing to remember in the first iteration
        if iters == 10:
            assert st1[1] < 0.8 * st0[1]
