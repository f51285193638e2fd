"""GPU parity tests of the VGG-19 trunk against oracle/vgg.py (FP32, tolerance like Caffe's own conv test: 1e-4)."""
import numpy as np
import pytest

from oracle import synth, vgg

pytestmark = pytest.mark.gpu


def to_dev(x, dev):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    torch.cuda.synchronize()  # libnct contexts run on their own non-blocking stream: the copy must have landed
    return t


@pytest.fixture(scope="module")
def weights():
    return synth.vgg19_weights(19)


@pytest.fixture(scope="module")
def vctx(pkg, weights):
    c = pkg.Context(0)
    c.load_vgg19_weights(weights)
    yield c
    c.close()


@pytest.mark.parametrize("h,w", [(64, 48), (97, 131), (256, 256)])
def test_vgg19_features_match_oracle(vctx, dev, weights, h, w):
    img, _ = synth.pair(0, h, w)
    feats = vctx.predict(to_dev(img, dev), 0)
    vctx.synchronize()
    ref = vgg.features(img, weights, 0)
    dims = vctx.level_dims(h, w)
    for l in range(5):
        g = feats[l].cpu().numpy()
        assert g.shape == ref[l].shape == (dims[l][1], dims[l][2], dims[l][0])
        scale = np.abs(ref[l]).max()
        err = np.abs(g - ref[l]).max() / scale
        assert err < 1e-4, f"level {l}: max error {err:.2e} of the feature range"
        assert (g >= 0).all()  # post-ReLU


@pytest.mark.parametrize("h,w", [(64, 48), (97, 131), (160, 160)])
def test_fp32_engine_is_bit_exact_against_the_canonical_order_oracle(vctx, dev, weights, h, w):
    """Engine 0 (FP32 CUDA cores) sums tap-major / channel-minor with one fmaf chain per output -- the order
    oracle/conv_oracle.c defines -- so all five feature maps are bit-identical, ragged sizes included."""
    img, _ = synth.pair(11, h, w)
    vctx.set_vgg_engine(0)
    feats = vctx.predict(to_dev(img, dev), 0)
    vctx.synchronize()
    ref = vgg.features_canonical(img, weights, 0)
    for l in range(5):
        g = feats[l].cpu().numpy()
        assert np.array_equal(g.view(np.uint32), ref[l].view(np.uint32)), f"level {l}: {(g != ref[l]).sum()} of {g.size} values differ"


def test_vgg19_im2col_oracle_agrees_with_direct(weights):
    img, _ = synth.pair(1, 40, 36)
    a = vgg.features(img, weights, 0)
    b = vgg.features(img, weights, 0, im2col=True)
    for l in range(5):
        assert np.abs(a[l] - b[l]).max() / np.abs(a[l]).max() < 1e-5


@pytest.mark.parametrize("deepest", [1, 2, 3, 4])
def test_truncated_forward_equals_full_forward(vctx, dev, deepest):
    img, _ = synth.pair(2, 80, 72)
    t = to_dev(img, dev)
    full = vctx.predict(t, 0)
    part = vctx.predict(t, deepest)
    vctx.synchronize()
    for l in range(5):
        if l < deepest:
            assert part[l] is None
        else:
            assert np.array_equal(part[l].cpu().numpy(), full[l].cpu().numpy())


def test_level_dims_ceil_mode(vctx):
    assert [d[1] for d in vctx.level_dims(700, 700)] == [44, 88, 175, 350, 700]
    assert [d[2] for d in vctx.level_dims(520, 352)] == [22, 44, 88, 176, 352]
    assert [d[0] for d in vctx.level_dims(700, 700)] == [512, 512, 256, 128, 64]


def test_missing_weights_fail_loudly(pkg, dev):
    import torch

    c = pkg.Context(0)
    with pytest.raises(pkg.NctError):
        c.predict(to_dev(np.zeros((32, 32, 3), np.uint8), dev))
    c.close()


@pytest.mark.parametrize("engine,tol", [(1, 2e-2), (2, 5e-4)])
@pytest.mark.parametrize("h,w", [(64, 48), (97, 131), (256, 256)])
def test_tensorcore_engines_match_fp32_engine(pkg, vctx, dev, weights, h, w, engine, tol):
    """tcgen05 convolutions against the FP32 CUDA-core engine, borders (TMA zero fill) and ragged tiles included.
    engine 1 = plain kind::tf32 (operands truncated to a 10-bit mantissa): ~1e-2 of the feature range after 13 layers;
    engine 2 = 3xTF32 (exact hi/lo operand split, three MMAs): ~2e-4 of the range -- operand rounding is gone, what
    remains is the accumulation inside TMEM over K = 9*Cin."""
    img, _ = synth.pair(7, h, w)
    t = to_dev(img, dev)
    ref = vctx.predict(t, 0)
    vctx.synchronize()
    c = pkg.Context(0)
    c.load_vgg19_weights(weights)
    c.set_vgg_engine(engine)
    got = c.predict(t, 0)
    c.synchronize()
    worst = 0.0
    for l in range(5):
        r, g = ref[l].cpu().numpy(), got[l].cpu().numpy()
        assert np.isfinite(g).all()
        worst = max(worst, float(np.abs(g - r).max() / np.abs(r).max()))
    print(f"engine {engine} {h}x{w}: max err {worst:.2e} of the feature range")
    assert worst < tol
    c.close()
