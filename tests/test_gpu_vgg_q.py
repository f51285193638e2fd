"""GPU parity tests of the exact fixed-point tensor-core convolution engine (engine 3, conv_i8.cu, tcgen05 kind::i8)
against oracle/vgg.py: BIT-EXACT, layer by layer, for the whole trunk and end to end."""
import os

import numpy as np
import pytest

from oracle import pipeline, synth, vgg

pytestmark = pytest.mark.gpu


def to_dev(x, dev):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    torch.cuda.synchronize()
    return t


@pytest.fixture(scope="module")
def weights():
    return synth.vgg19_weights(19)


@pytest.fixture(scope="module")
def qctx(pkg, weights):
    c = pkg.Context(0)
    c.load_vgg19_weights(weights)
    c.set_vgg_engine(3)
    yield c
    c.close()


@pytest.mark.parametrize("H,W,cin,cout", [(37, 53, 64, 64), (19, 21, 64, 128), (24, 40, 128, 128), (33, 17, 128, 256),
                                          (45, 29, 256, 256), (16, 8, 256, 512), (11, 13, 512, 512), (8, 16, 512, 512)])
def test_single_layer_accumulators_and_outputs_are_bit_exact(qctx, dev, H, W, cin, cout):
    """One layer, ragged tiles and image borders (TMA zero fill) included: the four raw INT32 TMEM accumulators equal the
    oracle's integer sums, and the FP32 outputs are identical bit for bit."""
    rng = np.random.default_rng(H * 1000 + W + cin)
    x = np.abs(rng.standard_normal((H, W, cin))).astype(np.float32) * 1.7
    x[rng.random((H, W, cin)) < 0.3] = 0
    w = (rng.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / (9 * cin))).astype(np.float32)
    b = (0.1 * rng.standard_normal(cout)).astype(np.float32)
    ref, racc = vgg.q_conv3x3_relu(x, w, b, return_acc=True)
    out, acc = qctx.conv3x3_fixedpoint(to_dev(x, dev), w, b, debug_acc=True)
    qctx.synchronize()
    acc = acc.cpu().numpy().reshape(4, H, W, cout)
    for d in range(4):
        assert np.array_equal(acc[d].astype(np.int64), racc[d]), f"accumulator {d} differs"
    assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32))


def test_extreme_values_quantise_like_the_oracle(qctx, dev):
    """The top of a binade (xq = 2^31 - 128: the carry chain reaches the leading digit), exact powers of two, tiny values
    and an all-zero tensor."""
    cin = cout = 64
    rng = np.random.default_rng(5)
    w = (rng.standard_normal((cout, cin, 3, 3)) * 0.05).astype(np.float32)
    b = np.zeros(cout, np.float32)
    x = np.zeros((16, 16, cin), np.float32)
    x[..., 0] = np.float32(2.0) - np.float32(2.0 ** -23)     # largest FP32 below 2
    x[..., 1] = 1.0
    x[..., 2] = 2.0 ** -30
    x[..., 3] = np.nextafter(np.float32(1.0), np.float32(0.0))
    for xx in (x, np.zeros_like(x)):
        ref = vgg.q_conv3x3_relu(xx, w, b)
        out = qctx.conv3x3_fixedpoint(to_dev(xx, dev), w, b)
        qctx.synchronize()
        assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("h,w", [(64, 48), (97, 131), (160, 160)])
def test_trunk_is_bit_exact_against_the_fixed_point_oracle(qctx, dev, weights, h, w):
    """All five feature maps of the engine-3 trunk equal oracle/vgg.py::features_fixedpoint bit for bit."""
    img, _ = synth.pair(11, h, w)
    feats = qctx.predict(to_dev(img, dev), 0)
    qctx.synchronize()
    ref = vgg.features_fixedpoint(img, weights, 0)
    for l in range(5):
        g = feats[l].cpu().numpy()
        assert g.shape == ref[l].shape
        assert np.array_equal(g.view(np.uint32), ref[l].view(np.uint32)), f"level {l}: {(g != ref[l]).sum()} of {g.size} values differ"


@pytest.mark.parametrize("bn,kb", [(64, 64), (64, 128), (128, 64)])
def test_every_tile_configuration_gives_the_same_bits(qctx, dev, weights, bn, kb):
    """The result is a pure function of the inputs: every tile shape / swizzle mode (NCT_I8_BN, NCT_I8_KB) produces the
    same feature maps as the default configuration."""
    img, _ = synth.pair(12, 80, 112)
    t = to_dev(img, dev)
    base = qctx.predict(t, 0)
    qctx.synchronize()  # the context runs on its own non-blocking stream: results are only there after this
    base = [f.cpu().numpy() for f in base]
    os.environ["NCT_I8_BN"], os.environ["NCT_I8_KB"] = str(bn), str(kb)
    try:
        got = qctx.predict(t, 0)
        qctx.synchronize()
        for l in range(5):
            assert np.array_equal(base[l].view(np.uint32), got[l].cpu().numpy().view(np.uint32))
    finally:
        del os.environ["NCT_I8_BN"], os.environ["NCT_I8_KB"]


def test_fixed_point_features_are_closer_to_fp64_than_the_fp32_order(qctx, dev, weights):
    """Accuracy, not only reproducibility: against the FP32-order engine the maps agree to ~1e-5 of the range."""
    img, _ = synth.pair(7, 128, 128)
    t = to_dev(img, dev)
    q = qctx.predict(t, 0)
    qctx.synchronize()
    qctx.set_vgg_engine(0)
    f = qctx.predict(t, 0)
    qctx.synchronize()
    qctx.set_vgg_engine(3)
    for l in range(5):
        a, b = q[l].cpu().numpy(), f[l].cpu().numpy()
        assert np.abs(a - b).max() / np.abs(b).max() < 2e-5


@pytest.mark.parametrize("deepest", [1, 2, 3, 4])
def test_truncated_forward_equals_full_forward(qctx, dev, deepest):
    img, _ = synth.pair(2, 80, 72)
    t = to_dev(img, dev)
    full = qctx.predict(t, 0)
    part = qctx.predict(t, deepest)
    qctx.synchronize()
    for l in range(deepest, 5):
        assert np.array_equal(part[l].cpu().numpy(), full[l].cpu().numpy())


@pytest.mark.parametrize("seed,ch,cw,sh,sw", [(4, 128, 128, 128, 128), (8, 120, 152, 136, 104), (9, 256, 256, 256, 256)])
def test_pipeline_end_to_end_with_the_tensor_core_engine_against_independent_oracle(qctx, dev, weights, seed, ch, cw, sh, sw):
    """The north star's end-to-end bar with the DEFAULT (tensor-core) engine: fully independent runs -- the oracle computes
    its own fixed-point features, NNFs, votes, neighbours, canonical-order CG and a direct WLS solve."""
    cnt, stl = synth.pair(seed, ch, cw, sh, sw)
    out = qctx.transfer_pair(cnt, stl)
    ref = pipeline.transfer_pair(cnt, stl, None, features_fn=lambda img, deepest: vgg.features_fixedpoint(img, weights, deepest),
                                 cg_mode="canonical")
    ps = pipeline.psnr(out, ref)
    ndiff = int((out != ref).sum())
    print(f"engine 3 end-to-end vs independent fixed-point oracle ({ch}x{cw} / {sh}x{sw}): PSNR {ps:.1f} dB, {ndiff} of {out.size} bytes differ")
    assert ps >= 50.0


@pytest.mark.parametrize("persist", ["0", "1"])
def test_one_tile_and_persistent_kernels_give_the_same_bits(qctx, dev, weights, persist, monkeypatch):
    """conv3x3_i8_kernel (one tile per CTA) and conv3x3_i8_persistent_kernel (one CTA per SM walking the tile list, two
    accumulator sets in TMEM at BN = 64) compute the same integer sums: forcing either for EVERY layer reproduces the default
    mix (persistent for the shallow layers only) bit for bit."""
    img, _ = synth.pair(13, 112, 96)
    t = to_dev(img, dev)
    base = qctx.predict(t, 0)
    qctx.synchronize()
    base = [f.cpu().numpy() for f in base]
    monkeypatch.setenv("NCT_I8_PERSIST", persist)
    got = qctx.predict(t, 0)
    qctx.synchronize()
    for l in range(5):
        assert np.array_equal(base[l].view(np.uint32), got[l].cpu().numpy().view(np.uint32))
    ref = vgg.features_fixedpoint(img, weights, 0)
    for l in range(5):
        assert np.array_equal(got[l].cpu().numpy().view(np.uint32), ref[l].view(np.uint32))
