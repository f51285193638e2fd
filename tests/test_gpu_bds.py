"""GPU parity tests of the BDS votes against oracle/bds_oracle.c (bit-exact: the gather order is fixed)."""
import numpy as np
import pytest

import oracle
from oracle import synth

pytestmark = pytest.mark.gpu


def to_dev(x, dev):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    torch.cuda.synchronize()  # libnct contexts run on their own non-blocking stream: the copy must have landed
    return t


def rand_nnf(rng, n, th, tw):
    return ((rng.integers(0, th, n).astype(np.uint32) << 12) | rng.integers(0, tw, n).astype(np.uint32))


def clustered_nnf(rng, n, th, tw):
    """many sources map to few targets: long inverse lists"""
    ty = rng.integers(0, max(1, th // 6), n)
    tx = rng.integers(0, max(1, tw // 6), n)
    return ((ty.astype(np.uint32) << 12) | tx.astype(np.uint32))


@pytest.mark.parametrize("ah,aw,bh,bw", [(18, 22, 20, 19), (44, 44, 44, 44), (175, 175, 160, 190), (1, 5, 5, 1)])
@pytest.mark.parametrize("bds", [0.0, 1.0, 2.0, 8.0])
@pytest.mark.parametrize("gen", [rand_nnf, clustered_nnf])
def test_reconstruct_bds_bit_exact(ctx, dev, ah, aw, bh, bw, bds, gen):
    cnt, stl = synth.pair(2, ah, aw, bh, bw)
    rng = np.random.default_rng(ah * 7 + bw)
    ann, bnn = gen(rng, ah * aw, bh, bw), gen(rng, bh * bw, ah, aw)
    g = ctx.reconstruct_bds(to_dev(cnt, dev), to_dev(stl, dev), to_dev(ann.view(np.int32), dev), to_dev(bnn.view(np.int32), dev), 1.0, bds)
    ctx.synchronize()
    o = oracle.reconstruct_bds(cnt, stl, ann, bnn, 1.0, bds)
    assert np.array_equal(g.cpu().numpy(), o)


@pytest.mark.parametrize("Cn,ah,aw,bh,bw", [(64, 30, 34, 28, 37), (64, 21, 23, 19, 25), (128, 25, 25, 25, 25), (256, 20, 23, 22, 19), (512, 15, 15, 15, 15), (16, 9, 9, 9, 9)])
@pytest.mark.parametrize("gen", [rand_nnf, clustered_nnf])
def test_bds_feature_error_bit_exact(ctx, dev, Cn, ah, aw, bh, bw, gen):
    c = oracle.l2norm_hwc(synth.feature_volume(1, ah, aw, Cn))
    s = synth.feature_volume(2, bh, bw, Cn) * np.float32(5.0)
    rng = np.random.default_rng(Cn + ah)
    ann, bnn = gen(rng, ah * aw, bh, bw), gen(rng, bh * bw, ah, aw)
    ge, gv = ctx.bds_feature_error(to_dev(c, dev), to_dev(s, dev), to_dev(ann.view(np.int32), dev), to_dev(bnn.view(np.int32), dev), 1.0, 2.0, want_vote=True)
    ctx.synchronize()
    oe, ov = oracle.bds_feature_error(c, s, ann, bnn, 1.0, 2.0, mode=0, want_vote=True)
    assert np.array_equal(gv.cpu().numpy().view(np.uint32), ov.view(np.uint32))
    assert np.array_equal(ge.cpu().numpy().view(np.uint32), oe.view(np.uint32))
    # and within the documented tolerance of the reference's own summation order
    oe1 = oracle.bds_feature_error(c, s, ann, bnn, 1.0, 2.0, mode=1)
    assert np.abs(ge.cpu().numpy() - oe1).max() < 1e-5


def test_bds_after_patchmatch_level_shape(pkg, ctx, dev):
    """relu3_1 shape of a 700^2 pair (175^2 x 256): PatchMatch both ways, then both votes, vs the oracle chain."""
    import torch

    n, Cn = 175, 256
    a_raw = synth.feature_volume(51, n, n, Cn, smooth=8)
    b_raw = synth.feature_volume(52, n, n, Cn, smooth=8)
    cnt, stl = synth.pair(3, n, n)
    ta, tb = to_dev(a_raw, dev), to_dev(b_raw, dev)
    na, nb = ctx.norm(ta), ctx.norm(tb)
    ann = torch.empty(n * n, dtype=torch.int32, device=dev)
    bnn = torch.empty(n * n, dtype=torch.int32, device=dev)
    annd = torch.empty(n * n, dtype=torch.float32, device=dev)
    bnnd = torch.empty(n * n, dtype=torch.float32, device=dev)
    ctx.init_ann(ann, n, n, n, n)
    ctx.init_ann(bnn, n, n, n, n)
    ctx.patchmatch_bidir(na, nb, ann, annd, bnn, bnnd, pkg.make_params(Cn, n, n, n, n, iters=2, rs_max=10))
    err = ctx.bds_feature_error(na, tb, ann, bnn, 1.0, 2.0)
    rec = ctx.reconstruct_bds(to_dev(cnt, dev), to_dev(stl, dev), ann, bnn, 1.0, 2.0)
    ctx.synchronize()
    oa, ob = oracle.l2norm_hwc(a_raw), oracle.l2norm_hwc(b_raw)
    p = oracle.make_params(Cn, n, n, n, n, iters=2, rs_max=10)
    o_ann, _, _ = oracle.patchmatch(oa, ob, oracle.nnf_init(n, n, n, n), p)
    o_bnn, _, _ = oracle.patchmatch(ob, oa, oracle.nnf_init(n, n, n, n), p)
    assert np.array_equal(ann.cpu().numpy().view(np.uint32), o_ann)
    assert np.array_equal(bnn.cpu().numpy().view(np.uint32), o_bnn)
    assert np.array_equal(err.cpu().numpy().view(np.uint32), oracle.bds_feature_error(oa, b_raw, o_ann, o_bnn, 1.0, 2.0).view(np.uint32))
    assert np.array_equal(rec.cpu().numpy(), oracle.reconstruct_bds(cnt, stl, o_ann, o_bnn, 1.0, 2.0))
