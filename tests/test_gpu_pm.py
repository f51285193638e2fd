"""GPU parity tests (through the C ABI) of the correspondence stage against oracle/pm_oracle.c.
Bar: bit-exact (NNF indices and FP32 distances)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from oracle import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def to_dev(x, dev):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    torch.cuda.synchronize()  # libnct contexts run on their own non-blocking stream: the copy must have landed
    return t


def u32(t):
    return t.cpu().numpy().view(np.uint32)


def run_gpu_pm(pkg, ctx, dev, na, nb, ah, aw, bh, bw, Cn, iters, rs, bidir=True, init=None):
    import torch

    ann = torch.empty(ah * aw, dtype=torch.int32, device=dev)
    bnn = torch.empty(bh * bw, dtype=torch.int32, device=dev)
    annd = torch.zeros(ah * aw, dtype=torch.float32, device=dev)
    bnnd = torch.zeros(bh * bw, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()  # the fills run on torch's stream, PatchMatch on the context's
    if init is None:
        ctx.init_ann(ann, ah, aw, bh, bw)
        ctx.init_ann(bnn, bh, bw, ah, aw)
    else:
        ann.copy_(to_dev(init[0].view(np.int32), dev))
        bnn.copy_(to_dev(init[1].view(np.int32), dev))
    p_ab = pkg.make_params(Cn, ah, aw, bh, bw, iters=iters, rs_max=rs)
    if bidir:
        ctx.patchmatch_bidir(na, nb, ann, annd, bnn, bnnd, p_ab)
    else:
        ctx.patchmatch_single(na, nb, ann, annd, p_ab)
        ctx.patchmatch_single(nb, na, bnn, bnnd, pkg.make_params(Cn, bh, bw, ah, aw, iters=iters, rs_max=rs))
    ctx.synchronize()
    return u32(ann), annd.cpu().numpy(), u32(bnn), bnnd.cpu().numpy()


def test_xorwow_matches_curand_and_oracle(ctx, dev):
    import torch

    ncols, ndraws = 700, 120
    g = ctx.xorwow_table(ncols, ndraws)
    ctx.synchronize()
    o = oracle.xorwow_uniform_table(ncols, ndraws)
    assert np.array_equal(g.cpu().numpy().view(np.uint32), o.view(np.uint32))
    # the real generator: cuRAND device API, used exactly like NCT/GeneralizedPatchMatch.cu:54-66
    ref = C.CDLL(os.path.join(HERE, "cuda", "libcurand_ref.so"))
    out = torch.empty((ncols, ndraws), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    assert ref.curand_ref_table(C.c_void_p(out.data_ptr()), ncols, ndraws) == 0
    assert np.array_equal(out.cpu().numpy().view(np.uint32), o.view(np.uint32))


@pytest.mark.parametrize("shape", [(44, 44, 44, 44), (44, 30, 31, 47), (5, 7, 9, 11), (700, 700, 700, 700), (63, 63, 61, 65)])
def test_nnf_init(ctx, dev, shape):
    import torch

    ah, aw, bh, bw = shape
    ann = torch.empty(ah * aw, dtype=torch.int32, device=dev)
    ctx.init_ann(ann, ah, aw, bh, bw)
    ctx.synchronize()
    assert np.array_equal(u32(ann), oracle.nnf_init(ah, aw, bh, bw))


@pytest.mark.parametrize("shape", [(44, 44, 44, 44, 88, 88, 88, 88), (88, 88, 88, 88, 175, 175, 175, 175),
                                   (44, 45, 40, 47, 88, 89, 79, 93), (63, 63, 63, 63, 125, 125, 125, 125),
                                   (350, 350, 350, 350, 700, 700, 700, 700)])
def test_nnf_upsample(ctx, dev, shape):
    import torch

    ahh, awh, bhh, bwh, ah, aw, bh, bw = shape
    rng = np.random.default_rng(7)
    half = ((rng.integers(0, bhh, ahh * awh).astype(np.uint32) << 12) | rng.integers(0, bwh, ahh * awh).astype(np.uint32))
    t_half = to_dev(half.view(np.int32), dev)
    ann = torch.empty(ah * aw, dtype=torch.int32, device=dev)
    ctx.upsample(t_half, ahh, awh, ann, ah, aw, bh, bw)
    ctx.synchronize()
    assert np.array_equal(u32(ann), oracle.nnf_upsample(half, ahh, awh, ah, aw, bh, bw))


@pytest.mark.parametrize("Cn", [16, 64, 128, 256, 512])
def test_l2norm_bit_exact(ctx, dev, Cn):
    a = synth.feature_volume(3, 33, 29, Cn)
    a[5, 6, :] = 0  # zero-norm pixel -> zeros (oracle decision D5)
    g = ctx.norm(to_dev(a, dev))
    ctx.synchronize()
    assert np.array_equal(g.cpu().numpy().view(np.uint32), oracle.l2norm_hwc(a).view(np.uint32))


def test_layout_round_trip(ctx, dev):
    a = synth.feature_volume(4, 37, 41, 64)
    t = to_dev(a, dev)
    chw = ctx.hwc_to_chw(t)
    back = ctx.chw_to_hwc(chw)
    ctx.synchronize()
    assert np.array_equal(chw.cpu().numpy(), a.transpose(2, 0, 1))
    assert np.array_equal(back.cpu().numpy(), a)


@pytest.mark.parametrize("Cn,ah,aw,bh,bw,iters,rs", [
    (64, 20, 24, 22, 19, 3, 6), (128, 16, 16, 16, 16, 10, 4), (256, 12, 14, 13, 12, 2, 32), (512, 11, 11, 11, 11, 4, 2),
    (32, 10, 12, 12, 10, 2, 4), (16, 10, 10, 10, 10, 2, 4),
    (64, 50, 61, 47, 66, 10, 32), (128, 40, 40, 37, 45, 10, 16), (256, 31, 33, 30, 36, 10, 8), (512, 25, 25, 24, 27, 10, 3),
    (64, 3, 3, 3, 3, 2, 1), (128, 1, 9, 9, 1, 2, 4),  # ragged / degenerate sizes
])
@pytest.mark.parametrize("bidir", [True, False])
def test_patchmatch_bit_exact(pkg, ctx, dev, Cn, ah, aw, bh, bw, iters, rs, bidir):
    a = oracle.l2norm_hwc(synth.feature_volume(11, ah, aw, Cn))
    b = oracle.l2norm_hwc(synth.feature_volume(12, bh, bw, Cn))
    if ah == 1 or aw == 1 or bh == 1 or bw == 1:
        # init_Ann divides by (aw-1): undefined for 1-wide images in the reference; start from zeros
        init = (np.zeros(ah * aw, np.uint32), np.zeros(bh * bw, np.uint32))
    else:
        init = (oracle.nnf_init(ah, aw, bh, bw), oracle.nnf_init(bh, bw, ah, aw))
    g_ann, g_annd, g_bnn, g_bnnd = run_gpu_pm(pkg, ctx, dev, to_dev(a, dev), to_dev(b, dev), ah, aw, bh, bw, Cn, iters, rs, bidir, init)
    o_ann, o_annd, _ = oracle.patchmatch(a, b, init[0], oracle.make_params(Cn, ah, aw, bh, bw, iters=iters, rs_max=rs))
    o_bnn, o_bnnd, _ = oracle.patchmatch(b, a, init[1], oracle.make_params(Cn, bh, bw, ah, aw, iters=iters, rs_max=rs))
    assert np.array_equal(g_ann, o_ann), f"{(g_ann != o_ann).sum()} NNF entries differ"
    assert np.array_equal(g_annd.view(np.uint32), o_annd.view(np.uint32))
    assert np.array_equal(g_bnn, o_bnn)
    assert np.array_equal(g_bnnd.view(np.uint32), o_bnnd.view(np.uint32))


def test_patchmatch_golden(pkg, ctx, dev):
    g = np.load(os.path.join(HERE, "golden", "pm_golden.npz"))
    for key in [k[:-4] for k in g.files if k.endswith("_ann")]:
        Cn, ah, aw, bh, bw, iters, rs = [int(v) for v in g[key + "_cfg"]]
        a = oracle.l2norm_hwc(synth.feature_volume(11, ah, aw, Cn))
        b = oracle.l2norm_hwc(synth.feature_volume(12, bh, bw, Cn))
        g_ann, g_annd, _, _ = run_gpu_pm(pkg, ctx, dev, to_dev(a, dev), to_dev(b, dev), ah, aw, bh, bw, Cn, iters, rs)
        assert np.array_equal(g_ann, g[key + "_ann"]), key
        assert np.array_equal(g_annd.view(np.uint32), g[key + "_annd"].view(np.uint32)), key


def test_iters_zero_is_initial_distance(pkg, ctx, dev):
    Cn, H, W = 128, 21, 23
    a = oracle.l2norm_hwc(synth.feature_volume(11, H, W, Cn))
    b = oracle.l2norm_hwc(synth.feature_volume(12, H, W, Cn))
    g_ann, g_annd, _, _ = run_gpu_pm(pkg, ctx, dev, to_dev(a, dev), to_dev(b, dev), H, W, H, W, Cn, 0, 8)
    o_ann, o_annd, _ = oracle.patchmatch(a, b, oracle.nnf_init(H, W, H, W), oracle.make_params(Cn, H, W, H, W, iters=0, rs_max=8))
    assert np.array_equal(g_ann, o_ann) and np.array_equal(g_annd.view(np.uint32), o_annd.view(np.uint32))


@pytest.mark.parametrize("iters", list(range(1, 11)))
def test_config5_iteration_sweep(pkg, ctx, dev, iters):
    """BASELINE config 5: relu3_1 geometry of a 512^2 image, 256 x 128 x 128, rs_max = 8; parity at every
    iteration count and the ground-truth shift as a sanity check."""
    Cn, H, W = 256, 128, 128
    a, b = synth.pm_sweep_volumes(Cn, H, W)
    ta, tb = ctx.norm(to_dev(a, dev)), ctx.norm(to_dev(b, dev))
    ctx.count_evals(True)
    g_ann, g_annd, _, _ = run_gpu_pm(pkg, ctx, dev, ta, tb, H, W, H, W, Cn, iters, 8, bidir=False)
    ctx.count_evals(False)
    oa, ob = oracle.l2norm_hwc(a), oracle.l2norm_hwc(b)
    o_ann, o_annd, st = oracle.patchmatch(oa, ob, oracle.nnf_init(H, W, H, W), oracle.make_params(Cn, H, W, H, W, iters=iters, rs_max=8))
    assert np.array_equal(g_ann, o_ann)
    assert np.array_equal(g_annd.view(np.uint32), o_annd.view(np.uint32))
    if iters == 10:
        x, y = oracle.unpack(g_ann)
        gx, gy = np.meshgrid(np.arange(W), np.arange(H))
        inner = ((gx.ravel() + 7 < W) & (gy.ravel() - 3 >= 0))
        ok = (x == gx.ravel() + 7) & (y == gy.ravel() - 3)
        assert ok[inner].mean() > 0.99


@pytest.mark.parametrize("level,iters", [(0, 10), (1, 10), (2, 10), (3, 2), (4, 1)])
def test_level_shapes_of_700(pkg, ctx, dev, level, iters):
    """The five level shapes of BASELINE config 2 (700^2): 44^2x512, 88^2x512, 175^2x256, 350^2x128, 700^2x64,
    with the reference's rs_max per level (NCT/main.cu:77-83), both directions, vs the oracle."""
    sizes = [44, 88, 175, 350, 700]
    chans = [512, 512, 256, 128, 64]
    ranges = [700 // 16, 700 // 32, 700 // 64, 32, 32]
    n, Cn, rs = sizes[level], chans[level], ranges[level]
    a = synth.feature_volume(21 + level, n, n, Cn, smooth=max(2, n // 22))
    b = synth.feature_volume(31 + level, n, n, Cn, smooth=max(2, n // 22))
    ta, tb = ctx.norm(to_dev(a, dev)), ctx.norm(to_dev(b, dev))
    ctx.count_evals(True)
    g_ann, g_annd, g_bnn, g_bnnd = run_gpu_pm(pkg, ctx, dev, ta, tb, n, n, n, n, Cn, iters, rs)
    ev, ev_ref = ctx.patchmatch_stats()
    ctx.count_evals(False)
    oa, ob = oracle.l2norm_hwc(a), oracle.l2norm_hwc(b)
    o_ann, o_annd, st_a = oracle.patchmatch(oa, ob, oracle.nnf_init(n, n, n, n), oracle.make_params(Cn, n, n, n, n, iters=iters, rs_max=rs))
    o_bnn, o_bnnd, st_b = oracle.patchmatch(ob, oa, oracle.nnf_init(n, n, n, n), oracle.make_params(Cn, n, n, n, n, iters=iters, rs_max=rs))
    assert np.array_equal(g_ann, o_ann) and np.array_equal(g_bnn, o_bnn)
    assert np.array_equal(g_annd.view(np.uint32), o_annd.view(np.uint32))
    assert np.array_equal(g_bnnd.view(np.uint32), o_bnnd.view(np.uint32))
    # the kernel's own evaluation counters agree with the oracle's (the roofline's unit count)
    assert ev == st_a[1] + st_b[1] and ev_ref == st_a[0] + st_b[0]


def test_full_size_properties(pkg, ctx, dev):
    """700^2 x 64, 10 iterations (BASELINE config 2 finest level): too big for the oracle to be fast, so check
    size-independent properties: determinism, in-bounds, annd == dist(ann) (idempotence through an iters=0 call),
    and monotone improvement over the initial field."""
    import torch

    n, Cn = 700, 64
    a = synth.feature_volume(41, n, n, Cn, smooth=16)
    b = synth.feature_volume(42, n, n, Cn, smooth=16)
    ta, tb = ctx.norm(to_dev(a, dev)), ctx.norm(to_dev(b, dev))
    r1 = run_gpu_pm(pkg, ctx, dev, ta, tb, n, n, n, n, Cn, 10, 32)
    r2 = run_gpu_pm(pkg, ctx, dev, ta, tb, n, n, n, n, Cn, 10, 32)
    for x, y in zip(r1, r2):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    x, y = oracle.unpack(r1[0])
    assert x.max() < n and y.max() < n
    # idempotence: distances recomputed by an iters=0 call on the final field are identical
    ann = to_dev(r1[0].view(np.int32), dev)
    annd = torch.empty(n * n, dtype=torch.float32, device=dev)
    ctx.patchmatch_single(ta, tb, ann, annd, pkg.make_params(Cn, n, n, n, n, iters=0, rs_max=32))
    ctx.synchronize()
    assert np.array_equal(annd.cpu().numpy().view(np.uint32), r1[1].view(np.uint32))
    r0 = run_gpu_pm(pkg, ctx, dev, ta, tb, n, n, n, n, Cn, 0, 32)
    assert np.all(r1[1] <= r0[1]) and r1[1].mean() < r0[1].mean()
    # spot-check 64 random entries against the oracle's distance function
    oa, ob = oracle.l2norm_hwc(a), oracle.l2norm_hwc(b)
    rng = np.random.default_rng(0)
    for p in rng.integers(0, n * n, 64):
        assert r1[1][p] == np.float32(oracle.dist_canon(oa, ob, int(p % n), int(p // n), int(x[p]), int(y[p])))


def test_bad_arguments_are_rejected(pkg, ctx, dev):
    import torch

    t = torch.zeros(16 * 16 * 48, dtype=torch.float32, device=dev)
    ann = torch.zeros(256, dtype=torch.int32, device=dev)
    annd = torch.zeros(256, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    with pytest.raises(pkg.NctError):  # C = 48 unsupported
        ctx.patchmatch_single(t, t, ann, annd, pkg.make_params(48, 16, 16, 16, 16))
    with pytest.raises(pkg.NctError):  # patch 5 unsupported (reference fixes 3, CT/Config.h:70)
        ctx.patchmatch_single(t, t, ann, annd, pkg.make_params(64, 16, 16, 16, 16, patch=5))


def test_finest_level_700x700x64_ten_iterations_matches_the_committed_golden(pkg, ctx, dev):
    """The finest PatchMatch level at BASELINE's headline size, all 10 iterations, both directions, against
    tests/golden/fullsize_golden.npz (oracle run offline by tests/golden/make_fullsize_golden.py): CRC32 of the whole
    NNF and distance arrays, every 97th entry stored in full, and the evaluation counters."""
    import os
    import zlib

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize_golden.npz"))
    Cn, ah, aw, bh, bw, iters, rs, sa, sb = (int(v) for v in g["pm700_cfg"])
    a = to_dev(oracle.l2norm_hwc(synth.feature_volume(sa, ah, aw, Cn)), dev)
    b = to_dev(oracle.l2norm_hwc(synth.feature_volume(sb, bh, bw, Cn)), dev)
    import torch

    ann = torch.empty(ah * aw, dtype=torch.int32, device=dev)
    bnn = torch.empty(bh * bw, dtype=torch.int32, device=dev)
    annd = torch.empty(ah * aw, dtype=torch.float32, device=dev)
    bnnd = torch.empty(bh * bw, dtype=torch.float32, device=dev)
    ctx.init_ann(ann, ah, aw, bh, bw)
    ctx.init_ann(bnn, bh, bw, ah, aw)
    ctx.count_evals(True)
    ctx.patchmatch_bidir(a, b, ann, annd, bnn, bnnd, pkg.make_params(Cn, ah, aw, bh, bw, iters=iters, rs_max=rs))
    ev, _ = ctx.patchmatch_stats()
    ctx.count_evals(False)
    ga, gb = ann.cpu().numpy().view(np.uint32), bnn.cpu().numpy().view(np.uint32)
    gad, gbd = annd.cpu().numpy(), bnnd.cpu().numpy()
    assert np.array_equal(ga[::97], g["pm700_ann_s97"]) and np.array_equal(gb[::97], g["pm700_bnn_s97"])
    assert np.array_equal(gad[::97].view(np.uint32), g["pm700_annd_s97"].view(np.uint32))
    crcs = [zlib.crc32(np.ascontiguousarray(v).tobytes()) for v in (ga, gad, gb, gbd)]
    assert crcs == [int(v) for v in g["pm700_crc"]]
    assert ev == int(g["pm700_evals"].sum())


@pytest.mark.parametrize("Cn,ah,aw,bh,bw,iters,rs", [(64, 41, 37, 39, 44, 4, 8), (128, 33, 35, 30, 31, 3, 4), (256, 20, 24, 22, 19, 3, 32),
                                                     (512, 12, 13, 11, 14, 10, 2), (64, 96, 96, 96, 96, 10, 32)])
def test_fp16_feature_store_is_bit_exact_against_the_oracle_on_rounded_volumes(pkg, ctx, dev, Cn, ah, aw, bh, bw, iters, rs):
    """Throughput mode beyond the reference: the PatchMatch volumes stored as FP16.  nct_l2norm_f16 = the
    canonical normalisation rounded to nearest-even half; the kernels convert every half to FP32 exactly and keep their
    arithmetic, so NNF, distances and evaluation counters equal the ORACLE run on the rounded volumes, bit for bit."""
    import torch

    a = synth.feature_volume(51, ah, aw, Cn)
    b = synth.feature_volume(52, bh, bw, Cn)
    na16, nb16 = ctx.norm_f16(to_dev(a, dev)), ctx.norm_f16(to_dev(b, dev))
    ctx.synchronize()
    oa16 = oracle.l2norm_hwc(a).astype(np.float16)
    ob16 = oracle.l2norm_hwc(b).astype(np.float16)
    assert np.array_equal(na16.cpu().numpy().view(np.uint16), oa16.view(np.uint16))
    assert np.array_equal(nb16.cpu().numpy().view(np.uint16), ob16.view(np.uint16))
    ann = torch.empty(ah * aw, dtype=torch.int32, device=dev)
    bnn = torch.empty(bh * bw, dtype=torch.int32, device=dev)
    annd = torch.empty(ah * aw, dtype=torch.float32, device=dev)
    bnnd = torch.empty(bh * bw, dtype=torch.float32, device=dev)
    ctx.init_ann(ann, ah, aw, bh, bw)
    ctx.init_ann(bnn, bh, bw, ah, aw)
    ctx.count_evals(True)
    ctx.patchmatch_bidir(na16, nb16, ann, annd, bnn, bnnd, pkg.make_params(Cn, ah, aw, bh, bw, iters=iters, rs_max=rs))
    ev, _ = ctx.patchmatch_stats()
    ctx.count_evals(False)
    fa, fb = oa16.astype(np.float32), ob16.astype(np.float32)
    o_ann, o_annd, st_a = oracle.patchmatch(fa, fb, oracle.nnf_init(ah, aw, bh, bw), oracle.make_params(Cn, ah, aw, bh, bw, iters=iters, rs_max=rs))
    o_bnn, o_bnnd, st_b = oracle.patchmatch(fb, fa, oracle.nnf_init(bh, bw, ah, aw), oracle.make_params(Cn, bh, bw, ah, aw, iters=iters, rs_max=rs))
    assert np.array_equal(ann.cpu().numpy().view(np.uint32), o_ann) and np.array_equal(bnn.cpu().numpy().view(np.uint32), o_bnn)
    assert np.array_equal(annd.cpu().numpy().view(np.uint32), o_annd.view(np.uint32))
    assert np.array_equal(bnnd.cpu().numpy().view(np.uint32), o_bnnd.view(np.uint32))
    assert ev == st_a[1] + st_b[1]
    # single direction through its own entry point
    ann1 = torch.empty(ah * aw, dtype=torch.int32, device=dev)
    annd1 = torch.empty(ah * aw, dtype=torch.float32, device=dev)
    ctx.init_ann(ann1, ah, aw, bh, bw)
    ctx.patchmatch_single(na16, nb16, ann1, annd1, pkg.make_params(Cn, ah, aw, bh, bw, iters=iters, rs_max=rs))
    ctx.synchronize()
    assert np.array_equal(ann1.cpu().numpy().view(np.uint32), o_ann)


def test_fp16_feature_store_rejects_what_it_does_not_cover(pkg, ctx, dev):
    import torch

    t = torch.zeros((16, 16, 32), dtype=torch.float16, device=dev)
    ann = torch.zeros(256, dtype=torch.int32, device=dev)
    annd = torch.zeros(256, dtype=torch.float32, device=dev)
    with pytest.raises(pkg.NctError):
        ctx.patchmatch_single(t, t, ann, annd, pkg.make_params(32, 16, 16, 16, 16))            # C < 64
    t64 = torch.zeros((16, 16, 64), dtype=torch.float16, device=dev)
    with pytest.raises(pkg.NctError):
        ctx.patchmatch_single(t64, t64, ann, annd, pkg.make_params(64, 16, 16, 16, 16, iters=0))  # iters = 0
