"""oracle/pipeline.py -- CPU restatement of transfer_color_single_bds (NCT/main.cu:47-454), composed from the
per-stage oracles.  TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).

`transfer_pair` is also the "port" CPU baseline of bench.py (cpu_baseline / --impl reference): VGG as im2col + SGEMM
(torch-CPU), PatchMatch / votes / clustering / k-NN in C with OpenMP, colour solves with scipy -- the reference's own
CPU-side algorithmic structure.

Hooks: `features_fn(image_bgr, deepest_level)` lets a test inject feature maps (e.g. the GPU's) so the downstream
stages can be compared in lock step; `on_level(level, dict)` receives every intermediate of a level.
"""
from __future__ import annotations

import time

import numpy as np

from . import color, vgg
from . import pm as _pm


def level_dims(h, w):
    ch = [512, 512, 256, 128, 64]
    dims = [None] * 5
    for l in range(4, -1, -1):
        dims[l] = (ch[l], h, w)
        h = -(-(h - 2) // 2) + 1
        w = -(-(w - 2) // 2) + 1
    return dims


DEFAULT_CFG = dict(bds=2.0, eps=0.6, nl=2.0, l=0.125, w=0.024, clusters=10, knum=8, alpha=1.2, pm_iters=10, kmeans_iters=11)


def transfer_pair(cnt, stl, weights=None, cfg=None, features_fn=None, on_level=None, stop_after_level=4, timings=None,
                  im2col=True, pm_mode="canonical", result_hook=None, cg_mode="reference", pm_fn=None, feature_store="f32"):
    """cnt, stl: uint8 BGR (H, W, 3).  Returns the uint8 BGR result (content size).
    pm_mode: "canonical" = deterministic jump-flood oracle (the parity target); "reference" = reference-semantics
    in-place serial PatchMatch in the reference's layout / summation order (CPU baseline only).
    cg_mode: "reference" = explicit A, A^T A by scipy (the reference's structure; summation order unspecified);
    "canonical" = the defined-order matrix-free CG of oracle/cg_oracle.c (decision N1) -- with canonical features
    (vgg.features_canonical) this makes the whole oracle run a bit-level target for the product.
    feature_store: "f32" (the reference's storage) or "f16" = the product's FP16 feature-store mode: the two normalised
    volumes PatchMatch gathers from are rounded to IEEE half (round to nearest even) and used as exact FP32 values; every
    other stage (BDS votes, feature error, clustering) keeps the FP32 volumes.
    pm_fn: optional replacement of the PatchMatch stage, pm_fn(nC, nS, ann, bnn, p_ab, p_ba) -> (ann, annd, bnn, bnnd)
    (tests plug the reference's own racy kernel in here to measure the reference's run-to-run noise floor)."""
    cfg = DEFAULT_CFG | (cfg or {})
    t_acc = timings if timings is not None else {}

    def tick(name, t0):
        t_acc[name] = t_acc.get(name, 0.0) + time.perf_counter() - t0

    if features_fn is None:
        def features_fn(img, deepest):
            return vgg.features(img, weights, deepest, im2col=im2col)

    ch, cw, _ = cnt.shape
    sh, sw, _ = stl.shape
    dc, ds = level_dims(ch, cw), level_dims(sh, sw)
    max_len = max(cw, ch, sw, sh)
    rng = [max_len // 16, max_len // 32, max_len // 64, 32, 32]
    cnt_lab_full_d = color.bgr2lab_u8(cnt).astype(np.float64) * (1.0 / 255.0)

    t0 = time.perf_counter()
    featC = features_fn(cnt, 0)
    featS = features_fn(stl, 0)
    tick("vgg", t0)
    cnt_imgs = color.pyramid(cnt, [(d[1], d[2]) for d in dc])
    stl_imgs = color.pyramid(stl, [(d[1], d[2]) for d in ds])

    t0 = time.perf_counter()
    nC0 = _pm.l2norm_hwc(featC[0])
    labels, nlabels = _pm.kmeans_labels(nC0.reshape(-1, dc[0][0]), cfg["clusters"], cfg["kmeans_iters"])
    tick("kmeans", t0)

    ann = bnn = None
    result = cnt
    for l in range(5):
        Cn, ah, aw = dc[l]
        _, bh, bw = ds[l]
        t0 = time.perf_counter()
        if l == 0:
            ann = _pm.nnf_init(ah, aw, bh, bw)
            bnn = _pm.nnf_init(bh, bw, ah, aw)
        else:
            ann = _pm.nnf_upsample(ann, dc[l - 1][1], dc[l - 1][2], ah, aw, bh, bw)
            bnn = _pm.nnf_upsample(bnn, ds[l - 1][1], ds[l - 1][2], bh, bw, ah, aw)
        nS = _pm.l2norm_hwc(featS[l])
        nC = _pm.l2norm_hwc(featC[l])
        p_ab = _pm.make_params(Cn, ah, aw, bh, bw, cfg["pm_iters"], rng[l])
        p_ba = _pm.make_params(Cn, bh, bw, ah, aw, cfg["pm_iters"], rng[l])
        nC_pm, nS_pm = nC, nS
        if feature_store == "f16":
            nC_pm = nC.astype(np.float16).astype(np.float32)
            nS_pm = nS.astype(np.float16).astype(np.float32)
        elif feature_store != "f32":
            raise ValueError("feature_store must be 'f32' or 'f16'")
        if pm_fn is not None:
            ann, annd, bnn, bnnd = pm_fn(nC, nS, ann, bnn, p_ab, p_ba)
            st_a = st_b = (0, 0)
        elif pm_mode == "canonical":
            ann, annd, st_a = _pm.patchmatch(nC_pm, nS_pm, ann, p_ab)
            bnn, bnnd, st_b = _pm.patchmatch(nS_pm, nC_pm, bnn, p_ba)
        else:
            nC_chw = np.ascontiguousarray(nC.transpose(2, 0, 1))
            nS_chw = np.ascontiguousarray(nS.transpose(2, 0, 1))
            ann, annd = _pm.patchmatch_ref_serial(nC_chw, nS_chw, ann, p_ab)
            bnn, bnnd = _pm.patchmatch_ref_serial(nS_chw, nC_chw, bnn, p_ba)
            st_a = st_b = (0, 0)
        tick("patchmatch", t0)
        t0 = time.perf_counter()
        bds_f = float(np.float32(cfg["bds"]))
        sml = _pm.reconstruct_bds(cnt_imgs[l], stl_imgs[l], ann, bnn, 1.0, bds_f)
        err = _pm.bds_feature_error(nC, featS[l], ann, bnn, 1.0, bds_f, mode=0)
        tick("bds", t0)
        t0 = time.perf_counter()
        cnt_lab = color.bgr2lab_u8(cnt_imgs[l])
        knn_id, knn_w = _pm.find_knns(labels, dc[0][2], dc[0][1], cnt_lab, 1 << l, cfg["clusters"])
        tick("knn", t0)
        t0 = time.perf_counter()
        stl_lab = color.bgr2lab_u8(sml)
        a0, b0 = color.local_fit(cnt_lab, stl_lab, cfg["eps"])
        weight = color.confidence_weights(err.reshape(ah, aw))
        norm_factor = float(cw * ch) / float(aw * ah)
        lam = cfg["w"] * norm_factor
        if cg_mode == "canonical":
            d2, wx2, wy2, kw2 = canonical_cg_weights(weight, cnt_lab, knn_id, knn_w, cfg["l"], cfg["alpha"], cfg["nl"], cfg["knum"], norm_factor)
            a1, b1, its = _pm.solve_nonlocal_canon(a0, b0, cnt_lab, stl_lab, d2, wx2, wy2, knn_id, kw2, 50 if l == 4 else 100)
        else:
            a1, b1, its = color.solve_nonlocal(a0, b0, weight, cnt_lab * (1.0 / 255.0), stl_lab * (1.0 / 255.0), knn_id, knn_w, l,
                                               cfg["l"], cfg["alpha"], cfg["nl"], cfg["knum"], norm_factor)
        tick("nonlocal", t0)
        t0 = time.perf_counter()
        a2, b2, rough = color.upsample_coefficients(a1, b1, cnt_lab_full_d, cw, ch)
        if ah == ch and aw == cw:
            lam = lam * 4
        a3, b3 = color.solve_wls(a2, b2, rough, cnt_lab_full_d[..., 0], lam, cfg["alpha"])
        result = color.apply_coefficients(cnt_lab_full_d, a3, b3)
        tick("wls", t0)
        if on_level is not None:
            on_level(l, dict(ann=ann, annd=annd, bnn=bnn, bnnd=bnnd, sml=sml, err=err, knn_id=knn_id, knn_w=knn_w, a0=a0, b0=b0,
                             weight=weight, a1=a1, b1=b1, a2=a2, b2=b2, rough=rough, a3=a3, b3=b3, result=result, cg_iters=its,
                             lam=lam, evals=(st_a, st_b), labels=labels, nC=nC, nS=nS, cnt_lab=cnt_lab, stl_lab=stl_lab))
        if result_hook is not None:  # lock-step tests: continue from the implementation-under-test's image
            result = result_hook(l, result)
        if l >= stop_after_level:
            break
        if l < 4:
            t0 = time.perf_counter()
            new = features_fn(result, l + 1)
            for k in range(l + 1, 5):
                featC[k] = new[k]
            tick("vgg", t0)
    return result


def canonical_cg_weights(weight, cnt_lab_u8, knn_id, knn_w, local_weight, alpha, nonlocal_weight, knum, d_weight):
    """The squared constraint weights of solve_nonlocal_downsample_gpu_gradient (CT/ColorTransfer.cpp:548-911) in the
    form cg_oracle.c consumes: d2 = (sqrt(weight) * sqrt(dWeight))^2 per pixel; wx2 / wy2 = 2 g^2 for the edge to x+1 /
    y+1 (each 4-neighbour edge is listed from both ends); kw2 = (sqrt(NN.w) * sqrt(nl / k))^2 per link.  Every
    operation is a separately rounded IEEE double operation (float parameters as in the reference's signature)."""
    h, w = weight.shape
    lam = float(np.float32(local_weight))
    gx, gy = color.gradient_weights(cnt_lab_u8[..., 0].astype(np.float64) * (1.0 / 255.0), lam, float(np.float32(alpha)))
    gx2, gy2 = gx * gx, gy * gy
    wx2 = (gx2 + gx2).ravel()
    wy2 = (gy2 + gy2).ravel()
    dw = np.sqrt(weight.ravel()) * float(np.sqrt(np.float32(d_weight)))
    d2 = dw * dw
    iw = np.sqrt(np.asarray(knn_w, np.float64)) * np.sqrt(nonlocal_weight / float(knum))
    kw2 = np.where(np.asarray(knn_id) >= 0, iw * iw, 0.0)
    return d2, wx2, wy2, np.ascontiguousarray(kw2)


def psnr(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = float((d * d).mean())
    return float("inf") if mse == 0 else 10.0 * np.log10(255.0 * 255.0 / mse)
