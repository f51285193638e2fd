"""ctypes wrappers over oracle/liboracle_pm.so (built by oracle/Makefile from pm_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle_pm.so")
_lib = None

_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_up = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False):
    """Compile the C restatement (gcc). Building the checker is not using it."""
    srcs = [os.path.join(_HERE, f) for f in ("pm_oracle.c", "bds_oracle.c", "cg_oracle.c", "cluster_oracle.c", "conv_oracle.c")]
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle_pm.so"], stdout=subprocess.DEVNULL)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_xorwow_uniform_table.argtypes = [C.c_int, C.c_int, _fp]
        L.orc_xorwow_raw.argtypes = [C.c_ulonglong, C.c_int, _up]
        L.orc_nnf_init.argtypes = [C.c_int] * 4 + [_up]
        L.orc_nnf_upsample.argtypes = [_up] + [C.c_int] * 6 + [_up]
        L.orc_dist_canon.argtypes = [_fp, _fp] + [C.c_int] * 9
        L.orc_dist_canon.restype = C.c_float
        L.orc_dist_ref_chw.argtypes = [_fp, _fp] + [C.c_int] * 10 + [C.c_float, C.c_int]
        L.orc_dist_ref_chw.restype = C.c_float
        L.orc_l2norm_hwc.argtypes = [_fp, _fp, C.c_int, C.c_int]
        L.orc_chw_to_hwc.argtypes = [_fp, _fp, C.c_int, C.c_int]
        L.orc_patchmatch.argtypes = [_fp, _fp, _up, _fp, _ip, _lp]
        L.orc_patchmatch.restype = C.c_int
        L.orc_patchmatch_opts.argtypes = [_fp, _fp, _up, _fp, _ip, _lp, C.c_int]
        L.orc_patchmatch_opts.restype = C.c_int
        L.orc_patchmatch_ref_serial.argtypes = [_fp, _fp, _up, _fp, _ip]
        L.orc_patchmatch_ref_serial.restype = C.c_int
        L.orc_num_threads.restype = C.c_int
        _bp = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
        L.orc_reconstruct_bds.argtypes = [_bp, _bp, _up, _up] + [C.c_int] * 4 + [C.c_double, C.c_double, _bp]
        L.orc_bds_feature_error.argtypes = [_fp, _fp, _up, _up] + [C.c_int] * 5 + [C.c_float, C.c_float, C.c_int, _fp, C.c_void_p]
        _dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        L.orc_solve_nonlocal_canon.argtypes = [_dp, _dp, _bp, _bp, _dp, _dp, _dp, _ip, _dp, C.c_int, C.c_int, C.c_int, C.c_double, _ip]
        L.orc_msvc_shuffle.argtypes = [C.c_int, _ip]
        L.orc_kmeans_labels.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, _ip, C.c_void_p]
        L.orc_kmeans_labels.restype = C.c_int
        for fn in (L.orc_find_knns, L.orc_find_knns_brute):
            fn.argtypes = [_ip, C.c_int, C.c_int, C.c_int, _bp, C.c_int, C.c_int, C.c_int, _ip, _dp]
        L.orc_conv3x3_relu_canon.argtypes = [_fp, _fp, _fp, _fp] + [C.c_int] * 4
        L.orc_maxpool2x2_ceil.argtypes = [_fp, _fp] + [C.c_int] * 3
        L.orc_preprocess_bgr.argtypes = [_bp, _fp, C.c_int]
        _lib = L
    return _lib


def make_params(Cn, ah, aw, bh, bw, iters=10, rs_max=32, patch=3):
    """NCT/main.cu:204-214"""
    return np.array([Cn, ah, aw, bh, bw, patch, iters, rs_max, 0, 10, 1], dtype=np.int32)


def xorwow_uniform_table(ncols, ndraws):
    out = np.empty((ncols, ndraws), np.float32)
    lib().orc_xorwow_uniform_table(ncols, ndraws, out)
    return out


def xorwow_raw(seed, ndraws):
    out = np.empty(ndraws, np.uint32)
    lib().orc_xorwow_raw(seed, ndraws, out)
    return out


def nnf_init(ah, aw, bh, bw):
    ann = np.empty(ah * aw, np.uint32)
    lib().orc_nnf_init(ah, aw, bh, bw, ann)
    return ann


def nnf_upsample(ann_half, ah_half, aw_half, ah, aw, bh, bw):
    ann = np.empty(ah * aw, np.uint32)
    lib().orc_nnf_upsample(np.ascontiguousarray(ann_half, np.uint32), ah_half, aw_half, ah, aw, bh, bw, ann)
    return ann


def l2norm_hwc(x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    lib().orc_l2norm_hwc(x, out, x.shape[0] * x.shape[1], x.shape[2])
    return out


def chw_to_hwc(x):
    x = np.ascontiguousarray(x, np.float32)
    Cn, H, W = x.shape
    out = np.empty((H, W, Cn), np.float32)
    lib().orc_chw_to_hwc(x, out, Cn, H * W)
    return out


def dist_canon(a, b, ax, ay, bx, by):
    ah, aw, Cn = a.shape
    bh, bw, _ = b.shape
    return float(lib().orc_dist_canon(a, b, Cn, ah, aw, bh, bw, ax, ay, bx, by))


def dist_ref_chw(a_chw, b_chw, ax, ay, bx, by, cutoff=float(2**31), use_fma=1):
    Cn, ah, aw = a_chw.shape
    _, bh, bw = b_chw.shape
    return float(lib().orc_dist_ref_chw(a_chw, b_chw, Cn, ah, aw, bh, bw, ax, ay, bx, by, 3, cutoff, use_fma))


def patchmatch(a_hwc, b_hwc, ann, params, d4=True):
    """Deterministic PatchMatch (decisions D1-D4). Returns (ann, annd, (evals_ref, evals_gpu, evals_dedup)):
    evals_ref = evaluations of the reference semantics, evals_gpu = left after D3 de-duplication + D4 unchanged-source
    skip (what the GPU kernel computes), evals_dedup = left after D3 alone.  d4=False switches D4 off (same field)."""
    a = np.ascontiguousarray(a_hwc, np.float32)
    b = np.ascontiguousarray(b_hwc, np.float32)
    ann = np.array(ann, dtype=np.uint32, copy=True).ravel()
    annd = np.empty(ann.shape[0], np.float32)
    stats = np.zeros(3, np.int64)
    rc = lib().orc_patchmatch_opts(a, b, ann, annd, np.ascontiguousarray(params, np.int32), stats, 1 if d4 else 0)
    if rc != 0:
        raise ValueError(f"orc_patchmatch: unsupported parameters ({rc})")
    return ann, annd, (int(stats[0]), int(stats[2]), int(stats[1]))


def patchmatch_ref_serial(a_chw, b_chw, ann, params):
    """Reference-semantics in-place PatchMatch in the reference's layout / summation order."""
    a = np.ascontiguousarray(a_chw, np.float32)
    b = np.ascontiguousarray(b_chw, np.float32)
    ann = np.array(ann, dtype=np.uint32, copy=True).ravel()
    annd = np.empty(ann.shape[0], np.float32)
    lib().orc_patchmatch_ref_serial(a, b, ann, annd, np.ascontiguousarray(params, np.int32))
    return ann, annd


def num_threads():
    return int(lib().orc_num_threads())


def unpack(ann):
    ann = np.asarray(ann, np.uint32)
    return (ann & 0xFFF).astype(np.int32), ((ann >> 12) & 0xFFF).astype(np.int32)


def reconstruct_bds(a_img, b_img, ann, bnn, w_cohen=1.0, w_complete=2.0):
    """reconstruct_bds, NCT/GeneralizedPatchMatch.cu:122-235 (uint8 BGR HxWx3)."""
    a_img = np.ascontiguousarray(a_img, np.uint8)
    b_img = np.ascontiguousarray(b_img, np.uint8)
    ah, aw, _ = a_img.shape
    bh, bw, _ = b_img.shape
    out = np.empty((ah, aw, 3), np.uint8)
    lib().orc_reconstruct_bds(a_img, b_img, np.ascontiguousarray(ann, np.uint32), np.ascontiguousarray(bnn, np.uint32),
                              ah, aw, bh, bw, float(w_cohen), float(w_complete), out)
    return out


def bds_feature_error(c_norm, s_raw, ann, bnn, w_cohen=1.0, w_complete=2.0, mode=0, want_vote=False):
    """avg_vote_bds_a/_b/avg_vote_bds + norm + feature_distance (NCT/main.cu:297-318).
    mode 0 = canonical reduction order (GPU parity target), 1 = the reference's sequential order."""
    c_norm = np.ascontiguousarray(c_norm, np.float32)
    s_raw = np.ascontiguousarray(s_raw, np.float32)
    ah, aw, Cn = c_norm.shape
    bh, bw, _ = s_raw.shape
    err = np.empty(ah * aw, np.float32)
    vote = np.empty((ah, aw, Cn), np.float32) if want_vote else None
    lib().orc_bds_feature_error(c_norm, s_raw, np.ascontiguousarray(ann, np.uint32), np.ascontiguousarray(bnn, np.uint32),
                                Cn, ah, aw, bh, bw, float(w_cohen), float(w_complete), mode, err,
                                vote.ctypes.data if want_vote else None)
    return (err, vote) if want_vote else err


def solve_nonlocal_canon(a0, b0, src_u8, ref_u8, d2, wx2, wy2, knn_id, kw2, maxit, tol=1e-6):
    """Canonical-order matrix-free CG of cg_oracle.c (decision N1). a0/b0 (h, w, 3) float64; returns (a, b, iters)."""
    h, w, _ = a0.shape
    a = np.array(a0, np.float64, copy=True)
    b = np.array(b0, np.float64, copy=True)
    its = np.zeros(3, np.int32)
    lib().orc_solve_nonlocal_canon(a, b, np.ascontiguousarray(src_u8, np.uint8), np.ascontiguousarray(ref_u8, np.uint8),
                                   np.ascontiguousarray(d2, np.float64), np.ascontiguousarray(wx2, np.float64),
                                   np.ascontiguousarray(wy2, np.float64), np.ascontiguousarray(knn_id, np.int32),
                                   np.ascontiguousarray(kw2, np.float64), h, w, int(maxit), float(tol), its)
    return a, b, [int(v) for v in its]


def msvc_shuffle(n):
    out = np.empty(n, np.int32)
    lib().orc_msvc_shuffle(n, out)
    return out


def kmeans_labels(features, k=10, iterations=11):
    """Root split of cvflann's hierarchical k-means (clusterFeastures). features (n, dim) float32.
    Returns (labels int32[n], number of clusters)."""
    f = np.ascontiguousarray(features, np.float32)
    labels = np.zeros(f.shape[0], np.int32)
    nl = lib().orc_kmeans_labels(f, f.shape[0], f.shape[1], k, iterations, labels, None)
    return labels, int(nl)


def find_knns(labels, lw, lh, lab_u8, samples, nlabels=10, brute=False):
    """ColorTransfer::findKnns. Returns (knn_id int32 [n, 8], knn_w float64 [n, 8])."""
    lab = np.ascontiguousarray(lab_u8, np.uint8)
    h, w, _ = lab.shape
    ids = np.empty((h * w, 8), np.int32)
    wts = np.empty((h * w, 8), np.float64)
    fn = lib().orc_find_knns_brute if brute else lib().orc_find_knns
    fn(np.ascontiguousarray(labels, np.int32), lw, lh, nlabels, lab, h, w, samples, ids, wts)
    return ids, wts
