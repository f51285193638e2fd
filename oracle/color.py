"""oracle/color.py -- CPU restatement of the reference's colour stage (numpy / scipy / cv2).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Shorthand: CT/ = code/windows/neural_color_transfer/source/ColorTransfer/, NCT/ = .../source/

Follows
  image pyramid                NCT/main.cu:104-108
  Lab conversions / convertTo  NCT/main.cu:351-375, CT/ColorTransfer.h:54-60, CT/ColorTransfer.cpp:1467-1469
  local gain/bias fit          CT/ColorTransfer.cpp:425-455 (tables), 1194-1265
  confidence weights           CT/ColorTransfer.cpp:1302-1340
  non-local least squares      CT/ColorTransfer.cpp:519-546, 548-949 + CG of CT/SparseSolver_GPU.cu:119-159
  upsample + roughness         CT/ColorTransfer.cpp:457-490, 1376-1380
  WLS                          CT/ColorTransfer.cpp:492-517, 951-1125 + exact SPD solve (CT/SparseSolver_CPU.cpp:104-286)
  apply                        CT/ColorTransfer.cpp:1436-1469

Third-party arithmetic ("parity unpinned" by the reference's own tests):
  * OpenCV 2.4.10 cvtColor / resize / convertTo -> cv2 4.13 with cv2.setUseOptimized(False), i.e. OpenCV's plain
    C++ paths (the IPP/AVX paths of the wheel use higher-precision bilinear coefficients for 64F images; the plain
    path computes them in float like 2.4.10 does).
  * MKL PARDISO (exact SPD solve)           -> scipy.sparse.linalg.splu, FP64.
  * cuSPARSE/cuBLAS CG                      -> scipy CSR products + numpy dots, same recurrences, FP64.

Spec decisions
  C1  m_knnid entries with id == -1 (padding when fewer than k neighbours were found, CT/ColorTransfer.cpp:107-108)
      have weight 0 and produce an out-of-range column in the reference; here they are dropped (zero rows).
  C2  sqrt(dWeight) is evaluated in float (dWeight is a float parameter, CT/ColorTransfer.cpp:550,621).
  C3  alpha of the non-local gradient weights is (double)(float)1.2 (float parameter, :550); WLS uses 1.2 (double).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _cv2():
    import cv2

    cv2.setUseOptimized(False)
    try:
        cv2.ocl.setUseOpenCL(False)
    except Exception:
        pass
    return cv2


# ------------------------------------------------------------------ OpenCV stand-ins
def bgr2lab_u8(img):
    cv2 = _cv2()
    return cv2.cvtColor(np.ascontiguousarray(img, np.uint8), cv2.COLOR_BGR2Lab)


def lab2bgr_u8(img):
    cv2 = _cv2()
    return cv2.cvtColor(np.ascontiguousarray(img, np.uint8), cv2.COLOR_Lab2BGR)


def resize_linear(img, w, h):
    """cv::resize(src, dst, Size(w, h), 0, 0, INTER_LINEAR) for 8UC3 or 64FC3."""
    cv2 = _cv2()
    return cv2.resize(np.ascontiguousarray(img), (int(w), int(h)), interpolation=cv2.INTER_LINEAR)


def to_u8_x255(x):
    """Mat::convertTo(CV_8U, 255.0): saturate_cast<uchar>(cvRound(x*255)) (round half to even)."""
    return np.clip(np.rint(np.asarray(x, np.float64) * 255.0), 0, 255).astype(np.uint8)


def pyramid(img, sizes_hw):
    """NCT/main.cu:104-108: successive INTER_LINEAR resizes from the finest image; sizes_hw[l] for
    l = 0 (coarsest) .. L-1 (finest == img)."""
    L = len(sizes_hw)
    out = [None] * L
    out[L - 1] = np.ascontiguousarray(img, np.uint8)
    for l in range(L - 2, -1, -1):
        h, w = sizes_hw[l]
        out[l] = resize_linear(out[l + 1], w, h)
    return out


# ------------------------------------------------------------------ local fit
def _window_sums(img_i64):
    """Sums over the clipped 3x3 window [x-1, x+2) x [y-1, y+2) (what getValue() of the raster-order prefix tables
    returns, CT/ColorTransfer.cpp:46-58, 1201-1230) and the window pixel count."""
    h, w, _ = img_i64.shape
    p = np.zeros((h + 2, w + 2, 3), np.int64)
    p[1:-1, 1:-1] = img_i64
    s = np.zeros((h, w, 3), np.int64)
    for dy in range(3):
        for dx in range(3):
            s += p[dy:dy + h, dx:dx + w]
    ys = np.minimum(np.arange(h) + 2, h) - np.maximum(np.arange(h) - 1, 0)
    xs = np.minimum(np.arange(w) + 2, w) - np.maximum(np.arange(w) - 1, 0)
    cnt = (ys[:, None] * xs[None, :]).astype(np.float64)
    return s, cnt


def local_fit(cnt_lab_u8, stl_lab_u8, eps=0.6):
    """Per-pixel gain/bias from 3x3 mean / std (CT/ColorTransfer.cpp:1194-1265). Returns a, b (h, w, 3) float64."""
    c = cnt_lab_u8.astype(np.int64)
    s = stl_lab_u8.astype(np.int64)
    c1, n = _window_sums(c)
    c2, _ = _window_sums(c * c)
    s1, _ = _window_sums(s)
    s2, _ = _window_sums(s * s)
    n = n[..., None]
    cm = c1 / n
    cv = np.sqrt(np.maximum(c2 / n - cm * cm, 0.0))
    sm = s1 / n
    sv = np.sqrt(np.maximum(s2 / n - sm * sm, 0.0))
    a = sv / (cv + eps)
    b = (sm - cm * a) * (1.0 / 255.0)
    return a, b


def confidence_weights(err):
    """m_weight = max(1 - (err - min)/(max - min), 1e-6), CT/ColorTransfer.cpp:1302-1340. err float32 (h, w)."""
    e = np.asarray(err, np.float32).astype(np.float64)
    lo, hi = e.min(), e.max()
    return np.maximum(1.0 - (e - lo) / (hi - lo), 1e-6)


def gradient_weights(L, lam, alpha):
    """compute_gradientMat, CT/ColorTransfer.cpp:519-546: g = sqrt(lam / (|dL|^alpha + 1e-4)); gx[y, x] couples
    (x, x+1), gy[y, x] couples (y, y+1); zero on the last column / row."""
    h, w = L.shape
    gx = np.zeros((h, w))
    gy = np.zeros((h, w))
    gx[:, :-1] = np.sqrt(lam / (pow_libm(np.abs(L[:, 1:] - L[:, :-1]), alpha) + 1e-4))
    gy[:-1, :] = np.sqrt(lam / (pow_libm(np.abs(L[1:, :] - L[:-1, :]), alpha) + 1e-4))
    return gx, gy


def pow_libm(x, alpha):
    """Element-wise pow through the C library (math.pow -> libm), not numpy's vector loops: numpy may dispatch
    float64 pow to SIMD code whose last bit differs between CPUs, and the colour least squares amplifies a 1-ulp
    difference in these weights to ~43 dB in the final image (tests/test_oracle_pipeline.py).  L is 8-bit / 255, so
    there are at most 65536 distinct arguments."""
    import math

    u, inv = np.unique(np.asarray(x, np.float64), return_inverse=True)
    pu = np.array([math.pow(float(v), float(alpha)) for v in u], np.float64)
    return pu[inv].reshape(np.shape(x))


# ------------------------------------------------------------------ non-local least squares
def assemble_nonlocal(weight, src, ref, knn_id, knn_w, local_weight=0.125, alpha=1.2, nonlocal_weight=2.0, knum=8,
                      d_weight=1.0):
    """The explicit constraint matrix of solve_nonlocal_downsample_gpu_gradient (CT/ColorTransfer.cpp:548-911).
    weight (h, w); src / ref (h, w, 3) Lab/255 doubles; knn_id (n, k) int (-1 = padding), knn_w (n, k) doubles (= NN.w).
    Returns (A0, A1, A2) CSR matrices [rows x 2n] and (B0, B1, B2) right-hand sides, rows in the reference's order."""
    h, w = weight.shape
    n = h * w
    alpha_f = float(np.float32(alpha))                         # C3
    gx, gy = gradient_weights(src[..., 0], float(np.float32(local_weight)), alpha_f)
    sqrt_dw = float(np.sqrt(np.float32(d_weight)))             # C2
    dw = np.sqrt(weight.ravel()) * sqrt_dw
    idx = np.arange(n)
    xs, ys = idx % w, idx // w

    rows, cols, vals = [], [], [[], [], []]
    rhs = [[], [], []]
    r = 0
    # data term: one row per pixel
    rr = np.arange(n)
    rows += [rr, rr]
    cols += [idx, idx + n]
    for c in range(3):
        vals[c] += [dw * src[..., c].ravel(), dw]
        rhs[c].append(dw * ref[..., c].ravel())
    r += n
    # local smoothness: per pixel, in the order x+1, x-1, y+1, y-1, each for a then b (:663-846)
    has = [xs + 1 < w, xs - 1 >= 0, ys + 1 < h, ys - 1 >= 0]
    per_pix = 2 * (has[0].astype(np.int64) + has[1] + has[2] + has[3])
    base = r + np.concatenate([[0], np.cumsum(per_pix)[:-1]])
    off = np.zeros(n, np.int64)
    gxr, gyr = gx.ravel(), gy.ravel()
    specs = [
        (has[0], gxr, idx, idx + 1),                         # -g at i, +g at i+1, g = gradX(y, x)
        (has[1], np.concatenate([[0.0], gxr[:-1]]), idx - 1, idx),  # g = gradX(y, x-1): -g at i-1, +g at i
        (has[2], gyr, idx, idx + w),
        (has[3], np.concatenate([np.zeros(w), gyr[:-w]]), idx - w, idx),
    ]
    for m, g, i0, i1 in specs:
        sel = np.nonzero(m)[0]
        for blk in (0, n):                                   # a rows, then b rows
            rws = base[sel] + off[sel]
            rows += [rws, rws]
            cols += [i0[sel] + blk, i1[sel] + blk]
            for c in range(3):
                vals[c] += [-g[sel], g[sel]]
            off[sel] += 1
    n_local = int(per_pix.sum())
    for c in range(3):
        rhs[c].append(np.zeros(n_local))
    r += n_local
    # non-local term (:849-911)
    nlw = np.sqrt(nonlocal_weight / float(knum))
    kid = np.asarray(knn_id, np.int64)
    kw = np.asarray(knn_w, np.float64)
    valid = kid >= 0                                          # C1
    ci, ki = np.nonzero(valid)
    i1 = kid[ci, ki]
    iw = np.sqrt(kw[ci, ki]) * nlw
    m = len(ci)
    ra = r + 2 * np.arange(m)
    lo, hi = np.minimum(ci, i1), np.maximum(ci, i1)
    rows += [ra, ra, ra + 1, ra + 1]
    cols += [lo, hi, lo + n, hi + n]
    for c in range(3):
        vals[c] += [iw, -iw, iw, -iw]
        rhs[c].append(np.zeros(2 * m))
    r += 2 * m
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    A = [sp.csr_matrix((np.concatenate(vals[c]), (rows, cols)), shape=(r, 2 * n)) for c in range(3)]
    B = [np.concatenate(rhs[c]) for c in range(3)]
    return A, B


def cg_normal_equations(A, b, x0, tol=1e-6, maxit=100):
    """solve_ls_cg_gpu (CT/SparseSolver_GPU.cu:3-198): A^T A x = A^T b by un-preconditioned CG from x0,
    `while (r1 > tol*tol && k <= maxit)`."""
    AtA = (A.T @ A).tocsr()
    r = A.T @ b
    x = np.array(x0, np.float64, copy=True)
    r = r - AtA @ x
    r1 = float(r @ r)
    r0 = 0.0
    k = 1
    p = None
    while r1 > tol * tol and k <= maxit:
        if k > 1:
            p = (r1 / r0) * p + r
        else:
            p = r.copy()
        Ap = AtA @ p
        va = r1 / float(p @ Ap)
        x = x + va * p
        r = r - va * Ap
        r0 = r1
        r1 = float(r @ r)
        k += 1
    return x, k - 1


def solve_nonlocal(a0, b0, weight, src, ref, knn_id, knn_w, layer, local_weight=0.125, alpha=1.2, nonlocal_weight=2.0,
                   knum=8, d_weight=1.0):
    """solve_nonlocal_downsample_gpu_gradient: returns refined (a, b) (h, w, 3) and the iteration counts."""
    h, w = weight.shape
    n = h * w
    A, B = assemble_nonlocal(weight, src, ref, knn_id, knn_w, local_weight, alpha, nonlocal_weight, knum, d_weight)
    maxit = 50 if layer == 4 else 100
    a = np.empty((h, w, 3))
    b = np.empty((h, w, 3))
    its = []
    for c in range(3):
        x0 = np.concatenate([a0[..., c].ravel(), b0[..., c].ravel()])
        x, k = cg_normal_equations(A[c], B[c], x0, 1e-6, maxit)
        a[..., c] = x[:n].reshape(h, w)
        b[..., c] = x[n:].reshape(h, w)
        its.append(k)
    return a, b, its


# ------------------------------------------------------------------ upsample / WLS / apply
def upsample_coefficients(a_lvl, b_lvl, cnt_lab_d, W, H):
    """upsample_color_coefficients_bilinear (CT/ColorTransfer.cpp:457-490): bilinear to full size and the roughness
    map, which only channel 2 decides (the loop over c overwrites)."""
    h, w, _ = a_lvl.shape
    if W > w or H > h:
        a = resize_linear(a_lvl, W, H)
        b = resize_linear(b_lvl, W, H)
    else:
        a, b = a_lvl.copy(), b_lvl.copy()
    nc = cnt_lab_d[..., 2] * a[..., 2] + b[..., 2]
    rough = np.where((nc < 0) | (nc > 1), 1e-6, 1.0)
    return a, b, rough


def wls_matrix(rough, L_full, lam, alpha=1.2):
    """(W + L_g) of solve_WLS_roughness_cpu (CT/ColorTransfer.cpp:951-1093), full symmetric CSR."""
    H, W = rough.shape
    n = H * W
    gx, gy = gradient_weights(L_full, lam, alpha)
    wx = (gx ** 2).ravel()   # pow(g, 2)
    wy = (gy ** 2).ravel()
    idx = np.arange(n)
    xs, ys = idx % W, idx // W
    diag = rough.ravel().copy()
    mx = xs + 1 < W
    my = ys + 1 < H
    diag[mx] += wx[mx]
    diag[idx[mx] + 1] += wx[mx]
    diag[my] += wy[my]
    diag[idx[my] + W] += wy[my]
    r = np.concatenate([idx, idx[mx], idx[mx] + 1, idx[my], idx[my] + W])
    c = np.concatenate([idx, idx[mx] + 1, idx[mx], idx[my] + W, idx[my]])
    v = np.concatenate([diag, -wx[mx], -wx[mx], -wy[my], -wy[my]])
    return sp.csc_matrix((v, (r, c)), shape=(n, n))


def solve_wls(a, b, rough, L_full, lam, alpha=1.2):
    """Exact solve of (W + L_g) x = W x0 for the six maps (PARDISO in the reference)."""
    H, W = rough.shape
    M = wls_matrix(rough, L_full, lam, alpha)
    lu = spla.splu(M, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    rhs = np.concatenate([a.reshape(-1, 3), b.reshape(-1, 3)], axis=1) * rough.reshape(-1, 1)
    x = lu.solve(rhs)
    # one step of iterative refinement (PARDISO does up to 2)
    x += lu.solve(rhs - M @ x)
    return x[:, :3].reshape(H, W, 3), x[:, 3:].reshape(H, W, 3)


def apply_coefficients(cnt_lab_d, a, b):
    """res = clamp(Lab*a + b, 0, 1) -> 8U (x255, round half even) -> Lab2BGR (CT/ColorTransfer.cpp:1436-1469)."""
    res = np.minimum(np.maximum(cnt_lab_d * a + b, 0.0), 1.0)
    return lab2bgr_u8(to_u8_x255(res))


def transfer_color_level(err, down_cnt, sml_res, cnt_full_lab_d, knn_id, knn_w, layer, cfg=None, return_all=False):
    """ColorTransfer::transfer_color_downsample (CT/ColorTransfer.cpp:1180-1478) for one level.
    err: float32 (h*w) BDS feature error; down_cnt / sml_res: uint8 BGR (h, w, 3) level-size content and
    BDS-reconstructed style; cnt_full_lab_d: (H, W, 3) m_cntLabD. Returns the refined full-resolution BGR image."""
    cfg = dict(eps=0.6, nl=2.0, l=0.125, w=0.024, knum=8, alpha=1.2) | (cfg or {})
    h, w, _ = down_cnt.shape
    H, W, _ = cnt_full_lab_d.shape
    cnt_lab = bgr2lab_u8(down_cnt)
    stl_lab = bgr2lab_u8(sml_res)
    cnt_lab_d = cnt_lab.astype(np.float64) * (1.0 / 255.0)
    stl_lab_d = stl_lab.astype(np.float64) * (1.0 / 255.0)
    a0, b0 = local_fit(cnt_lab, stl_lab, cfg["eps"])
    weight = confidence_weights(np.asarray(err, np.float32).reshape(h, w))
    norm_factor = float(W * H) / float(w * h)
    lam = cfg["w"] * norm_factor
    a1, b1, its = solve_nonlocal(a0, b0, weight, cnt_lab_d, stl_lab_d, knn_id, knn_w, layer, cfg["l"], cfg["alpha"],
                                 cfg["nl"], cfg["knum"], norm_factor)
    a2, b2, rough = upsample_coefficients(a1, b1, cnt_full_lab_d, W, H)
    if h == H and w == W:
        lam = lam * 4
    a3, b3 = solve_wls(a2, b2, rough, cnt_full_lab_d[..., 0], lam, cfg["alpha"])
    out = apply_coefficients(cnt_full_lab_d, a3, b3)
    if return_all:
        return dict(out=out, a0=a0, b0=b0, weight=weight, a1=a1, b1=b1, a2=a2, b2=b2, rough=rough, a3=a3, b3=b3,
                    cg_iters=its, lam=lam)
    return out
