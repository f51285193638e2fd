"""CPU tests of the colour-stage oracle (oracle/color.py): internal consistency of the restated algebra."""
import numpy as np
import pytest

from oracle import color, synth


def rand_knn(rng, n, k=8, frac_invalid=0.05):
    ids = rng.integers(0, n, (n, k)).astype(np.int32)
    ids[ids == np.arange(n)[:, None]] = 0
    w = np.exp(1.0 - rng.random((n, k)) * 0.3 / 3.0)
    inval = rng.random((n, k)) < frac_invalid
    ids[inval] = -1
    w[inval] = 0.0
    return ids, w


def test_local_fit_matches_brute_force():
    cnt, stl = synth.pair(0, 13, 17)
    cl, sl = color.bgr2lab_u8(cnt), color.bgr2lab_u8(stl)
    a, b = color.local_fit(cl, sl, 0.6)
    for (x, y) in [(0, 0), (16, 12), (5, 7), (0, 6), (16, 0)]:
        wc = cl[max(y - 1, 0):y + 2, max(x - 1, 0):x + 2].reshape(-1, 3).astype(np.float64)
        ws = sl[max(y - 1, 0):y + 2, max(x - 1, 0):x + 2].reshape(-1, 3).astype(np.float64)
        cm, sm = wc.mean(0), ws.mean(0)
        cv = np.sqrt(np.maximum((wc ** 2).mean(0) - cm ** 2, 0))
        sv = np.sqrt(np.maximum((ws ** 2).mean(0) - sm ** 2, 0))
        ea = sv / (cv + 0.6)
        assert np.allclose(a[y, x], ea, rtol=1e-12)
        assert np.allclose(b[y, x], (sm - cm * ea) / 255.0, rtol=1e-12, atol=1e-15)


def test_identical_images_give_near_identity_fit():
    cnt, _ = synth.pair(0, 20, 20)
    cl = color.bgr2lab_u8(cnt)
    a, b = color.local_fit(cl, cl, 0.6)
    assert np.all(a <= 1.0) and np.all(a >= 0.0)


def test_confidence_weights_range():
    e = np.random.default_rng(0).standard_normal(100).astype(np.float32)
    w = color.confidence_weights(e.reshape(10, 10))
    assert w.max() == 1.0 and w.min() == 1e-6


def test_nonlocal_rows_and_normal_matrix_structure():
    rng = np.random.default_rng(1)
    h, w = 7, 9
    n = h * w
    ids, kw = rand_knn(rng, n)
    weight = rng.random((h, w)) + 0.1
    src, ref = rng.random((h, w, 3)), rng.random((h, w, 3))
    A, B = color.assemble_nonlocal(weight, src, ref, ids, kw, d_weight=4.0)
    n_edges_x, n_edges_y = h * (w - 1), (h - 1) * w
    # size + localConstraints (each edge listed from both ends, for a and b) + 2 per valid non-local link
    assert A[0].shape == (n + 4 * (n_edges_x + n_edges_y) + 2 * int((ids >= 0).sum()), 2 * n)
    N = (A[0].T @ A[0]).toarray()
    assert np.allclose(N, N.T)
    # smoothness rows annihilate constants: N @ [1;0] only keeps the data term
    d2 = weight.ravel() * float(np.float32(4.0))
    v = N @ np.concatenate([np.ones(n), np.zeros(n)])
    assert np.allclose(v[:n], d2 * src[..., 0].ravel() ** 2, rtol=1e-6)
    assert np.allclose(v[n:], d2 * src[..., 0].ravel(), rtol=1e-6)


def test_cg_converges_to_least_squares_solution_when_given_budget():
    rng = np.random.default_rng(2)
    h, w = 6, 6
    n = h * w
    ids, kw = rand_knn(rng, n, frac_invalid=0.0)
    weight = rng.random((h, w)) + 0.5
    src, ref = rng.random((h, w, 3)), rng.random((h, w, 3))
    A, B = color.assemble_nonlocal(weight, src, ref, ids, kw)
    x, k = color.cg_normal_equations(A[1], B[1], np.zeros(2 * n), 1e-12, 5000)
    ls = np.linalg.lstsq(A[1].toarray(), B[1], rcond=None)[0]
    assert np.allclose(x, ls, atol=1e-6)
    x100, k100 = color.cg_normal_equations(A[1], B[1], np.zeros(2 * n), 1e-6, 100)
    assert k100 <= 100


def test_wls_matches_dense_solve_and_preserves_constants():
    rng = np.random.default_rng(3)
    H, W = 9, 8
    L = rng.random((H, W))
    rough = np.where(rng.random((H, W)) < 0.2, 1e-6, 1.0)
    a = rng.random((H, W, 3))
    b = rng.random((H, W, 3))
    M = color.wls_matrix(rough, L, 0.5).toarray()
    assert np.allclose(M, M.T) and np.all(np.linalg.eigvalsh(M) > 0)
    a2, b2 = color.solve_wls(a, b, rough, L, 0.5)
    ea = np.linalg.solve(M, a.reshape(-1, 3) * rough.reshape(-1, 1)).reshape(H, W, 3)
    assert np.allclose(a2, ea, atol=1e-10)
    # a constant map is a fixed point: (W + L) c = W c
    c = np.full((H, W, 3), 0.7)
    ca, _ = color.solve_wls(c, c, rough, L, 0.5)
    assert np.allclose(ca, 0.7, atol=1e-9)


def test_transfer_color_level_runs_end_to_end_small():
    rng = np.random.default_rng(4)
    H = W = 24
    h = w = 12
    cnt, stl = synth.pair(0, H, W)
    down_cnt = color.resize_linear(cnt, w, h)
    sml = color.resize_linear(stl, w, h)
    ids, kw = rand_knn(rng, h * w)
    err = rng.random(h * w).astype(np.float32) - 1.0
    cnt_lab_d = color.bgr2lab_u8(cnt).astype(np.float64) / 255.0
    r = color.transfer_color_level(err, down_cnt, sml, cnt_lab_d, ids, kw, layer=3, return_all=True)
    assert r["out"].shape == (H, W, 3) and r["out"].dtype == np.uint8
    assert max(r["cg_iters"]) <= 100 and np.isfinite(r["a3"]).all()


def test_unconverged_cg_iterate_is_rounding_sensitive_and_canonical_order_is_close():
    """Evidence for decision N1 (oracle/cg_oracle.c): (i) perturbing A^T A by 1e-15 relative moves the 100-iteration
    iterate by far more than 1e-15 -- the reference's own result is only defined up to that; (ii) the canonical-order
    matrix-free CG agrees with the reference-order CG to within that same sensitivity band."""
    import oracle

    rng = np.random.default_rng(44)
    h = w = 32
    n = h * w
    cnt, stl = synth.pair(2, h, w)
    cl, sl = color.bgr2lab_u8(cnt), color.bgr2lab_u8(stl)
    ids, kw = rand_knn(rng, n, frac_invalid=0.0)
    weight = np.maximum(rng.random((h, w)), 1e-6)
    a0, b0 = color.local_fit(cl, sl, 0.6)
    A, B = color.assemble_nonlocal(weight, cl / 255.0, sl / 255.0, ids, kw, d_weight=253.0)
    x0 = np.concatenate([a0[..., 0].ravel(), b0[..., 0].ravel()])
    x_ref, _ = color.cg_normal_equations(A[0], B[0], x0, 1e-6, 100)
    Ap = A[0].copy()
    Ap.data = Ap.data * (1.0 + 1e-15 * rng.standard_normal(Ap.data.shape))
    x_per, _ = color.cg_normal_equations(Ap, B[0], x0, 1e-6, 100)
    sens = np.abs(x_per - x_ref).max() / np.abs(x_ref).max()
    assert sens > 1e-9  # amplification by >= 6 orders of magnitude
    gx, gy = color.gradient_weights(cl[..., 0] / 255.0, 0.125, float(np.float32(1.2)))
    dw = np.sqrt(weight.ravel()) * float(np.sqrt(np.float32(253.0)))
    iw = np.sqrt(kw) * np.sqrt(2.0 / 8)
    ca, cb, its = oracle.solve_nonlocal_canon(a0, b0, cl, sl, dw * dw, 2 * (gx * gx).ravel(), 2 * (gy * gy).ravel(), ids, iw * iw, 100)
    xc = np.concatenate([ca[..., 0].ravel(), cb[..., 0].ravel()])
    diff = np.abs(xc - x_ref).max() / np.abs(x_ref).max()
    assert its == [100, 100, 100] and diff < max(50 * sens, 5e-3)
    # run to convergence and the two agree tightly: the disagreement above is the truncation, not the operator
    x_conv, _ = color.cg_normal_equations(A[0], B[0], x0, 1e-11, 20000)
    cca, ccb, _ = oracle.solve_nonlocal_canon(a0, b0, cl, sl, dw * dw, 2 * (gx * gx).ravel(), 2 * (gy * gy).ravel(), ids, iw * iw, 20000, 1e-11)
    assert np.abs(np.concatenate([cca[..., 0].ravel(), ccb[..., 0].ravel()]) - x_conv).max() / np.abs(x_conv).max() < 1e-6
