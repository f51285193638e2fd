"""GPU parity tests of the colour stage against oracle/color.py (cv2 plain C++ paths, scipy).
Bars: bit-exact for the 8-bit / integer work and the closed-form FP64 steps; 1e-4 relative (max-norm, per map)
for the iterative solves -- the tolerance BASELINE.json's north_star states for the colour least squares."""
import numpy as np
import pytest

from oracle import color, synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


def to_dev(x, dev):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    torch.cuda.synchronize()  # libnct contexts run on their own non-blocking stream: the copy must have landed
    return t


def relerr(x, ref):
    return float(np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-12))


def rand_knn(rng, n, k=8, frac_invalid=0.05):
    ids = rng.integers(0, n, (n, k)).astype(np.int32)
    ids[ids == np.arange(n)[:, None]] = 0
    w = np.exp(1.0 - rng.random((n, k)) * 0.3 / 3.0)
    inval = rng.random((n, k)) < frac_invalid
    ids[inval] = -1
    w[inval] = 0.0
    return ids, w


def test_bgr2lab_lab2bgr_exhaustive(ctx, dev):
    """all 2^24 8-bit colours, both directions, against cv2"""
    import torch

    v = torch.arange(1 << 24, dtype=torch.int32, device=dev)
    img = torch.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], dim=1).to(torch.uint8).contiguous().view(4096, 4096, 3)
    lab = ctx.bgr2lab(img)
    bgr = ctx.lab2bgr(img)
    ctx.synchronize()
    host = img.cpu().numpy()
    for s in range(0, 4096, 512):
        assert np.array_equal(lab[s:s + 512].cpu().numpy(), color.bgr2lab_u8(host[s:s + 512]))
        assert np.array_equal(bgr[s:s + 512].cpu().numpy(), color.lab2bgr_u8(host[s:s + 512]))


@pytest.mark.parametrize("sh,sw,dh,dw", [(700, 700, 350, 350), (350, 350, 175, 175), (175, 175, 88, 88), (88, 88, 44, 44),
                                         (125, 125, 63, 63), (520, 352, 260, 176), (65, 44, 33, 22), (1200, 900, 1000, 750),
                                         (37, 41, 90, 77)])
def test_resize_u8_bit_exact(ctx, dev, sh, sw, dh, dw):
    src = np.random.default_rng(sh + dw).integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    g = ctx.resize_linear(to_dev(src, dev), dh, dw)
    ctx.synchronize()
    assert np.array_equal(g.cpu().numpy(), color.resize_linear(src, dw, dh))


@pytest.mark.parametrize("sh,sw,dh,dw", [(44, 44, 700, 700), (88, 88, 700, 700), (175, 175, 700, 700), (350, 350, 700, 700),
                                         (63, 63, 1000, 1000), (65, 44, 520, 352), (38, 60, 600, 960)])
def test_resize_f64_bit_exact(ctx, dev, sh, sw, dh, dw):
    src = np.random.default_rng(sh).standard_normal((sh, sw, 3))
    g = ctx.resize_linear(to_dev(src, dev), dh, dw)
    ctx.synchronize()
    assert np.array_equal(g.cpu().numpy(), color.resize_linear(src, dw, dh))


def test_pyramid_matches_reference_chain(pkg, ctx, dev):
    cnt, _ = synth.pair(0, 700, 700)
    sizes = pkg.level_sizes(700)[::-1]
    ref = color.pyramid(cnt, [(s, s) for s in sizes])
    cur = to_dev(cnt, dev)
    for l in range(3, -1, -1):
        cur = ctx.resize_linear(cur, sizes[l], sizes[l])
        ctx.synchronize()
        assert np.array_equal(cur.cpu().numpy(), ref[l])


@pytest.mark.parametrize("h,w", [(44, 44), (13, 17), (175, 175), (1, 7), (700, 700)])
def test_local_fit_and_weights_bit_exact(ctx, dev, h, w):
    cnt, stl = synth.pair(1, h, w)
    cl, sl = color.bgr2lab_u8(cnt), color.bgr2lab_u8(stl)
    a, b = ctx.local_fit(to_dev(cl, dev), to_dev(sl, dev), 0.6)
    err = (np.random.default_rng(0).random(h * w).astype(np.float32) - 1.0)
    wg = ctx.confidence_weights(to_dev(err, dev))
    ctx.synchronize()
    oa, ob = color.local_fit(cl, sl, 0.6)
    assert np.array_equal(a.cpu().numpy(), oa) and np.array_equal(b.cpu().numpy(), ob)
    assert np.array_equal(wg.cpu().numpy(), color.confidence_weights(err.reshape(h, w)).ravel())


@pytest.mark.parametrize("h,w,layer,dwt", [(44, 44, 0, 253.0), (60, 52, 2, 16.0), (96, 96, 4, 1.0), (175, 175, 2, 16.0)])
def test_solve_nonlocal(ctx, dev, h, w, layer, dwt):
    """The un-converged CG iterate is chaotically sensitive to rounding (see oracle/cg_oracle.c), so parity is
    pinned bit-for-bit against the canonical-order oracle fed with the kernel's own weight arrays; those arrays and
    the result are additionally checked against the reference-order oracle (explicit A^T A, scipy)."""
    import oracle

    rng = np.random.default_rng(h)
    n = h * w
    cnt, stl = synth.pair(2, h, w)
    cl, sl = color.bgr2lab_u8(cnt), color.bgr2lab_u8(stl)
    ids, kw = rand_knn(rng, n)
    weight = np.maximum(rng.random((h, w)), 1e-6)
    a0, b0 = color.local_fit(cl, sl, 0.6)
    ga, gb = to_dev(a0, dev), to_dev(b0, dev)
    its = ctx.solve_nonlocal(ga, gb, to_dev(weight.ravel(), dev), to_dev(cl, dev), to_dev(sl, dev), to_dev(ids, dev),
                             to_dev(kw, dev), layer, d_weight=dwt, want_iters=True)
    d2, wx2, wy2 = (ctx.read_scratch(k, np.float64, n) for k in ("nl_d2", "nl_wx2", "nl_wy2"))
    kw2 = ctx.read_scratch("nl_kw2", np.float64, n * 8)
    # the kernel's operator coefficients equal the oracle's bit for bit (pow() comes from a host-libm table on both sides)
    from oracle import pipeline
    od2, owx2, owy2, okw2 = pipeline.canonical_cg_weights(weight, cl, ids, kw, 0.125, 1.2, 2.0, 8, dwt)
    assert np.array_equal(wx2, owx2) and np.array_equal(wy2, owy2)
    assert np.array_equal(d2, od2)
    assert np.array_equal(kw2.reshape(n, 8), okw2)
    # bit-exact against the canonical-order oracle
    maxit = 50 if layer == 4 else 100
    ca, cb, cits = oracle.solve_nonlocal_canon(a0, b0, cl, sl, d2, wx2, wy2, ids, kw2, maxit)
    assert its == cits
    assert np.array_equal(ga.cpu().numpy(), ca) and np.array_equal(gb.cpu().numpy(), cb)
    # against the reference-order oracle: same iteration count, agreement limited by the iterate's own sensitivity
    oa, ob, oits = color.solve_nonlocal(a0, b0, weight, cl / 255.0, sl / 255.0, ids, kw, layer, d_weight=dwt)
    assert its == oits
    worst = max(max(relerr(ga.cpu().numpy()[..., c], oa[..., c]), relerr(gb.cpu().numpy()[..., c], ob[..., c])) for c in range(3))
    print(f"non-local CG {h}x{w} layer {layer}: iters {its}, max rel. diff vs reference-order oracle {worst:.2e}")
    assert worst < 5e-3


def test_solve_ls_cg_with_the_reference_argument_list(pkg, ctx):
    """nct_solve_ls_cg = solve_ls_cg_gpu's own signature (CT/SparseSolver_GPU.cuh:12): host arrays, one-based CSR of
    the explicit constraint matrix the reference assembles (here: oracle assemble_nonlocal), x0 in / x out.  Against the
    scipy restatement of the same loop: same iteration count; a few iterations agree to rounding, the full 100 to the
    band the un-converged iterate is defined to."""
    rng = np.random.default_rng(7)
    h, w = 24, 20
    n = h * w
    cnt, stl = synth.pair(2, h, w)
    cl, sl = color.bgr2lab_u8(cnt), color.bgr2lab_u8(stl)
    ids, kw = rand_knn(rng, n)
    weight = np.maximum(rng.random((h, w)), 1e-6)
    a0, b0 = color.local_fit(cl, sl, 0.6)
    A, B = color.assemble_nonlocal(weight, cl * (1.0 / 255.0), sl * (1.0 / 255.0), ids, kw, d_weight=16.0)
    for c in range(3):
        M = A[c].tocsr()
        M.sort_indices()
        rows, size = M.shape
        x0 = np.concatenate([a0[..., c].ravel(), b0[..., c].ravel()])
        for maxit, tol in ((5, 1e-9), (100, 5e-3)):
            x = x0.copy()
            its = ctx.solve_ls_cg(size, rows, M.data, M.indices + 1, M.indptr + 1, x, B[c], 1e-6, maxit)
            ox, oits = color.cg_normal_equations(M, B[c], x0, 1e-6, maxit)
            assert its == oits == maxit
            assert relerr(x, ox) < tol, f"channel {c}, {maxit} iterations: {relerr(x, ox):.2e}"
    # a small well-conditioned system: stops on the tolerance, at the least-squares solution
    import scipy.sparse as sp

    M = sp.vstack([sp.identity(30), sp.random(90, 30, density=0.1, random_state=3)]).tocsr()
    M.sort_indices()
    bb = rng.standard_normal(120)
    x = np.zeros(30)
    its = ctx.solve_ls_cg(30, 120, M.data, M.indices + 1, M.indptr + 1, x, bb, 1e-8, 100)
    ox, oits = color.cg_normal_equations(M, bb, np.zeros(30), 1e-8, 100)
    assert 0 < its < 100 and abs(its - oits) <= 1
    assert relerr(x, np.linalg.lstsq(M.toarray(), bb, rcond=None)[0]) < 1e-7
    # zero iterations when the start vector already satisfies the tolerance
    x2 = x.copy()
    assert ctx.solve_ls_cg(30, 120, M.data, M.indices + 1, M.indptr + 1, x2, bb, 1e-3, 100) == 0 and np.array_equal(x2, x)
    # zero-based indices are rejected (the reference's arrays are one-based, CUSPARSE_INDEX_BASE_ONE)
    with pytest.raises(pkg.NctError):
        ctx.solve_ls_cg(30, 120, M.data, M.indices, M.indptr, x, bb, 1e-8, 10)


@pytest.mark.parametrize("h,w,H,W", [(44, 44, 700, 700), (30, 25, 120, 100), (64, 64, 64, 64)])
def test_upsample_roughness_apply_bit_exact(ctx, dev, h, w, H, W):
    rng = np.random.default_rng(H)
    cnt, _ = synth.pair(3, H, W)
    lab = color.bgr2lab_u8(cnt)
    a = 1.0 + 0.5 * rng.standard_normal((h, w, 3))
    b = 0.2 * rng.standard_normal((h, w, 3))
    ga, gb, gr = ctx.upsample_coefficients(to_dev(a, dev), to_dev(b, dev), to_dev(lab, dev))
    out, out_lab = ctx.apply_coefficients(to_dev(lab, dev), ga, gb, want_lab=True)
    ctx.synchronize()
    oa, ob, orr = color.upsample_coefficients(a, b, lab / 255.0, W, H)
    assert np.array_equal(ga.cpu().numpy(), oa) and np.array_equal(gb.cpu().numpy(), ob)
    assert np.array_equal(gr.cpu().numpy(), orr)
    assert (orr == 1e-6).any() and (orr == 1.0).any()
    res = np.minimum(np.maximum(lab / 255.0 * oa + ob, 0.0), 1.0)
    assert np.array_equal(out_lab.cpu().numpy(), color.to_u8_x255(res))
    assert np.array_equal(out.cpu().numpy(), color.apply_coefficients(lab / 255.0, oa, ob))


@pytest.mark.parametrize("H,W,lam", [(96, 80, 6.07), (128, 128, 0.096), (175, 160, 1.5)])
def test_solve_wls_matches_direct_solve(ctx, dev, H, W, lam):
    rng = np.random.default_rng(W)
    cnt, _ = synth.pair(4, H, W)
    lab = color.bgr2lab_u8(cnt)
    a = 1.0 + 0.5 * rng.standard_normal((H, W, 3))
    b = 0.2 * rng.standard_normal((H, W, 3))
    rough = np.where(rng.random((H, W)) < 0.1, 1e-6, 1.0)
    ga, gb = to_dev(a, dev), to_dev(b, dev)
    its, res = ctx.solve_wls(ga, gb, to_dev(rough, dev), to_dev(lab, dev), lam, 1.2)
    oa, ob = color.solve_wls(a, b, rough, lab[..., 0] / 255.0, lam, 1.2)
    assert res <= 1e-10 and its > 0
    for c in range(3):
        assert relerr(ga.cpu().numpy()[..., c], oa[..., c]) < REL_TOL
        assert relerr(gb.cpu().numpy()[..., c], ob[..., c]) < REL_TOL
    ja, jb = to_dev(a, dev), to_dev(b, dev)
    jits, jres = ctx.solve_wls(ja, jb, to_dev(rough, dev), to_dev(lab, dev), lam, 1.2, jacobi=True)
    assert relerr(ja.cpu().numpy(), ga.cpu().numpy()) < 1e-7  # multigrid and Jacobi PCG agree
    print(f"WLS {H}x{W} lam={lam}: {its} MG-PCG iterations (Jacobi-PCG: {jits}), rel.res {res:.2e}, "
          f"max rel.err a {max(relerr(ga.cpu().numpy()[..., c], oa[..., c]) for c in range(3)):.2e}")


@pytest.mark.parametrize("H,W,lam,hole", [(700, 700, 6.144, 0), (700, 700, 1.536, 0), (700, 700, 0.384, 0), (700, 700, 0.096, 0), (350, 280, 0.096, 0),
                                          (96, 80, 6.07, 0), (700, 700, 0.096, 380), (700, 700, 0.384, 560), (350, 280, 0.096, 150)])
def test_wls_adaptive_bottom_depth(ctx, dev, H, W, lam, hole, monkeypatch):
    """NCT_WLS_DEPTH = t: the single-block bottom of the V-cycle stops at the first level whose off-diagonal share has
    mean <= t and maximum <= 0.99 instead of descending to one node.  Same solution to the tolerance, iteration count
    within 2 of the full depth -- also with a contiguous hole x hole region of roughness 1e-6 (out-of-range colours),
    which needs the deep levels and must keep them (a mean-only criterion costs 8-15 iterations there)."""
    rng = np.random.default_rng(H + W)
    cnt, _ = synth.pair(6, H, W)
    lab = color.bgr2lab_u8(cnt)
    a = 1.0 + 0.5 * rng.standard_normal((H, W, 3))
    b = 0.2 * rng.standard_normal((H, W, 3))
    rough = np.where(rng.random((H, W)) < 0.1, 1e-6, 1.0)
    if hole:
        rough[H // 8:H // 8 + hole, W // 10:W // 10 + hole] = 1e-6
    got = {}
    for thr in ("0", "0.96", "0.96"):
        monkeypatch.setenv("NCT_WLS_DEPTH", thr)
        ga, gb = to_dev(a, dev), to_dev(b, dev)
        its, res = ctx.solve_wls(ga, gb, to_dev(rough, dev), to_dev(lab, dev), lam, 1.2, rel_tol=1e-8)
        assert res <= 1e-8
        r = (its, ga.cpu().numpy().copy(), gb.cpu().numpy().copy())
        if thr in got:   # deterministic: the depth decision is integer arithmetic
            assert r[0] == got[thr][0] and np.array_equal(r[1].view(np.uint64), got[thr][1].view(np.uint64))
        got[thr] = r
    print(f"WLS {H}x{W} lam={lam} hole={hole}: {got['0'][0]} iterations at full depth, {got['0.96'][0]} with the adaptive bottom")
    assert abs(got["0.96"][0] - got["0"][0]) <= 2
    assert relerr(got["0.96"][1], got["0"][1]) < 1e-6 and relerr(got["0.96"][2], got["0"][2]) < 1e-6


@pytest.mark.parametrize("H,W,lam", [(128, 128, 0.096), (350, 280, 1.5), (96, 80, 6.07)])
def test_solve_wls_loop_modes_are_bit_identical(ctx, dev, H, W, lam, monkeypatch):
    """The stopping test runs on the device after every iteration, so the solution and the iteration count do not depend on
    how the launches reach the GPU: device-side WHILE loop (CUDA conditional graph node, the default), replayed
    two-iteration graph with host checks per batch, plain stream launches."""
    rng = np.random.default_rng(H + W)
    cnt, _ = synth.pair(5, H, W)
    lab = color.bgr2lab_u8(cnt)
    a = 1.0 + 0.5 * rng.standard_normal((H, W, 3))
    b = 0.2 * rng.standard_normal((H, W, 3))
    rough = np.where(rng.random((H, W)) < 0.1, 1e-6, 1.0)
    got = {}
    for mode in ("2", "1", "0", "2"):   # mode 2 twice: the second call replays the cached graph
        monkeypatch.setenv("NCT_WLS_LOOP", mode)
        ga, gb = to_dev(a, dev), to_dev(b, dev)
        n0 = ctx.launch_count
        its, res = ctx.solve_wls(ga, gb, to_dev(rough, dev), to_dev(lab, dev), lam, 1.2, rel_tol=1e-8)
        ctx.synchronize()
        r = (its, ga.cpu().numpy().copy(), gb.cpu().numpy().copy(), ctx.launch_count - n0)
        assert res <= 1e-8
        if mode in got:
            assert r[0] == got[mode][0] and np.array_equal(r[1], got[mode][1])
        got[mode] = r
    for mode in ("1", "0"):
        assert got[mode][0] == got["2"][0], f"iteration counts differ: {got[mode][0]} vs {got['2'][0]}"
        assert np.array_equal(got[mode][1].view(np.uint64), got["2"][1].view(np.uint64))
        assert np.array_equal(got[mode][2].view(np.uint64), got["2"][2].view(np.uint64))
    print(f"WLS {H}x{W}: {got['2'][0]} iterations in every loop mode; launches counted: while {got['2'][3]}, graph {got['1'][3]}, stream {got['0'][3]}")


def test_wls_converged_start_returns_immediately(ctx, dev):
    """x0 that already solves the system (constant maps, roughness 1 everywhere): zero iterations, in particular the
    device-side loop must terminate when nothing ever sets a new residual."""
    H, W = 64, 64
    cnt, _ = synth.pair(6, H, W)
    lab = color.bgr2lab_u8(cnt)
    a = np.full((H, W, 3), 1.25)
    b = np.full((H, W, 3), -0.5)
    ga, gb = to_dev(a, dev), to_dev(b, dev)
    its, res = ctx.solve_wls(ga, gb, to_dev(np.ones((H, W)), dev), to_dev(lab, dev), 0.5, 1.2, rel_tol=1e-8)
    assert its == 0 and res <= 1e-8
    assert np.array_equal(ga.cpu().numpy(), a) and np.array_equal(gb.cpu().numpy(), b)


@pytest.mark.parametrize("one_based", [True, False])
def test_solve_direct_with_the_reference_argument_list(ctx, one_based):
    """nct_solve_direct = solve_direct_cpu's own numerical arguments (CT/SparseSolver_CPU.h:35-43): host arrays, UPPER
    triangle of the SPD matrix in CSR (one-based as the reference builds it, CT/ColorTransfer.cpp:951-1099), six
    right-hand sides.  (1) The WLS system the reference assembles, against scipy's direct solve (standing in for PARDISO);
    (2) an unrelated random sparse SPD matrix: the entry point does not depend on the 5-point structure."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    rng = np.random.default_rng(7)
    H, W = 40, 36
    cnt, _ = synth.pair(4, H, W)
    lab = color.bgr2lab_u8(cnt)
    rough = np.where(rng.random((H, W)) < 0.1, 1e-6, 1.0)
    M1 = color.wls_matrix(rough, lab[..., 0] / 255.0, 1.5, 1.2).tocsr()
    n2 = 500
    R = sp.random(n2, n2, density=0.01, random_state=3, format="csr")
    M2 = (R @ R.T + sp.identity(n2) * 0.5).tocsr()
    for M in (M1, M2):
        n = M.shape[0]
        U = sp.triu(M, format="csr")
        U.sort_indices()
        base = 1 if one_based else 0
        B = rng.standard_normal((6, n))
        X, its, res = ctx.solve_direct(U.data, U.indptr + base, U.indices + base, B, one_based=one_based, rel_tol=1e-11)
        ref = spla.splu(M.tocsc()).solve(B.T).T
        err = np.abs(X - ref).max() / np.abs(ref).max()
        print(f"nct_solve_direct n={n} nnz(upper)={U.nnz}: {its} iterations, rel.res {res:.1e}, max rel. error vs direct solve {err:.1e}")
        assert res <= 1e-11 and err < 1e-8


def test_solve_direct_rejects_a_matrix_that_is_not_upper_triangular(pkg, ctx):
    A = np.array([1.0, 0.5, 1.0])
    with pytest.raises(pkg.NctError):
        ctx.solve_direct(A, np.array([0, 1, 3]), np.array([0, 0, 1]), np.zeros((6, 2)), one_based=False)   # entry (1, 0) is below the diagonal


@pytest.mark.parametrize("mode", ["2", "1", "0"])
def test_wls_iteration_budget_is_enforced_on_the_device(pkg, ctx, dev, mode, monkeypatch):
    """max_iters is checked by the kernels themselves (PcgScalars::done): with a budget far below what the system needs the
    solve stops after exactly that many iterations -- in particular the device-side WHILE loop terminates -- and reports
    the failure instead of returning an unconverged result as if it were one."""
    monkeypatch.setenv("NCT_WLS_LOOP", mode)
    rng = np.random.default_rng(1)
    H, W = 96, 112
    cnt, _ = synth.pair(9, H, W)
    lab = color.bgr2lab_u8(cnt)
    a = 1.0 + 0.5 * rng.standard_normal((H, W, 3))
    b = 0.2 * rng.standard_normal((H, W, 3))
    ga, gb = to_dev(a, dev), to_dev(b, dev)
    with pytest.raises(pkg.NctError, match="did not reach"):
        ctx.solve_wls(ga, gb, to_dev(np.ones((H, W)), dev), to_dev(lab, dev), 3.0, 1.2, rel_tol=1e-12, max_iters=5)
    # the context is still usable and the next solve converges
    ga, gb = to_dev(a, dev), to_dev(b, dev)
    its, res = ctx.solve_wls(ga, gb, to_dev(np.ones((H, W)), dev), to_dev(lab, dev), 3.0, 1.2, rel_tol=1e-8)
    assert res <= 1e-8 and its > 5
