"""CPU model (numpy / scipy) of the WLS multigrid-PCG of csrc/wls_mg.cu, for exploring preconditioner variants without a
GPU: the same aggregation hierarchy (2x2 cells, coarse 5-point operators with edge scale 0.5), cell-centred linear
interpolation P (0.75 / 0.25 per axis, clamped), R = P^T, damped-Jacobi sweeps with per-sweep damping factors, one
symmetric V-cycle as the preconditioner of CG on six right-hand sides, stop at max_k ||r_k|| / ||b_k|| <= tol.

  python tools/wls_mg_model.py systems.npz            (systems dumped from the oracle pipeline: a2_l, b2_l, rough_l, lam_l, L)

Dev tool only (never imported by the product or the tests)."""
import sys

import numpy as np
import scipy.sparse as sp


def gradient_weights(L, lam, alpha=1.2):
    h, w = L.shape
    wx = np.zeros((h, w))
    wy = np.zeros((h, w))
    wx[:, :-1] = lam / (np.abs(L[:, 1:] - L[:, :-1]) ** alpha + 1e-4)
    wy[:-1, :] = lam / (np.abs(L[1:, :] - L[:-1, :]) ** alpha + 1e-4)
    return wx, wy  # squared g: the operator's edge weights


def laplacian(rsum, wx, wy):
    h, w = rsum.shape
    n = h * w
    idx = np.arange(n).reshape(h, w)
    d = rsum.copy()
    d[:, :-1] += wx[:, :-1]
    d[:, 1:] += wx[:, :-1]
    d[:-1, :] += wy[:-1, :]
    d[1:, :] += wy[:-1, :]
    r = [idx.ravel(), idx[:, :-1].ravel(), idx[:, 1:].ravel(), idx[:-1, :].ravel(), idx[1:, :].ravel()]
    c = [idx.ravel(), idx[:, 1:].ravel(), idx[:, :-1].ravel(), idx[1:, :].ravel(), idx[:-1, :].ravel()]
    v = [d.ravel(), -wx[:, :-1].ravel(), -wx[:, :-1].ravel(), -wy[:-1, :].ravel(), -wy[:-1, :].ravel()]
    return sp.csr_matrix((np.concatenate(v), (np.concatenate(r), np.concatenate(c))), shape=(n, n)), d


def coarsen(rsum, wx, wy, edge_scale):
    h, w = rsum.shape
    hc, wc = (h + 1) // 2, (w + 1) // 2
    pad = lambda a: np.pad(a, ((0, 2 * hc - h), (0, 2 * wc - w)))  # noqa: E731
    rs = pad(rsum).reshape(hc, 2, wc, 2).sum((1, 3))
    wxp, wyp = pad(wx), pad(wy)
    cx = np.zeros((hc, wc))
    cy = np.zeros((hc, wc))
    # edges crossing from aggregate J to J+1: fine edges (x = 2J+1 -> 2J+2) of both rows of the aggregate
    cx[:, :-1] = (wxp[:, 1::2].reshape(hc, 2, wc).sum(1))[:, :-1]
    cy[:-1, :] = (wyp[1::2, :].reshape(hc, wc, 2).sum(2))[:-1, :]
    return rs, edge_scale * cx, edge_scale * cy


def interp_1d(nf, nc):
    rows, cols, vals = [], [], []
    for x in range(nf):
        jp = x >> 1
        jn = min(max(jp + (1 if x & 1 else -1), 0), nc - 1)
        rows += [x, x]
        cols += [jp, jn]
        vals += [0.75, 0.25]
    return sp.csr_matrix((vals, (rows, cols)), shape=(nf, nc))


class Hierarchy:
    def __init__(self, rough, wx, wy, edge_scale=0.5, min_n=1):
        self.levels = []
        rs = rough
        while True:
            M, d = laplacian(rs, wx, wy)
            self.levels.append(dict(M=M, invd=1.0 / d.ravel(), shape=rs.shape))
            if rs.size <= min_n:
                break
            rs, wx, wy = coarsen(rs, wx, wy, edge_scale)
        for k in range(len(self.levels) - 1):
            (hf, wf), (hc, wc) = self.levels[k]["shape"], self.levels[k + 1]["shape"]
            self.levels[k]["P"] = sp.kron(interp_1d(hf, hc), interp_1d(wf, wc)).tocsr()

    def vcycle(self, b, omegas, k=0, gamma=1):
        L = self.levels[k]
        M, invd = L["M"], L["invd"][:, None]
        if k == len(self.levels) - 1:
            if M.shape[0] == 1:
                return b * invd
        x = np.zeros_like(b)
        for om in omegas:  # pre-smoothing from a zero guess
            x = x + om * invd * (b - M @ x)
        if k < len(self.levels) - 1:
            P = L["P"]
            for _ in range(gamma):
                x = x + P @ self.vcycle(P.T @ (b - M @ x), omegas, k + 1, gamma)
        for om in omegas:  # post-smoothing (the factors commute, so the order does not matter for symmetry)
            x = x + om * invd * (b - M @ x)
        return x


def pcg(M, rhs, x0, precond, tol=1e-8, maxit=400):
    x = x0.copy()
    r = rhs - M @ x
    bb = (rhs * rhs).sum(0)
    z = precond(r)
    p = z.copy()
    rz = (r * z).sum(0)
    for it in range(1, maxit + 1):
        Ap = M @ p
        alpha = rz / (p * Ap).sum(0)
        x += alpha * p
        r -= alpha * Ap
        rel = np.sqrt((r * r).sum(0) / bb).max()
        if rel <= tol:
            return x, it, rel
        z = precond(r)
        rz_new = (r * z).sum(0)
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, maxit, rel


def main():
    g = np.load(sys.argv[1])
    Lc = g["L"]
    variants = {
        "jacobi 0.8/0.8 (round-1 start)": dict(omegas=(0.8, 0.8)),
        "pair 0.55/1.7 (round-1 final)": dict(omegas=(0.55, 1.7)),
        "3 sweeps 0.52/0.8/1.9": dict(omegas=(0.52, 0.8, 1.9)),
        "3 sweeps 0.55/1.0/2.2": dict(omegas=(0.55, 1.0, 2.2)),
        "pair, W-cycle": dict(omegas=(0.55, 1.7), gamma=2),
        "pair, edge scale 0.5 -> Galerkin-like 0.25": dict(omegas=(0.55, 1.7), edge_scale=0.25),
    }
    for l in range(5):
        lam = float(g[f"lam_{l}"]) * (4 if l == 4 else 1)
        rough = g[f"rough_{l}"]
        wx, wy = gradient_weights(Lc, lam)
        x0 = np.concatenate([g[f"a2_{l}"].reshape(-1, 3), g[f"b2_{l}"].reshape(-1, 3)], 1)
        rhs = rough.reshape(-1, 1) * x0
        row = [f"level {l} lam {lam:.3f}:"]
        for name, v in variants.items():
            H = Hierarchy(rough, wx, wy, edge_scale=v.get("edge_scale", 0.5))
            M = H.levels[0]["M"]
            _, its, rel = pcg(M, rhs, x0, lambda r: H.vcycle(r, v["omegas"], 0, v.get("gamma", 1)))
            row.append(f"{name}: {its}")
        print("  ".join(row), flush=True)


if __name__ == "__main__":
    main()
