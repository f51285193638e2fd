"""Dev tool (never imported by the product or the tests): coarse-space variants for the WLS multigrid-PCG on the model of
tools/wls_mg_model.py - harmonic coarse edge weights, operator-dependent (resistance-weighted) interpolation, Galerkin coarse
operators.  python tools/wls_coarse_space_model.py systems.npz [levels...]; results in profiles/r2_wls_tuning.md."""
import sys, numpy as np, scipy.sparse as sp
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
from wls_mg_model import Hierarchy, gradient_weights, pcg, laplacian, coarsen

def half_res(wa):  # 1/(2w), 0 where edge missing (w==0 from padding)
    out = np.zeros_like(wa); m = wa > 0; out[m] = 0.5 / wa[m]; return out

def axis_weights(wx):
    """per fine node: a = interpolation weight towards the neighbouring aggregate along x; plus coarse conductances"""
    h, w = wx.shape
    wc = (w + 1) // 2
    # edge arrays: e[x] = wx[:,x] is edge (x,x+1), valid for x < w-1
    def E(x):  # returns column or zeros
        if np.isscalar(x):
            return wx[:, x] if 0 <= x < w - 1 else np.zeros(h)
    a = np.zeros((h, w))
    for x in range(w):
        if x & 1:  # right node of aggregate J -> neighbour J+1 if exists (x+1 < w)
            if x + 1 >= w: continue
            near = half_res(E(x - 1)); cross = E(x); far = half_res(E(x + 1))
        else:
            if x - 1 < 0: continue
            near = half_res(E(x)); cross = E(x - 1); far = half_res(E(x - 2))
        tot = near + 1.0 / cross + far
        a[:, x] = near / tot
    # coarse conductance per fine row between J and J+1: cross edge x=2J+1
    cfine = np.zeros((h, wc))
    for J in range(wc - 1):
        x = 2 * J + 1
        tot = half_res(E(x - 1)) + 1.0 / E(x) + half_res(E(x + 1))
        cfine[:, J] = 1.0 / tot
    return a, cfine

class HH(Hierarchy):
    def __init__(self, rough, wx, wy, harmonic_edges=True, opdep_P=True, galerkin=False, min_n=1):
        self.levels = []
        rs = rough
        M, d = laplacian(rs, wx, wy)
        while True:
            h, w = rs.shape
            lev = dict(M=M, invd=1.0 / M.diagonal(), shape=rs.shape)
            self.levels.append(lev)
            if rs.size <= min_n: break
            hc, wc = (h + 1) // 2, (w + 1) // 2
            ax, cxf = axis_weights(wx)
            ayT, cyfT = axis_weights(wy.T.copy())
            ay, cyf = ayT.T, cyfT.T   # cyf shape (hc, w)
            if not opdep_P:
                ax = np.where(ax > 0, 0.25, 0.0); ay = np.where(ay > 0, 0.25, 0.0)
            # build P
            yy, xx = np.mgrid[0:h, 0:w]
            Jx, Jy = xx >> 1, yy >> 1
            Nx = np.clip(Jx + np.where(xx & 1, 1, -1), 0, wc - 1); Ny = np.clip(Jy + np.where(yy & 1, 1, -1), 0, hc - 1)
            fi = (yy * w + xx).ravel()
            rows = np.concatenate([fi] * 4)
            cols = np.concatenate([(Jy * wc + Jx).ravel(), (Jy * wc + Nx).ravel(), (Ny * wc + Jx).ravel(), (Ny * wc + Nx).ravel()])
            vals = np.concatenate([((1 - ax) * (1 - ay)).ravel(), (ax * (1 - ay)).ravel(), ((1 - ax) * ay).ravel(), (ax * ay).ravel()])
            P = sp.csr_matrix((vals, (rows, cols)), shape=(h * w, hc * wc)); P.sum_duplicates()
            lev["P"] = P
            # coarse
            rs_c, wx_c, wy_c = coarsen(rs, wx, wy, 0.5)
            if harmonic_edges:
                pad_r = np.pad(cxf, ((0, 2 * hc - h), (0, 0)))
                wx_c = pad_r.reshape(hc, 2, wc).sum(1)
                pad_c = np.pad(cyf, ((0, 0), (0, 2 * wc - w)))
                wy_c = pad_c.reshape(hc, wc, 2).sum(2)
            rs, wx, wy = rs_c, wx_c, wy_c
            if galerkin:
                M = (P.T @ M @ P).tocsr()
            else:
                M, d = laplacian(rs, wx, wy)

g = np.load(sys.argv[1]); Lc = g["L"]
levels = [int(a) for a in sys.argv[2:]] or [0, 2, 4]
for l in levels:
    lam = float(g[f"lam_{l}"]) * (4 if l == 4 else 1)
    rough = g[f"rough_{l}"]
    wx, wy = gradient_weights(Lc, lam)
    x0 = np.concatenate([g[f"a2_{l}"].reshape(-1, 3), g[f"b2_{l}"].reshape(-1, 3)], 1)
    rhs = rough.reshape(-1, 1) * x0
    row = [f"level {l} lam {lam:.3f}:"]
    H = Hierarchy(rough, wx, wy, edge_scale=0.5); M = H.levels[0]["M"]
    _, its, _ = pcg(M, rhs, x0, lambda r: H.vcycle(r, (0.55, 1.7), 0, 1)); row.append(f"product: {its}")
    for name, kw in [("harm edges only", dict(opdep_P=False)), ("opdep P only", dict(harmonic_edges=False)), ("harm+opdep", dict()), ("opdep+galerkin", dict(galerkin=True)), ("linear P + galerkin", dict(opdep_P=False, galerkin=True))]:
        H2 = HH(rough, wx, wy, **kw)
        _, its, rel = pcg(M, rhs, x0, lambda r: H2.vcycle(r, (0.55, 1.7), 0, 1), maxit=150); row.append(f"{name}: {its} ({rel:.0e})")
    print("  ".join(row), flush=True)
