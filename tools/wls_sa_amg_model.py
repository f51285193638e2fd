"""Dev tool (never imported by the product or the tests): smoothed-aggregation AMG (strength threshold theta, greedy
aggregation, Jacobi-smoothed prolongator on the filtered matrix, Galerkin coarse operators) as the PCG preconditioner of the
WLS systems, against the product's geometric hierarchy.  python tools/wls_sa_amg_model.py systems.npz [solves...];
results in profiles/r2_wls_tuning.md."""
import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
from wls_mg_model import Hierarchy, gradient_weights, pcg

def strength(A, theta):
    A = A.tocsr(); n = A.shape[0]
    C = A.copy().tocoo()
    off = C.row != C.col
    absoff = np.abs(C.data) * off
    rowmax = np.zeros(n); np.maximum.at(rowmax, C.row, absoff)
    keep = off & (np.abs(C.data) >= theta * rowmax[C.row]) & (rowmax[C.row] > 0)
    S = sp.csr_matrix((np.ones(keep.sum()), (C.row[keep], C.col[keep])), shape=(n, n))
    S = ((S + S.T) > 0).astype(np.int8).tocsr()   # symmetrise
    return S

def aggregate(S):
    n = S.shape[0]; indptr, idx = S.indptr, S.indices
    agg = -np.ones(n, np.int64); na = 0
    for i in range(n):   # pass 1
        if agg[i] >= 0: continue
        nb = idx[indptr[i]:indptr[i + 1]]
        if nb.size and (agg[nb] < 0).all():
            agg[i] = na; agg[nb] = na; na += 1
    for i in range(n):   # pass 2: attach to a neighbouring aggregate
        if agg[i] >= 0: continue
        nb = idx[indptr[i]:indptr[i + 1]]
        a = agg[nb]; a = a[a >= 0]
        if a.size: agg[i] = -2 - a[0]
    m = agg <= -2; agg[m] = -2 - agg[m]
    for i in range(n):   # pass 3
        if agg[i] >= 0: continue
        agg[i] = na
        nb = idx[indptr[i]:indptr[i + 1]]
        for j in nb:
            if agg[j] < 0: agg[j] = na
        na += 1
    return agg, na

def sa_level(A, theta, smooth=True):
    S = strength(A, theta)
    agg, na = aggregate(S)
    n = A.shape[0]
    Pt = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, na))
    if not smooth: return Pt
    # filtered matrix: weak off-diagonals lumped into the diagonal
    Ac = A.tocoo(); strong = S.tocsr()
    isstrong = np.asarray(strong[Ac.row, Ac.col]).ravel() > 0
    offd = Ac.row != Ac.col
    keep = isstrong | ~offd
    AF = sp.csr_matrix((Ac.data[keep], (Ac.row[keep], Ac.col[keep])), shape=A.shape)
    lump = np.zeros(n); np.add.at(lump, Ac.row[~keep], Ac.data[~keep])
    AF = AF + sp.diags(lump)
    Dinv = 1.0 / AF.diagonal()
    DA = sp.diags(Dinv) @ AF
    rho = abs(spla.eigs(DA, k=1, which='LM', return_eigenvectors=False, maxiter=50, tol=1e-2)[0]) if n > 10 else 2.0
    return (Pt - (4.0 / 3.0 / rho) * (DA @ Pt)).tocsr()

class SA:
    def __init__(self, A, theta=0.25, smooth=True, min_n=200):
        self.lv = []
        while True:
            L = dict(A=A.tocsr(), invd=1.0 / A.diagonal())
            self.lv.append(L)
            if A.shape[0] <= min_n: L["lu"] = spla.splu(A.tocsc()); break
            P = sa_level(A, theta, smooth)
            if P.shape[1] >= 0.9 * A.shape[0]: L["lu"] = spla.splu(A.tocsc()); break
            L["P"] = P
            A = (P.T @ A @ P).tocsr()
    def cyc(self, b, om, k=0):
        L = self.lv[k]; A = L["A"]; invd = L["invd"][:, None]
        if "lu" in L: return L["lu"].solve(b)
        x = np.zeros_like(b)
        for o in om: x = x + o * invd * (b - A @ x)
        P = L["P"]
        x = x + P @ self.cyc(P.T @ (b - A @ x), om, k + 1)
        for o in om: x = x + o * invd * (b - A @ x)
        return x
    def info(self):
        return " -> ".join(f"{L['A'].shape[0]}({L['A'].nnz / L['A'].shape[0]:.1f})" for L in self.lv)

g = np.load(sys.argv[1]); Lc = g["L"]
for l in [int(a) for a in sys.argv[2:]] or [0, 2, 4]:
    lam = float(g[f"lam_{l}"]) * (4 if l == 4 else 1)
    rough = g[f"rough_{l}"]
    wx, wy = gradient_weights(Lc, lam)
    x0 = np.concatenate([g[f"a2_{l}"].reshape(-1, 3), g[f"b2_{l}"].reshape(-1, 3)], 1)
    rhs = rough.reshape(-1, 1) * x0
    H = Hierarchy(rough, wx, wy, edge_scale=0.5); M = H.levels[0]["M"]
    _, its, _ = pcg(M, rhs, x0, lambda r: H.vcycle(r, (0.55, 1.7), 0, 1))
    print(f"solve {l}: product {its}", flush=True)
    for theta in (0.5, 0.7):
        for smooth in (True,):
            t = time.time(); h = SA(M, theta, smooth)
            for om, name in (((0.55, 1.7), "V(2,2) pair"), ((0.8,), "V(1,1) 0.8")):
                _, its, rel = pcg(M, rhs, x0, lambda r: h.cyc(r, om), maxit=200)
                print(f"   theta {theta} {'SA' if smooth else 'UA'} {name}: {its} ({rel:.0e})  levels {h.info()}  setup {time.time()-t:.0f}s", flush=True)
