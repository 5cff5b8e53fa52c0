"""Scratch study (CPU, numpy/scipy): PCG step counts of the data-term solve under different preconditioners on a real frame of the
bench sequence (oracle loop).  Not part of the product or the tests."""
import sys, time
from pathlib import Path
import numpy as np
import scipy.sparse as sp
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench
from oracle import orc, orc_pipe

T = int(sys.argv[1]) if len(sys.argv) > 1 else 30
orc.build()
frames = bench.make_frames(T + 1, 0)
k = orc_pipe.KinFu(bench.cpu_params())
for t in range(T):
    k(frames[t])
nodes_before = k.buffer("nodes").copy()
k(frames[T])
vis = k.buffer("canonical_visible").reshape(-1, 4)
live = k.buffer("curr_points").reshape(-1, 4)
R, tt = k.getCameraPose(-1)
Ri = R.T; ti = -Ri @ tt
canon = vis.copy(); canon[:, :3] = vis[:, :3] @ Ri.T + ti
M = len(nodes_before)
print("frame", T, "nodes", M, "oracle stats", k.buffer("solve_stats"))
idx, d2 = orc_pipe.knn8_fast(nodes_before, canon)
valid = ~(np.isnan(canon[:, :3]).any(1) | np.isnan(live[:, :3]).any(1))
nw = nodes_before[:, 11]
w = np.exp(-(d2.astype(np.float64)) / (2 * nw[np.maximum(idx, 0)].astype(np.float64) ** 2)).astype(np.float32).astype(np.float64)
w[(idx < 0) | ~valid[:, None]] = 0
rows = np.repeat(np.arange(len(canon)), 8)
W = sp.csr_matrix((w.ravel(), (rows, np.maximum(idx, 0).ravel())), shape=(len(canon), M))
A = (W.T @ W).tocsr()
b = np.where(valid[:, None], (live[:, :3] - canon[:, :3]).astype(np.float64), 0.0)
gb = W.T @ b
print("nnz", A.nnz, "nnz/row", A.nnz / M, "valid", valid.sum())
# current translations
x0 = np.zeros((M, 3))
for m in range(M):
    pass
tr = orc.node_translations(nodes_before)[:, 1:].astype(np.float64)
x0 = tr
diag = A.diagonal()
c0 = 0.5 * (b * b).sum()

def cost_of(x): return c0 + (x * (0.5 * (A @ x) - gb)).sum()

def morton_order(pts):
    p = pts - pts.min(0); p = (p / (p.max() + 1e-9) * 1023).astype(np.int64)
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF; v = (v | (v << 8)) & 0x0300F00F; v = (v | (v << 4)) & 0x030C30C3; v = (v | (v << 2)) & 0x09249249; return v
    return np.argsort(spread(p[:, 0]) | (spread(p[:, 1]) << 1) | (spread(p[:, 2]) << 2), kind="stable")

def make_prec(kind, radius, bs=8):
    cd = np.clip(diag, 1e-6, 1e32) / radius
    if kind == "jacobi":
        mi = 1.0 / (diag + cd)
        return lambda r: r * mi[:, None]
    if kind == "block":
        order = morton_order(nodes_before[:, 0:3] if False else node_pos)
        Ad = (A + sp.diags(cd)).tocsr()
        blocks = [order[i:i + bs] for i in range(0, M, bs)]
        invs = [np.linalg.inv(Ad[bk][:, bk].toarray()) for bk in blocks]
        def app(r):
            z = np.empty_like(r)
            for bk, iv in zip(blocks, invs): z[bk] = iv @ r[bk]
            return z
        return app
    if kind == "ssor":
        Ad = (A + sp.diags(cd)).tocsr()
        from scipy.sparse.linalg import spsolve_triangular
        L = sp.tril(Ad).tocsr(); U = sp.triu(Ad).tocsr(); D = Ad.diagonal()
        return lambda r: spsolve_triangular(U, D[:, None] * spsolve_triangular(L, r, lower=True), lower=False)
    raise ValueError

def node_positions():
    # node position = translation part of dq? nodes row: [pos? ...] -- take from the oracle helper if present
    return nodes_before[:, 0:3].astype(np.float64)
node_pos = node_positions()

def lm(kind, bs=8, lm_iters=5, lin_iters=100, qtol=1e-4):
    x = x0.copy(); cost = cost_of(x); radius, decrease = 1e4, 2.0; total = 0; per = []
    for it in range(lm_iters):
        cd = np.clip(diag, 1e-6, 1e32) / radius
        g = gb - A @ x
        prec = make_prec(kind, radius, bs)
        dl = np.zeros_like(x); r = g.copy(); z = prec(r); p = z.copy(); rz = (r * z).sum(); Q0 = 0.0; n = 0
        for l in range(lin_iters):
            if not rz > 0: break
            Ap = A @ p + cd[:, None] * p
            pAp = (p * Ap).sum()
            if not pAp > 0: break
            alpha = rz / pAp
            dl += alpha * p; r -= alpha * Ap; z = prec(r); rz_new = (r * z).sum()
            Q1 = -0.5 * (dl * (r + g)).sum(); beta = rz_new / rz; p = z + beta * p; rz = rz_new; n += 1
            zeta = (l + 1) * (Q1 - Q0) / Q1; Q0 = Q1
            if zeta < qtol: break
        total += n; per.append(n)
        model = 0.5 * (dl * (g + r + cd[:, None] * dl)).sum()
        new_cost = cost - (dl * g).sum() + 0.5 * (dl * (g - r - cd[:, None] * dl)).sum()
        change = cost - new_cost; rho = change / model if model > 0 else 0.0
        stop = False
        if change >= 0 and rho > 1e-3:
            x = x + dl; stop = change <= cost * 1e-6; cost = new_cost
            f = 1 - (2 * rho - 1) ** 3; radius = min(radius / max(f, 1 / 3), 1e16); decrease = 2.0
        else:
            radius /= decrease; decrease *= 2
        if stop: break
    return total, per, cost, x

from scipy.sparse.linalg import spsolve
xs = np.stack([spsolve((A + sp.diags(np.full(M, 1e-12))).tocsc(), gb[:, d]) for d in range(3)], 1)
print("cost0", cost_of(x0), "optimum", cost_of(xs))
for kind, bs in (("jacobi", 0), ("block", 4), ("block", 8), ("block", 16), ("block", 32), ("ssor", 0)):
    t0 = time.time()
    total, per, cost, x = lm(kind, bs)
    print(f"{kind:7s} bs={bs:2d}: pcg steps {total:4d} {per}  final cost {cost:.9g}  rel to optimum {(cost - cost_of(xs)) / cost_of(xs):.3e}  max|dx| vs jacobi n/a  ({time.time() - t0:.1f}s)")
