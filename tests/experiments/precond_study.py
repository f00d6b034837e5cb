"""Preconditioner study for the reduced pose system (round-1 verdict, item 7): PCG iteration counts on the Schur
complement S = (Hpp + lambda I) - Hpl (Hll + lambda I)^-1 Hlp of a late LM iteration, for candidates that stay inside
or next to the block-Jacobi contract. CPU only (oracle + scipy); not part of the product or of the test suite.

    python tests/experiments/precond_study.py c2 c5s

Convergence criterion = the product's: sqrt(r.z / r0.z0) <= 1e-10 (k_pcg, Chronopoulos-Gear recurrences aside).
"""
import sys, os, time
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu_oracle as co  # noqa: E402
from sparse_gslam_b200 import graphgen as gg  # noqa: E402


def sparse_h(st, lin):
    """scipy CSR of the full symmetric H from the oracle's block list (column-major blocks, upper triangle)."""
    n = st["dim"]
    nr, nc = st["nrows"].astype(np.int64), st["ncols"].astype(np.int64)
    size = nr * nc
    start = np.concatenate([[0], np.cumsum(size)[:-1]])
    rows, cols, vals = [], [], []
    for (a, b) in ((3, 3), (3, 2), (2, 2)):
        m = (nr == a) & (nc == b)
        if not m.any():
            continue
        idx = start[m][:, None] + np.arange(a * b)[None, :]
        v = lin["H"][idx].reshape(-1, b, a).transpose(0, 2, 1)  # [blk, r, c]
        ro = st["offset"][st["row"][m]].astype(np.int64)[:, None, None] + np.arange(a)[None, :, None]
        cof = st["offset"][st["col"][m]].astype(np.int64)[:, None, None] + np.arange(b)[None, None, :]
        ro, cof = np.broadcast_arrays(ro, cof)
        rows.append(ro.ravel()); cols.append(cof.ravel()); vals.append(v.ravel())
    r, c, v = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    U = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
    D = sp.diags(U.diagonal())
    Us = sp.triu(U, 1)
    return (Us + Us.T + D).tocsr()


def late_lm_system(name, lm_iters):
    g = gg.make_c5(rows=200, cols=200) if name == "c5s" else (gg.make_c5(rows=100, cols=100) if name == "c5xs" else gg.make(name))
    o = co.Oracle(g)
    assert o.initialize_optimization()
    lam = None
    if lm_iters > 0:
        n, stats = o.optimize(lm_iters, co.ALGO_LM, co.JAC_ANALYTIC)
        lam = stats[-1]["lambda_"]
    st = o.structure()
    lin = o.linearize(co.JAC_ANALYTIC)
    H = sparse_h(st, lin)
    Pf = int((st["kind"] == 0).sum())
    np3 = 3 * Pf
    if lam is None:
        lam = 1e-5 * H.diagonal().max()
    n = H.shape[0]
    Hpp = H[:np3, :np3] + lam * sp.identity(np3)
    Hpl = H[:np3, np3:]
    Hll = (H[np3:, np3:] + lam * sp.identity(n - np3)).tocsc()
    if n > np3:
        Hll_inv = sp.csr_matrix(spla.inv(Hll)) if n - np3 < 20000 else block2_inv(Hll)
        S = (Hpp - Hpl @ Hll_inv @ Hpl.T).tocsr()
        b = lin["b"][:np3] - Hpl @ (Hll_inv @ lin["b"][np3:])
    else:
        S, b = Hpp.tocsr(), lin["b"][:np3]
    return g, S, b, lam, Pf


def block2_inv(Hll):
    """inverse of a block-diagonal matrix of 2x2 blocks"""
    d = Hll.tocsr()
    n = d.shape[0]
    a = d.diagonal()[0::2]; c = d.diagonal()[1::2]
    off = np.asarray(d[np.arange(0, n, 2), np.arange(1, n, 2)]).ravel()
    det = a * c - off * off
    i = np.arange(0, n, 2)
    r = np.concatenate([i, i + 1, i, i + 1]); cc = np.concatenate([i, i + 1, i + 1, i])
    v = np.concatenate([c / det, a / det, -off / det, -off / det])
    return sp.csr_matrix((v, (r, cc)), shape=(n, n))


def pcg(S, b, M, tol=1e-10, maxit=20000):
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy()
    g0 = g = r @ z
    for it in range(1, maxit + 1):
        q = S @ p
        a = g / (p @ q)
        x += a * p; r -= a * q
        z = M(r); gn = r @ z
        if np.sqrt(abs(gn) / g0) <= tol:
            return it, x
        p = z + (gn / g) * p; g = gn
    return maxit, x


def block_mask(n, starts):
    """sparse 0/1 pattern of the diagonal blocks [starts[i], starts[i+1])"""
    lab = np.zeros(n, np.int64)
    lab[starts[1:-1]] = 1
    return np.cumsum(lab)


def restrict(S, keep_fn):
    C = S.tocoo()
    m = keep_fn(C.row, C.col)
    return sp.csc_matrix((C.data[m], (C.row[m], C.col[m])), shape=S.shape)


def factor(A):
    lu = spla.splu(A.tocsc(), permc_spec="NATURAL", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    return lu.solve


def bj(S, poses):
    n = S.shape[0]
    lab = (np.arange(n) // (3 * poses))
    return factor(restrict(S, lambda r, c: lab[r] == lab[c]))


def band(S, w, seg=None):
    """exact solve with the part of S within w poses of the diagonal (optionally cut at segment borders)"""
    pr = np.arange(S.shape[0]) // 3
    if seg is None:
        A = restrict(S, lambda r, c: np.abs(pr[r] - pr[c]) <= w)
    else:
        sg = pr // seg
        A = restrict(S, lambda r, c: (np.abs(pr[r] - pr[c]) <= w) & (sg[r] == sg[c]))
    return factor(A)


def two_level(S, fine, agg, mode="const"):
    """fine block-Jacobi + additive coarse correction on aggregates of `agg` consecutive poses (3 dof each: the rigid
    x / y / theta shifts of the aggregate, piecewise constant)"""
    n = S.shape[0]
    P = n // 3
    a = np.arange(P) // agg
    na = a.max() + 1
    rows = np.arange(n)
    cols = 3 * a[rows // 3] + rows % 3
    R = sp.csr_matrix((np.ones(n), (cols, rows)), shape=(3 * na, n))
    Ac = (R @ S @ R.T).tocsc()
    cs = spla.splu(Ac).solve
    return (lambda r: fine(r) + R.T @ cs(R @ r)), Ac


def two_level_lin(S, fine, agg):
    """coarse space = piecewise LINEAR hat functions over the chain (nodes every `agg` poses)"""
    n = S.shape[0]
    P = n // 3
    t = np.arange(P) / agg
    i0 = np.floor(t).astype(np.int64); w1 = t - i0; w0 = 1 - w1
    na = i0.max() + 2
    rows = np.arange(n); pr = rows // 3; d = rows % 3
    R = sp.csr_matrix((np.concatenate([w0[pr], w1[pr]]), (np.concatenate([3 * i0[pr] + d, 3 * (i0[pr] + 1) + d]),
                                                          np.concatenate([rows, rows]))), shape=(3 * na, n))
    Ac = (R @ S @ R.T).tocsc() + 1e-12 * sp.identity(3 * na)
    cs = spla.splu(Ac).solve
    return (lambda r: fine(r) + R.T @ cs(R @ r)), Ac


def level_r(n, agg):
    a = np.arange(n // 3) // agg
    rows = np.arange(n)
    return sp.csr_matrix((np.ones(n), (3 * a[rows // 3] + rows % 3, rows)), shape=(3 * (a.max() + 1), n))


def multilevel_diagonal(S, fine, aggs):
    """BPX-style additive multilevel: fine block-Jacobi + sum over levels of R_l^T blockdiag3(R_l S R_l^T)^-1 R_l (no
    coarse solve anywhere: what a kernel could apply with segmented sums only)"""
    ops = []
    for agg in aggs:
        R = level_r(S.shape[0], agg)
        Ac = (R @ S @ R.T).tocsr()
        lab = np.arange(Ac.shape[0]) // 3
        ops.append((R, factor(restrict(Ac, lambda r, c: lab[r] == lab[c]))))

    def M(r):
        z = fine(r)
        for R, D in ops:
            z = z + R.T @ D(R @ r)
        return z
    return M


def capped(S, fine, nagg_max, align=8):
    """two-level, at most nagg_max aggregates (a coarse matrix small enough to invert densely on chip)"""
    P = S.shape[0] // 3
    agg = -(-P // nagg_max)
    agg = -(-agg // align) * align
    return two_level(S, fine, agg)[0]


def main():
    names = sys.argv[1:] or ["c2"]
    for name in names:
        t0 = time.time()
        g, S, b, lam, Pf = late_lm_system(name, int(os.environ.get("LM_ITERS", "8")))
        print(f"== {name}: {Pf} free poses, nnz(S) {S.nnz}, lambda {lam:.3e}  ({time.time() - t0:.1f} s)")
        base = None
        cands = [("block-Jacobi 4 poses (product)", lambda: bj(S, 4)), ("block-Jacobi 8", lambda: bj(S, 8)),
                 ("block-Jacobi 16", lambda: bj(S, 16)), ("block-Jacobi 32", lambda: bj(S, 32)),
                 ("chain tridiagonal, 32-pose segments", lambda: band(S, 1, 32)),
                 ("chain tridiagonal, whole chain", lambda: band(S, 1)),
                 ("band +-4 poses, whole chain", lambda: band(S, 4)),
                 ("band +-16 poses, whole chain", lambda: band(S, 16)),
                 ("BJ4 + coarse const agg 16", lambda: two_level(S, bj(S, 4), 16)[0]),
                 ("BJ4 + coarse const agg 64", lambda: two_level(S, bj(S, 4), 64)[0]),
                 ("BJ4 + coarse linear agg 16", lambda: two_level_lin(S, bj(S, 4), 16)[0]),
                 ("BJ4 + coarse linear agg 64", lambda: two_level_lin(S, bj(S, 4), 64)[0]),
                 ("BJ4 + coarse const, <= 32 aggregates", lambda: capped(S, bj(S, 4), 32)),
                 ("BJ4 + coarse const, <= 128 aggregates", lambda: capped(S, bj(S, 4), 128)),
                 ("BJ4 + multilevel diagonal 16/64/256", lambda: multilevel_diagonal(S, bj(S, 4), [16, 64, 256]))]
        for label, mk in cands:
            t0 = time.time()
            try:
                M = mk()
                it, x = pcg(S, b, M)
                res = np.linalg.norm(S @ x - b) / np.linalg.norm(b)
            except Exception as e:  # a band of an SPD matrix need not be SPD
                print(f"   {label:40s} failed: {type(e).__name__} {e}")
                continue
            base = base or it
            print(f"   {label:40s} {it:6d} iterations  ({it / base:5.2f} x)  true residual {res:.1e}  {time.time() - t0:.1f} s", flush=True)


if __name__ == "__main__":
    main()
