"""Independent numpy restatement of the reference's optimiser path. TEST INFRASTRUCTURE, not product code.

PARITY UNPINNED (see oracle/sgo_oracle.h): g2o/Eigen are not under /root/reference. This second restatement is
written in a different style from oracle/sgo_oracle.cpp (dense Hessian, SE2 as 3x3 homogeneous matrices where
possible, `numpy.linalg.solve`) so that the two oracles cannot share an indexing or sign bug. Small graphs only.

What it follows:
  * reference src/sparse_gslam/src/g2o_bindings/edge_se2_rhotheta.cpp:9-16 (pose-line error)
  * reference src/ls_extractor/include/ls_extractor/utils.h:23-45 (transform_line, checkRhoTheta)
  * reference src/sparse_gslam/src/g2o_bindings/vertex_rhotheta.cpp:30-34 (additive update, no wrap)
  * g2o 2020.5.29 (un-vendored): EdgeSE2, BaseBinaryEdge numeric Jacobian and quadratic form,
    RobustKernelDCS, buildIndexMapping, OptimizationAlgorithmLevenberg / GaussNewton.
"""
from __future__ import annotations

import math

import numpy as np

PI = math.pi


def normalize_theta(t: float) -> float:
    if -PI <= t < PI:
        return t
    m = math.floor(t / (2 * PI))
    t = t - m * 2 * PI
    if t >= PI:
        t -= 2 * PI
    if t < -PI:
        t += 2 * PI
    return t


def se2_mat(x, y, th):
    c, s = math.cos(th), math.sin(th)
    return np.array([[c, -s, x], [s, c, y], [0, 0, 1.0]])


def se2_inverse(p):
    """g2o SE2::inverse(): angle normalised first, translation rotated by the inverse rotation."""
    th = normalize_theta(-p[2])
    c, s = math.cos(th), math.sin(th)
    return np.array([c * (-p[0]) - s * (-p[1]), s * (-p[0]) + c * (-p[1]), th])


def se2_compose(a, b):
    c, s = math.cos(a[2]), math.sin(a[2])
    return np.array([a[0] + c * b[0] - s * b[1], a[1] + s * b[0] + c * b[1], normalize_theta(a[2] + b[2])])


def pp_error(xi, xj, z):
    zinv = se2_inverse(z)
    return se2_compose(zinv, se2_compose(se2_inverse(xi), xj))


def pp_jacobians(xi, xj, z):
    zinv = se2_inverse(z)
    c, s = math.cos(xi[2]), math.sin(xi[2])
    dx, dy = xj[0] - xi[0], xj[1] - xi[1]
    Ai = np.array([[-c, -s, -s * dx + c * dy], [s, -c, -c * dx - s * dy], [0, 0, -1.0]])
    Bj = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1.0]])
    Z = np.eye(3)
    Z[:2, :2] = se2_mat(0, 0, zinv[2])[:2, :2]
    return Z @ Ai, Z @ Bj


def transform_line(line, t, ang):
    rho, al = float(line[0]), float(line[1])
    al += ang
    if al > PI:
        al -= 2 * PI
    if al < -PI:
        al += 2 * PI
    rho += t[0] * math.cos(al) + t[1] * math.sin(al)
    if rho < 0.0:
        rho = -rho
        al += PI
        if al > PI:
            al -= 2 * PI
    return np.array([rho, al])


def pl_error(pose, line, z):
    pinv = se2_inverse(pose)
    r = transform_line(line, pinv[:2], pinv[2])
    e = np.array([z[0] - r[0], z[1] - r[1]])
    e[1] = normalize_theta(e[1])
    return e


def pose_oplus(p, u):
    return np.array([p[0] + u[0], p[1] + u[1], normalize_theta(p[2] + u[2])])


def lm_oplus(l, u):
    return np.array([l[0] + u[0], l[1] + u[1]])  # vertex_rhotheta.cpp:33 discards the wrap


def pl_jac_numeric(pose, line, z, pose_free=True, lm_free=True):
    """BaseBinaryEdge::linearizeOplus default: central differences, delta = 1e-9."""
    delta = 1e-9
    scalar = 1 / (2 * delta)
    A = np.zeros((2, 3))
    B = np.zeros((2, 2))
    if pose_free:
        for d in range(3):
            add = np.zeros(3)
            add[d] = delta
            e1 = pl_error(pose_oplus(pose, add), line, z)
            add[d] = -delta
            e2 = pl_error(pose_oplus(pose, add), line, z)
            A[:, d] = scalar * (e1 - e2)
    if lm_free:
        for d in range(2):
            add = np.zeros(2)
            add[d] = delta
            e1 = pl_error(pose, lm_oplus(line, add), z)
            add[d] = -delta
            e2 = pl_error(pose, lm_oplus(line, add), z)
            B[:, d] = scalar * (e1 - e2)
    return A, B


def pl_jac_analytic(pose, line):
    ca, sa = math.cos(line[1]), math.sin(line[1])
    q = line[0] - pose[0] * ca - pose[1] * sa
    s = 1.0 if q >= 0 else -1.0
    A = np.array([[s * ca, s * sa, 0.0], [0.0, 0.0, 1.0]])
    B = np.array([[-s, -s * (pose[0] * sa - pose[1] * ca)], [0.0, -1.0]])
    return A, B


def dcs(phi, e2):
    scale = (2.0 * phi) / (phi + e2)
    if scale >= 1.0:
        return e2, 1.0
    return scale * e2 * scale, scale * scale


def _full3(u):
    return np.array([[u[0], u[1], u[2]], [u[1], u[3], u[4]], [u[2], u[4], u[5]]])


def _full2(u):
    return np.array([[u[0], u[1]], [u[1], u[2]]])


class PyOracle:
    def __init__(self, g):
        self.g = g
        self.pose = np.array(g.pose_est, dtype=np.float64, copy=True)
        self.lm = np.array(g.lm_est, dtype=np.float64, copy=True)
        self.lam = 0.0
        self.ni = 2.0

    # ---- SparseOptimizer::initializeOptimization / buildIndexMapping
    def initialize_optimization(self):
        g = self.g
        edges = []
        for k in range(g.n_pp):
            if g.pose_fixed[g.pp_i[k]] and g.pose_fixed[g.pp_j[k]]:
                continue
            edges.append((int(g.pp_seq[k]), 0, k))
        for k in range(g.n_pl):
            if g.pose_fixed[g.pl_pose[k]] and g.lm_fixed[g.pl_lm[k]]:
                continue
            edges.append((int(g.pl_seq[k]), 1, k))
        edges.sort()
        self.edges = edges
        act = {}
        for _, t, k in edges:
            if t == 0:
                act[(0, int(g.pp_i[k]))] = int(g.pose_id[g.pp_i[k]])
                act[(0, int(g.pp_j[k]))] = int(g.pose_id[g.pp_j[k]])
            else:
                act[(0, int(g.pl_pose[k]))] = int(g.pose_id[g.pl_pose[k]])
                act[(1, int(g.pl_lm[k]))] = int(g.lm_id[g.pl_lm[k]])
        verts = sorted(act.items(), key=lambda kv: kv[1])
        self.hidx = {}
        self.order = []
        off = 0
        self.offset = []
        for (kind, idx), _ in verts:
            fixed = g.pose_fixed[idx] if kind == 0 else g.lm_fixed[idx]
            if fixed:
                self.hidx[(kind, idx)] = -1
                continue
            self.hidx[(kind, idx)] = len(self.order)
            self.order.append((kind, idx))
            self.offset.append(off)
            off += 3 if kind == 0 else 2
        self.dim = off
        return len(edges) > 0 and len(self.order) > 0

    def block_pattern(self):
        """Ordered (row, col, nrows, ncols) list of BlockSolver::buildStructure (column-major, rows ascending)."""
        dims = [3 if k == 0 else 2 for k, _ in self.order]
        cols = [set([i]) for i in range(len(self.order))]
        for _, t, k in self.edges:
            i0, i1 = self._edge_hidx(t, k)
            if i0 >= 0 and i1 >= 0:
                cols[max(i0, i1)].add(min(i0, i1))
        out = []
        for c, rows in enumerate(cols):
            for r in sorted(rows):
                out.append((r, c, dims[r], dims[c]))
        return out

    def _edge_hidx(self, t, k):
        g = self.g
        if t == 0:
            return self.hidx[(0, int(g.pp_i[k]))], self.hidx[(0, int(g.pp_j[k]))]
        return self.hidx[(0, int(g.pl_pose[k]))], self.hidx[(1, int(g.pl_lm[k]))]

    def errors(self):
        g = self.g
        out = []
        for _, t, k in self.edges:
            if t == 0:
                out.append(pp_error(self.pose[g.pp_i[k]], self.pose[g.pp_j[k]], g.pp_z[k]))
            else:
                out.append(pl_error(self.pose[g.pl_pose[k]], self.lm[g.pl_lm[k]], g.pl_z[k]))
        return out

    def chi2(self):
        g = self.g
        c = cr = 0.0
        for (_, t, k), e in zip(self.edges, self.errors()):
            om = _full3(g.pp_info[k]) if t == 0 else _full2(g.pl_info[k])
            v = float(e @ om @ e)
            c += v
            if t == 0 and g.pp_phi[k] > 0:
                cr += dcs(g.pp_phi[k], v)[0]
            else:
                cr += v
        return c, cr

    def build_system(self, numeric=True):
        g = self.g
        H = np.zeros((self.dim, self.dim))
        b = np.zeros(self.dim)
        errs = self.errors()
        for (_, t, k), e in zip(self.edges, errs):
            i0, i1 = self._edge_hidx(t, k)
            if t == 0:
                A, B = pp_jacobians(self.pose[g.pp_i[k]], self.pose[g.pp_j[k]], g.pp_z[k])
                om = _full3(g.pp_info[k])
                w = 1.0
                if g.pp_phi[k] > 0:
                    w = dcs(g.pp_phi[k], float(e @ om @ e))[1]
                om = w * om
                d0 = d1 = 3
            else:
                pose, line = self.pose[g.pl_pose[k]], self.lm[g.pl_lm[k]]
                if numeric:
                    A, B = pl_jac_numeric(pose, line, g.pl_z[k], i0 >= 0, i1 >= 0)
                else:
                    A, B = pl_jac_analytic(pose, line)
                om = _full2(g.pl_info[k])
                d0, d1 = 3, 2
            omega_r = -(om @ e)
            if i0 >= 0:
                o0 = self.offset[i0]
                b[o0:o0 + d0] += A.T @ omega_r
                H[o0:o0 + d0, o0:o0 + d0] += A.T @ om @ A
            if i1 >= 0:
                o1 = self.offset[i1]
                b[o1:o1 + d1] += B.T @ omega_r
                H[o1:o1 + d1, o1:o1 + d1] += B.T @ om @ B
            if i0 >= 0 and i1 >= 0:
                blk = A.T @ om @ B
                H[o0:o0 + d0, o1:o1 + d1] += blk
                H[o1:o1 + d1, o0:o0 + d0] += blk.T
        return H, b

    def update(self, x):
        for (kind, idx), off in zip(self.order, self.offset):
            if kind == 0:
                self.pose[idx] = pose_oplus(self.pose[idx], x[off:off + 3])
            else:
                self.lm[idx] = lm_oplus(self.lm[idx], x[off:off + 2])

    def optimize(self, iters, algo="lm", numeric=True):
        stats = []
        for it in range(iters):
            current = self.chi2()[1]
            before = current
            H, b = self.build_system(numeric)
            if algo == "gn":
                try:
                    x = np.linalg.solve(H, b)
                    ok = bool(np.all(np.isfinite(x)))
                except np.linalg.LinAlgError:
                    ok = False
                    x = np.zeros_like(b)
                self.update(x)
                stats.append(dict(iteration=it, trials=1, result=1 if ok else -1, chi2=current, lambda_=0.0, rho=0.0,
                                  chi2_before=before))
                if not ok:
                    return 0, stats
                continue
            if it == 0:
                self.lam = 1e-5 * float(np.max(np.abs(np.diag(H))))
                self.ni = 2.0
            rho = 0.0
            trials = 0
            while True:
                bak = (self.pose.copy(), self.lm.copy())
                try:
                    x = np.linalg.solve(H + self.lam * np.eye(self.dim), b)
                    ok2 = bool(np.all(np.isfinite(x)))
                except np.linalg.LinAlgError:
                    ok2 = False
                    x = np.zeros_like(b)
                self.update(x)
                temp = self.chi2()[1]
                if not ok2:
                    temp = np.finfo(np.float64).max
                rho = current - temp
                scale = float(np.sum(x * (self.lam * x + b))) + 1e-3
                rho /= scale
                if rho > 0 and np.isfinite(temp):
                    alpha = 1.0 - (2 * rho - 1) ** 3
                    alpha = min(alpha, 2.0 / 3.0)
                    self.lam *= max(1.0 / 3.0, alpha)
                    self.ni = 2.0
                    current = temp
                else:
                    self.lam *= self.ni
                    self.ni *= 2
                    self.pose, self.lm = bak
                    if not np.isfinite(self.lam):
                        break
                trials += 1
                if not (rho < 0 and trials < 10):
                    break
            result = 2 if (trials == 10 or rho == 0 or not np.isfinite(self.lam)) else 1
            stats.append(dict(iteration=it, trials=trials, result=result, chi2=current, lambda_=self.lam, rho=rho,
                              chi2_before=before))
            if result != 1:
                break
        return len(stats), stats
