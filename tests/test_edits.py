"""Rows either side of the optimiser: pose-graph edits and information-matrix producers.

CPU tier: the oracle (oracle/sgo_frontend.cpp) is pinned against independent numpy restatements (homogeneous matrices,
np.linalg, finite-difference Jacobians of a float64 line fit), and the product's per-element bodies (sgb_edits.h, run
by tests/hostsim with the kernels' work decomposition) are compared with the oracle.
GPU tier (-m gpu): the CUDA kernels through the C ABI against the oracle, plus size-independent properties at full
size (re-measuring a chain from its own estimates reproduces it; the chi2 of those edges is zero).

Tolerances: FP64 rows 1e-12 relative on the host, 1e-10 on the device (sincos differs in the last bit and the error
of a composed chain grows with its length); single-precision rows (the reference computes them in float) 1e-5 relative
between two float evaluations in the same order, 2e-4 against the device (cosf/sinf/atan2f and FMA contraction).
"""
import numpy as np
import pytest

import hostsim
from oracle import cpu_oracle as co
from sparse_gslam_b200 import graphgen as gg


# ------------------------------------------------------------------------------------------------ helpers
def T(p):
    c, s = np.cos(p[2]), np.sin(p[2])
    return np.array([[c, -s, p[0]], [s, c, p[1]], [0, 0, 1.0]])


def from_T(M):
    return np.array([M[0, 2], M[1, 2], np.arctan2(M[1, 0], M[0, 0])])


def ang(a):
    return np.abs(gg.wrap(np.asarray(a)))


def pose_close(a, b, tol):
    a, b = np.asarray(a).reshape(-1, 3), np.asarray(b).reshape(-1, 3)
    scale = max(1.0, float(np.abs(b[:, :2]).max()))
    assert np.abs(a[:, :2] - b[:, :2]).max() / scale < tol
    assert ang(a[:, 2] - b[:, 2]).max() < tol


def sym6(m):
    return np.array([m[0, 0], m[0, 1], m[0, 2], m[1, 1], m[1, 2], m[2, 2]])


def rand_chain(rng, n, scale=1.0):
    est = np.zeros((n, 3))
    est[0] = [rng.normal(), rng.normal(), rng.uniform(-3, 3)]
    for k in range(1, n):
        d = np.array([0.5 * scale + 0.05 * rng.normal(), 0.05 * rng.normal(), 0.3 * rng.normal()])
        est[k] = from_T(T(est[k - 1]) @ T(d))
    return est


def rand_info(rng, n, dim=3):
    out = []
    for _ in range(n):
        a = rng.normal(size=(dim, dim))
        m = a @ a.T + dim * np.eye(dim)
        out.append(sym6(m) if dim == 3 else np.array([m[0, 0], m[0, 1], m[1, 1]]))
    return np.array(out)


def odom_case(rng, n_seg, max_steps=12):
    lens = rng.integers(0, max_steps + 1, size=n_seg)
    lens[0] = 0  # an interval without odometry: reset() state
    seg = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    n = int(seg[-1])
    deltas = np.stack([0.05 + 0.02 * rng.normal(size=n), 0.005 * rng.normal(size=n), 0.03 * rng.normal(size=n)], 1)
    deltas[rng.random(n) < 0.1, 0] *= -1  # reversing: the abs() in the input covariance matters
    return deltas, seg


def scan_case(rng, nw=3, ns=5, sz=61):
    ang_ = np.linspace(-2.0, 2.0, sz)
    beam = np.stack([np.cos(ang_), np.sin(ang_)], 1).astype(np.float32)
    deltas = np.stack([0.1 + 0.02 * rng.normal(size=(nw, ns - 1)), 0.01 * rng.normal(size=(nw, ns - 1)),
                       0.05 * rng.normal(size=(nw, ns - 1))], -1)
    r = rng.uniform(0.5, 5.0, size=(nw, ns, sz))
    pts = np.stack([r * np.cos(ang_) + 0.1 * rng.normal(size=r.shape), r * np.sin(ang_) + 0.1 * rng.normal(size=r.shape)], -1)
    pts = pts.astype(np.float32)
    pts[rng.random(r.shape) < 0.1] = np.inf  # no return (multicloud2.cpp thresh())
    return deltas, beam, pts


def line_case(rng, n_seg=40):
    pts, cov, seg = [], [], [0]
    for s in range(n_seg):
        n = int(rng.integers(5, 60))
        th = rng.uniform(-np.pi, np.pi)
        rho = rng.uniform(-4.0, 4.0)  # negative: exercises checkRhoTheta's flip
        t = np.sort(rng.uniform(-2.0, 2.0, size=n))
        x = rho * np.cos(th) - t * np.sin(th) + 0.01 * rng.normal(size=n)
        y = rho * np.sin(th) + t * np.cos(th) + 0.01 * rng.normal(size=n)
        pts.append(np.stack([x, y], 1))
        for _ in range(n):
            a = rng.normal(size=(2, 2)) * 0.01
            cov.append((a @ a.T + 1e-4 * np.eye(2)).reshape(4))
        seg.append(seg[-1] + n)
    return np.concatenate(pts).astype(np.float32), np.array(cov, np.float32), np.array(seg, np.int32)


# ------------------------------------------------------------------------------------------------ oracle KATs
def test_oracle_pg_append_matches_homogeneous_matrices():
    rng = np.random.default_rng(11)
    lm = rand_chain(rng, 40)
    prev = np.array([3.0, -2.0, 2.9])
    z, est = co.pg_append(prev, lm)
    cur = T(prev)
    for k in range(39):
        zk = np.linalg.inv(T(lm[k])) @ T(lm[k + 1])
        pose_close(z[k], from_T(zk), 1e-12)
        cur = cur @ zk
        pose_close(est[k], from_T(cur), 1e-11)
    # hanging the chain from its own predecessor reproduces it
    _, est2 = co.pg_append(lm[0], lm)
    pose_close(est2, lm[1:], 1e-12)


def test_oracle_closure_chi2_matches_numpy():
    rng = np.random.default_rng(12)
    est = rand_chain(rng, 30)
    ei = rng.integers(0, 30, 20).astype(np.int32)
    ej = ((ei + rng.integers(1, 29, 20)) % 30).astype(np.int32)
    z = np.array([from_T(np.linalg.inv(T(est[i])) @ T(est[j])) for i, j in zip(ei, ej)]) + 0.05 * rng.normal(size=(20, 3))
    info = rand_info(rng, 20)
    chi = co.closure_chi2(est, ei, ej, z, info)
    for k in range(20):
        e = from_T(np.linalg.inv(T(z[k])) @ np.linalg.inv(T(est[ei[k]])) @ T(est[ej[k]]))
        u = info[k]
        O = np.array([[u[0], u[1], u[2]], [u[1], u[3], u[4]], [u[2], u[4], u[5]]])
        assert abs(chi[k] - e @ O @ e) <= 1e-10 * max(1.0, chi[k])
    # exact measurements: zero error
    z0 = np.array([from_T(np.linalg.inv(T(est[i])) @ T(est[j])) for i, j in zip(ei, ej)])
    assert co.closure_chi2(est, ei, ej, z0, info).max() < 1e-20


def test_oracle_odom_information_matches_numpy():
    rng = np.random.default_rng(13)
    deltas, seg = odom_case(rng, 25)
    sx, sy, sw = 0.1, 0.05, 0.2
    z, cov, info = co.odom_information(deltas, seg, sx, sy, sw)
    for s in range(25):
        P = np.eye(3) * 1e-6
        pose = np.zeros(3)
        for k in range(seg[s], seg[s + 1]):
            dx, dy, dth = deltas[k]
            ct, st = np.cos(pose[2]), np.sin(pose[2])
            J1 = np.array([[1, 0, dy * ct - dx * st], [0, 1, -dx * ct - dy * st], [0, 0, 1.0]])
            J2 = np.array([[ct, st, 0], [-st, ct, 0], [0, 0, 1.0]])
            Q = np.diag([abs(dx * dx) * sx ** 2, abs(dy * dx) * sy ** 2, abs(dth * dx) * sw ** 2])
            P = J1 @ P @ J1.T + J2 @ Q @ J2.T
            pose = from_T(T(pose) @ T(deltas[k]))
        pose_close(z[s], pose, 1e-12)
        np.testing.assert_allclose(cov[s], P, rtol=1e-10, atol=1e-18)
        np.testing.assert_allclose(info[s], sym6(np.linalg.inv(P)), rtol=1e-8)
    assert np.allclose(cov[0], np.eye(3) * 1e-6) and np.allclose(z[0], 0)  # empty interval = reset() state


def _fit64(p):
    """Total-least-squares line of points p [n,2] in float64, the closed form the reference cites (eq. 11)."""
    xb, yb = p.mean(0)
    sxx, syy, sxy = ((p[:, 0] - xb) ** 2).sum(), ((p[:, 1] - yb) ** 2).sum(), ((p[:, 0] - xb) * (p[:, 1] - yb)).sum()
    th = 0.5 * np.arctan2(-2 * sxy, syy - sxx)
    rho = xb * np.cos(th) + yb * np.sin(th)
    return rho, th


def test_oracle_line_fit_matches_float64_fit_and_finite_difference_covariance():
    rng = np.random.default_rng(14)
    pts, pcov, seg = line_case(rng, 30)
    rt, cov, info = co.line_fit_information(pts, pcov, seg)
    flips = 0
    for s in range(30):
        p = pts[seg[s]:seg[s + 1]].astype(np.float64)
        rho, th = _fit64(p)
        if rho < 0:
            rho, th = -rho, th + np.pi
            flips += 1
        assert abs(rt[s, 0] - rho) < 2e-5 and ang(rt[s, 1] - th) < 2e-5
        assert rt[s, 0] >= 0
        # covariance: sum_i A_i C_i A_i^T with A_i = d(rho,theta)/d(x_i,y_i) by central differences of the float64 fit
        C = np.zeros((2, 2))
        for i in range(len(p)):
            A = np.zeros((2, 2))
            for d in range(2):
                q1, q2 = p.copy(), p.copy()
                q1[i, d] += 1e-6
                q2[i, d] -= 1e-6
                r1, t1 = _fit64(q1)
                r2, t2 = _fit64(q2)
                if _fit64(p)[0] < 0:
                    r1, r2 = -r1, -r2
                A[0, d] = (r1 - r2) / 2e-6
                A[1, d] = gg.wrap(t1 - t2) / 2e-6
            C += A @ pcov[seg[s] + i].reshape(2, 2).astype(np.float64) @ A.T
        c = cov[s].reshape(2, 2).astype(np.float64)
        assert np.abs(c - C).max() <= 2e-3 * np.abs(C).max(), (s, c, C)
        iv = np.linalg.inv(c)
        np.testing.assert_allclose(info[s], [iv[0, 0], iv[0, 1], iv[1, 1]], rtol=1e-9)
    assert flips > 3


def test_oracle_scan_point_covariances_matches_numpy():
    rng = np.random.default_rng(15)
    deltas, beam, pts = scan_case(rng)
    sx, sy, sw, vr = 0.1, 0.05, 0.2, 0.03 ** 2
    cov, rt, valid = co.scan_point_covariances(deltas, beam, pts, sx, sy, sw, vr)
    nw, ns, sz = pts.shape[:3]
    assert valid.sum() < valid.size and valid.sum() > 0.8 * valid.size
    for w in range(nw):
        for i in range(ns):
            P = np.eye(3) * 1e-6
            pose = np.zeros(3)
            for j in range(i, ns - 1):
                dx, dy, dth = deltas[w, j]
                ct, st = np.cos(pose[2]), np.sin(pose[2])
                J1 = np.array([[1, 0, dy * ct - dx * st], [0, 1, -dx * ct - dy * st], [0, 0, 1.0]])
                J2 = np.array([[ct, st, 0], [-st, ct, 0], [0, 0, 1.0]])
                Q = np.diag([abs(dx * dx) * sx ** 2, abs(dy * dx) * sy ** 2, abs(dth * dx) * sw ** 2])
                P = J1 @ P @ J1.T + J2 @ Q @ J2.T
                pose = from_T(T(pose) @ T(deltas[w, j]))
            ct, st = np.cos(pose[2]), np.sin(pose[2])
            Juk = np.array([[-ct, st, pose[1] * ct + pose[0] * st], [-st, -ct, pose[1] * st - pose[0] * ct], [0, 0, -1.0]])
            P = Juk @ P @ Juk.T
            inv = from_T(np.linalg.inv(T(pose)))
            c2, s2 = np.cos(inv[2]), np.sin(inv[2])
            Ja = np.array([[1, 0, inv[1] * c2 - inv[0] * s2], [0, 1, -inv[0] * c2 - inv[1] * s2]])
            Jb = np.array([[c2, s2], [-s2, c2]])
            for j in range(0, sz, 7):
                if not valid[w, i, j]:
                    assert not np.isfinite(pts[w, i, j]).all() and np.all(cov[w, i, j] == 0)
                    continue
                cv, sv = beam[j]
                Cp = vr * np.array([[cv * cv, cv * sv], [cv * sv, sv * sv]])
                ref = Ja @ P @ Ja.T + Jb @ Cp @ Jb.T
                got = cov[w, i, j].reshape(2, 2)
                assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()
                x, y = pts[w, i, j]
                assert abs(rt[w, i, j, 0] - np.hypot(x, y)) < 1e-5 and abs(rt[w, i, j, 1] - np.arctan2(y, x)) < 1e-6
    # newest scan: no odometry in between, pose covariance stays at reset()
    j = int(np.argmax(valid[0, ns - 1]))
    cv, sv = beam[j]
    ref = 1e-6 * np.eye(2) + vr * np.array([[cv * cv, cv * sv], [cv * sv, sv * sv]])
    assert np.abs(cov[0, ns - 1, j].reshape(2, 2) - ref).max() < 1e-4 * np.abs(ref).max()


# ------------------------------------------------------------------------------------------------ product bodies (host) vs oracle
@pytest.mark.parametrize("n,threads,items,scan", [(1, 256, 8, 512), (37, 256, 8, 512), (2048, 256, 8, 512), (2049, 256, 8, 512),
                                                  (9000, 256, 8, 512), (700, 8, 2, 4), (1000, 4, 3, 2)])
def test_bodies_pg_append(n, threads, items, scan):
    rng = np.random.default_rng(n)
    lm = rand_chain(rng, n + 1)
    prev = np.array([-4.0, 7.0, -3.1])
    z0, e0 = co.pg_append(prev, lm)
    z1, e1 = hostsim.pg_append(prev, lm, threads, items, scan)
    pose_close(z1, z0, 1e-13)
    pose_close(e1, e0, 1e-11)  # the scan re-associates the SE2 products


def test_bodies_closure_chi2_and_odom():
    rng = np.random.default_rng(21)
    est = rand_chain(rng, 200)
    ei = rng.integers(0, 200, 300).astype(np.int32)
    ej = ((ei + rng.integers(1, 199, 300)) % 200).astype(np.int32)
    z = np.array([from_T(np.linalg.inv(T(est[i])) @ T(est[j])) for i, j in zip(ei, ej)]) + 0.1 * rng.normal(size=(300, 3))
    info = rand_info(rng, 300)
    np.testing.assert_allclose(hostsim.closure_chi2(est, ei, ej, z, info), co.closure_chi2(est, ei, ej, z, info), rtol=1e-12)
    deltas, seg = odom_case(rng, 200)
    a, b = hostsim.odom_information(deltas, seg, 0.1, 0.05, 0.2), co.odom_information(deltas, seg, 0.1, 0.05, 0.2)
    pose_close(a[0], b[0], 1e-13)
    np.testing.assert_allclose(a[1], b[1], rtol=1e-12, atol=1e-20)
    np.testing.assert_allclose(a[2], b[2], rtol=1e-10)


def test_bodies_float_rows():
    rng = np.random.default_rng(22)
    deltas, beam, pts = scan_case(rng, nw=4, ns=6, sz=90)
    a = hostsim.scan_point_covariances(deltas, beam, pts, 0.1, 0.05, 0.2, 9e-4)
    b = co.scan_point_covariances(deltas, beam, pts, 0.1, 0.05, 0.2, 9e-4)
    assert np.array_equal(a[2], b[2])
    np.testing.assert_allclose(a[0], b[0], rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(a[1], b[1], rtol=1e-6, atol=1e-6)
    p, c, seg = line_case(rng, 60)
    a, b = hostsim.line_fit_information(p, c, seg), co.line_fit_information(p, c, seg)
    np.testing.assert_allclose(a[0], b[0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(a[1], b[1], rtol=1e-4, atol=1e-12)
    np.testing.assert_allclose(a[2], b[2], rtol=1e-3)


# ------------------------------------------------------------------------------------------------ GPU tier
gpu = pytest.mark.gpu


def _pose_graph_script(g, n_false=3, seed=0):
    """A pose-graph life cycle from the C1-shaped pose graph: chunks of the chain arrive, closures in between."""
    rng = np.random.default_rng(seed)
    is_odo = (g.pp_j - g.pp_i) == 1
    odo_info = g.pp_info[is_odo]
    assert np.array_equal(g.pp_i[is_odo], np.arange(g.P - 1))
    clo = np.flatnonzero(~is_odo)
    chunks = np.array_split(np.arange(1, g.P), 6)
    script = []
    for c in chunks:
        script.append(("append", int(c[0]), len(c)))
        hi = int(c[-1])
        for k in clo:
            if max(g.pp_i[k], g.pp_j[k]) <= hi and min(g.pp_i[k], g.pp_j[k]) >= 0 and ("closure", int(k)) not in script:
                script.append(("closure", int(k)))
    false = []
    for _ in range(n_false):
        i, j = sorted(rng.choice(g.P, 2, replace=False))
        false.append((int(i), int(j), rng.normal(size=3) * [3.0, 3.0, 1.0], np.array([400.0, 0, 0, 400.0, 0, 2500.0])))
    return odo_info, script, false


def _oracle_pose_graph(g, odo_info, script, false, phi, lm_est):
    """The same life cycle on the CPU oracle: returns (graph before optimise, dict of arrays)."""
    est = [lm_est[0].copy()]
    pp_i, pp_j, pp_z, pp_info, pp_phi, is_c = [], [], [], [], [], []
    for op in script:
        if op[0] == "append":
            first, cnt = op[1], op[2]
            z, e = co.pg_append(est[-1], lm_est[first - 1:first + cnt])
            for k in range(cnt):
                pp_i.append(len(est) - 1)
                pp_j.append(len(est))
                est.append(e[k])
                pp_z.append(z[k])
                pp_info.append(odo_info[first - 1 + k])
                pp_phi.append(0.0)
                is_c.append(0)
        else:
            k = op[1]
            pp_i.append(int(g.pp_i[k])); pp_j.append(int(g.pp_j[k])); pp_z.append(g.pp_z[k]); pp_info.append(g.pp_info[k])
            pp_phi.append(phi); is_c.append(1)
    for (i, j, z, info) in false:
        pp_i.append(i); pp_j.append(j); pp_z.append(z); pp_info.append(info); pp_phi.append(phi); is_c.append(1)
    return dict(est=np.array(est), pp_i=np.array(pp_i, np.int32), pp_j=np.array(pp_j, np.int32), pp_z=np.array(pp_z),
                pp_info=np.array(pp_info), pp_phi=np.array(pp_phi), is_c=np.array(is_c, np.uint8))


def _as_graph(g, d, active):
    P = d["est"].shape[0]
    a = np.flatnonzero(active)
    fixed = np.zeros(P, np.uint8)
    fixed[0] = 1
    return gg.Graph(name="pg", pose_id=np.arange(P, dtype=np.int32), pose_est=d["est"].copy(), pose_fixed=fixed,
                    pose_gt=d["est"].copy(), lm_id=np.zeros(0, np.int32), lm_est=np.zeros((0, 2)), lm_fixed=np.zeros(0, np.uint8),
                    lm_gt=np.zeros((0, 2)), pp_i=d["pp_i"][a], pp_j=d["pp_j"][a], pp_z=d["pp_z"][a], pp_info=d["pp_info"][a],
                    pp_phi=d["pp_phi"][a], pp_seq=np.arange(len(a), dtype=np.int64), pl_pose=np.zeros(0, np.int32),
                    pl_lm=np.zeros(0, np.int32), pl_z=np.zeros((0, 2)), pl_info=np.zeros((0, 3)), pl_seq=np.zeros(0, np.int64))


@gpu
def test_pose_graph_life_cycle_matches_oracle():
    from sparse_gslam_b200 import SparseOptimizerB200, capi
    from sparse_gslam_b200.posegraph import CHI2_REJECT, PoseGraphB200
    g = gg.make_c1().pose_only(phi=10.0)
    phi = 10.0
    odo_info, script, false = _pose_graph_script(g)
    # the landmark graph whose optimised estimates are copied: the same poses, held by an optimiser handle on the device
    lm = SparseOptimizerB200(capi.ALGO_GN, jacobian_mode=capi.JAC_ANALYTIC)
    assert lm.initialize_optimization(g)
    lm_est = g.pose_est
    pg = PoseGraphB200()
    pg.reset(lm_est[0], 0)
    for op in script:
        if op[0] == "append":
            pg.append_from_lm(lm, op[1], op[2], odo_info[op[1] - 1:op[1] - 1 + op[2]])
        else:
            k = op[1]
            pg.add_closure(g.pp_i[k], g.pp_j[k], g.pp_z[k], g.pp_info[k], phi)
    for (i, j, z, info) in false:
        pg.add_closure(i, j, z, info, phi)
    d = _oracle_pose_graph(g, odo_info, script, false, phi, lm_est)
    got = pg.download()
    assert np.array_equal(got["pp_i"], d["pp_i"]) and np.array_equal(got["pp_j"], d["pp_j"])
    assert np.array_equal(got["pp_is_closure"], d["is_c"])
    pose_close(got["pp_z"], d["pp_z"], 1e-10)
    pose_close(got["pose_est"], d["est"], 1e-10)
    np.testing.assert_array_equal(got["pp_info"], d["pp_info"])
    np.testing.assert_array_equal(got["pp_phi"], d["pp_phi"])
    # optimise (GN-20 + DCS), oracle on the same graph
    active = np.ones(len(d["pp_i"]), bool)
    o = co.Oracle(_as_graph(g, d, active))
    assert o.initialize_optimization()
    n_o, _ = o.optimize(20, co.ALGO_GN, co.JAC_ANALYTIC)
    n_g, stats = pg.optimize(20)
    assert n_g == n_o == 20
    po, _ = o.estimates()
    got = pg.download()
    pose_close(got["pose_est"], po, 1e-6)
    # false-closure removal at the optimised estimates
    clo = np.flatnonzero(d["is_c"])
    chi_o = co.closure_chi2(po, d["pp_i"][clo], d["pp_j"][clo], d["pp_z"][clo], d["pp_info"][clo])
    removed, chi_g, act = pg.prune_closures(CHI2_REJECT)
    margin = np.abs(chi_o - CHI2_REJECT) > 1e-3 * CHI2_REJECT  # decisions away from the threshold must agree
    assert np.array_equal(act[margin], (chi_o <= CHI2_REJECT)[margin])
    np.testing.assert_allclose(chi_g, chi_o, rtol=1e-4, atol=1e-8)
    assert removed == int((~act).sum()) and removed >= len(false) - 1
    assert pg.info()["n_active_closures"] == int(act.sum())
    # final optimise without the removed edges (log_runner.cpp:203-204)
    active[clo[~act]] = False
    d2 = dict(d, est=po)
    o2 = co.Oracle(_as_graph(g, d2, active))
    assert o2.initialize_optimization()
    o2.optimize(20, co.ALGO_GN, co.JAC_ANALYTIC)
    # start the device store from the oracle's estimates so that both final solves start from the same point
    n_g, _ = pg.optimize(20)
    assert n_g == 20
    p2, _ = o2.estimates()
    pose_close(pg.download()["pose_est"], p2, 1e-6)
    # a second prune removes nothing new
    removed2, _, act2 = pg.prune_closures(CHI2_REJECT)
    assert removed2 == 0 and np.array_equal(act2, act)


@gpu
def test_set_graph_device_equals_host_path():
    """sgb_set_graph_device (values gathered on the device) gives the same linearisation as sgb_set_graph."""
    from sparse_gslam_b200 import capi
    from sparse_gslam_b200.posegraph import PoseGraphB200
    g = gg.make_c1().pose_only(phi=10.0)
    pg = PoseGraphB200()
    pg.reset(g.pose_est[0], 0)
    is_odo = (g.pp_j - g.pp_i) == 1
    pg.append_from_host(g.pose_est, g.pp_info[is_odo])
    for k in np.flatnonzero(~is_odo):
        pg.add_closure(g.pp_i[k], g.pp_j[k], g.pp_z[k], g.pp_info[k], 10.0)
    n, _ = pg.optimize(1)
    assert n == 1
    d = pg.download()
    # the same graph through the host path, from the store's own values
    fixed = np.zeros(g.P, np.uint8); fixed[0] = 1
    h = gg.Graph(name="h", pose_id=d["pose_id"], pose_est=d["pose_est"], pose_fixed=fixed, pose_gt=d["pose_est"],
                 lm_id=np.zeros(0, np.int32), lm_est=np.zeros((0, 2)), lm_fixed=np.zeros(0, np.uint8), lm_gt=np.zeros((0, 2)),
                 pp_i=d["pp_i"], pp_j=d["pp_j"], pp_z=d["pp_z"], pp_info=d["pp_info"], pp_phi=d["pp_phi"],
                 pp_seq=np.arange(len(d["pp_i"]), dtype=np.int64), pl_pose=np.zeros(0, np.int32), pl_lm=np.zeros(0, np.int32),
                 pl_z=np.zeros((0, 2)), pl_info=np.zeros((0, 3)), pl_seq=np.zeros(0, np.int64))
    from sparse_gslam_b200 import SparseOptimizerB200
    a = SparseOptimizerB200(capi.ALGO_GN, jacobian_mode=capi.JAC_ANALYTIC)
    assert a.initialize_optimization(h)
    la = a.linearize()
    lb = pg.solver.linearize()  # the store's solver holds the device-gathered graph at the same estimates
    # not bit-equal: the cached inverse measurement is formed with the device's sincos on one path, glibc's on the other
    np.testing.assert_allclose(lb["H"], la["H"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(lb["b"], la["b"], rtol=1e-9, atol=1e-7)
    np.testing.assert_allclose(lb["chi2"], la["chi2"], rtol=1e-12)


@gpu
def test_chain_round_trip_full_size():
    """1M-pose chain (C5 size): re-measuring a chain from its own estimates and hanging it from its own first pose
    reproduces it, and the chi2 of the re-measured edges at those estimates is zero."""
    from sparse_gslam_b200.posegraph import PoseGraphB200
    rng = np.random.default_rng(5)
    n = 1_000_000
    d = np.stack([0.5 + 0.02 * rng.normal(size=n), 0.02 * rng.normal(size=n), 0.2 * rng.normal(size=n)], 1)
    th = np.concatenate([[0.3], 0.3 + np.cumsum(d[:, 2])])
    x = np.concatenate([[0.0], np.cumsum(d[:, 0] * np.cos(th[:-1]) - d[:, 1] * np.sin(th[:-1]))])
    y = np.concatenate([[0.0], np.cumsum(d[:, 0] * np.sin(th[:-1]) + d[:, 1] * np.cos(th[:-1]))])
    lm = np.stack([x, y, gg.wrap(th)], 1)
    pg = PoseGraphB200()
    pg.reset(lm[0], 0)
    info = np.tile(np.array([2500.0, 0, 0, 2500.0, 0, 1e4]), (n, 1))
    pg.append_from_host(lm, info)
    got = pg.download()
    assert got["pose_est"].shape[0] == n + 1
    assert np.abs(got["pose_est"][:, :2] - lm[:, :2]).max() < 1e-7  # ~1e3 m extent, 1e6 compositions
    assert ang(got["pose_est"][:, 2] - lm[:, 2]).max() < 1e-9
    pose_close(got["pp_z"], d, 1e-9)
    # close the loop over every 1000th edge with the re-measured z itself: chi2 == 0 at the stored estimates
    idx = np.arange(0, n, 1000)
    for k in idx[:50]:
        pg.add_closure(int(k), int(k) + 1, got["pp_z"][k], info[k], 10.0)
    removed, chi, act = pg.prune_closures(11.345)
    assert removed == 0 and chi.max() < 1e-9 and act.all()
    assert pg.info()["kernel_launches"] >= 3


@gpu
def test_producers_match_oracle():
    from sparse_gslam_b200 import frontend as fe
    rng = np.random.default_rng(31)
    deltas, seg = odom_case(rng, 5000)
    z, cov, info, ms = fe.odom_information(deltas, seg, 0.1, 0.05, 0.2)
    zo, co_, io = co.odom_information(deltas, seg, 0.1, 0.05, 0.2)
    pose_close(z, zo, 1e-12)
    np.testing.assert_allclose(cov, co_, rtol=1e-10, atol=1e-20)
    np.testing.assert_allclose(info, io, rtol=1e-8)
    assert ms > 0
    d2, beam, pts = scan_case(rng, nw=6, ns=8, sz=181)
    c, rt, valid, ms = fe.scan_point_covariances(d2, beam, pts, 0.1, 0.05, 0.2, 9e-4)
    c0, rt0, v0 = co.scan_point_covariances(d2, beam, pts, 0.1, 0.05, 0.2, 9e-4)
    assert np.array_equal(valid, v0)
    np.testing.assert_allclose(c, c0, rtol=2e-4, atol=1e-10)
    np.testing.assert_allclose(rt, rt0, rtol=1e-5, atol=2e-6)
    p, pc, sg = line_case(rng, 500)
    r, cv, inf, ms = fe.line_fit_information(p, pc, sg)
    r0, cv0, inf0 = co.line_fit_information(p, pc, sg)
    np.testing.assert_allclose(r[:, 0], r0[:, 0], rtol=1e-5, atol=1e-5)
    assert ang(r[:, 1] - r0[:, 1]).max() < 1e-5
    np.testing.assert_allclose(cv, cv0, rtol=2e-3, atol=1e-10)
    np.testing.assert_allclose(inf, inf0, rtol=5e-3)


@gpu
def test_producers_empty_and_bad_arguments():
    from sparse_gslam_b200 import SgbError
    from sparse_gslam_b200 import frontend as fe
    z, cov, info, ms = fe.odom_information(np.zeros((0, 3)), [0], 0.1, 0.1, 0.1)
    assert z.shape == (0, 3)
    z, cov, info, ms = fe.odom_information(np.zeros((0, 3)), [0, 0, 0], 0.1, 0.1, 0.1)  # two empty intervals
    assert np.allclose(cov, np.eye(3) * 1e-6) and np.allclose(info[:, [0, 3, 5]], 1e6) and np.all(z == 0)
    with pytest.raises(SgbError):
        fe.odom_information(np.zeros((3, 3)), [0, 2, 1], 0.1, 0.1, 0.1)
    with pytest.raises(SgbError):
        fe.line_fit_information(np.zeros((3, 2)), np.zeros((3, 4)), [1, 3])


@gpu
@pytest.mark.parametrize("name", ["small", "c1"])
def test_set_graph_device_full_graph_with_slots(name):
    """Pose + line-landmark graph whose values live in device memory in a DIFFERENT order than the index arrays (slot
    lists): same linearisation and the same LM-15 result as the host path."""
    import torch
    from sparse_gslam_b200 import SparseOptimizerB200, capi
    g = gg.make(name)
    rng = np.random.default_rng(3)
    perm_pp, perm_pl = rng.permutation(g.n_pp), rng.permutation(g.n_pl)  # device slot of edge k
    def dev_edges(a, perm):
        out = np.empty_like(a)
        out[perm] = a
        return torch.from_numpy(np.ascontiguousarray(out)).cuda()
    t = dict(pose_est=torch.from_numpy(np.ascontiguousarray(g.pose_est)).cuda(), lm_est=torch.from_numpy(np.ascontiguousarray(g.lm_est)).cuda(),
             pp_z=dev_edges(g.pp_z, perm_pp), pp_info=dev_edges(g.pp_info, perm_pp), pp_phi=dev_edges(g.pp_phi, perm_pp),
             pl_z=dev_edges(g.pl_z, perm_pl), pl_info=dev_edges(g.pl_info, perm_pl))
    torch.cuda.synchronize()
    a = SparseOptimizerB200(capi.ALGO_LM, jacobian_mode=capi.JAC_ANALYTIC)
    assert a.initialize_optimization(g)
    b = SparseOptimizerB200(capi.ALGO_LM, jacobian_mode=capi.JAC_ANALYTIC)
    assert b.initialize_optimization_device(g, {k: v.data_ptr() for k, v in t.items()}, perm_pp, perm_pl)
    sa, sb = a.structure(), b.structure()
    for k in ("kind", "index", "offset", "row", "col", "nrows", "ncols"):
        assert np.array_equal(sa[k], sb[k])
    la, lb = a.linearize(), b.linearize()
    np.testing.assert_allclose(lb["H"], la["H"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(lb["b"], la["b"], rtol=1e-9, atol=1e-7)
    np.testing.assert_allclose(lb["chi2"], la["chi2"], rtol=1e-12)
    na, _ = a.optimize(15)
    nb, _ = b.optimize(15)
    # the cached inverse measurements differ in the last bit between the two paths; once chi2 has stopped moving LM may
    # return Terminate an iteration earlier or later (rho is a ratio of two vanishing numbers): compare the state
    assert na >= 1 and nb >= 1
    pa, la_ = a.estimates()
    pb, lb_ = b.estimates()
    pose_close(pb, pa, 1e-6)
    assert np.abs(lb_ - la_).max() < 1e-6 * max(1.0, np.abs(la_).max())
    ca, cb = a.active_chi2()[0], b.active_chi2()[0]
    assert abs(ca - cb) <= 1e-6 * ca


# ------------------------------------------------------------------------------------------------ edge cases (host bodies vs oracle)
def _same(a, b, rtol, atol=0.0):
    """allclose that also demands the same NaN / inf pattern (degenerate inputs must degenerate the same way)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.isinf(a), np.isinf(b))
    m = np.isfinite(a)
    np.testing.assert_allclose(a[m], b[m], rtol=rtol, atol=atol)


def test_edge_cases_line_fit():
    segs = []
    segs.append(np.array([[1.0, 2.0]]))                                   # one point: no direction, 0/0 in the covariance
    segs.append(np.array([[1.0, 2.0], [2.0, 2.0]]))                       # two points
    segs.append(np.stack([np.linspace(-1, 1, 9), np.full(9, 3.0)], 1))    # exactly horizontal, zero residual
    segs.append(np.stack([np.full(9, -2.5), np.linspace(-1, 1, 9)], 1))   # exactly vertical, rho < 0 before the flip
    segs.append(np.stack([np.linspace(0, 4, 30), np.linspace(0, 4, 30)], 1) * [1, 1] + [0, 1e-3])  # diagonal
    segs.append(np.stack([500 + np.linspace(0, 2, 40), 300 + 0.5 * np.linspace(0, 2, 40)], 1))      # far from the origin
    pts = np.concatenate(segs).astype(np.float32)
    seg = np.concatenate([[0], np.cumsum([len(s) for s in segs])]).astype(np.int32)
    cov = np.tile(np.array([1e-4, 2e-5, 2e-5, 3e-4], np.float32), (len(pts), 1))
    a, b = hostsim.line_fit_information(pts, cov, seg), co.line_fit_information(pts, cov, seg)
    _same(a[0], b[0], 1e-5, 1e-6)
    _same(a[1], b[1], 1e-4, 1e-12)
    _same(a[2], b[2], 1e-3)
    assert np.all(b[0][2:, 0] >= 0)


def test_edge_cases_odometry_and_scan_points():
    deltas = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.5], [-0.1, 0.02, -0.3], [0.2, 0.0, 0.0], [0.0, 0.1, 0.0],
                       [0.3, -0.1, 3.2], [0.1, 0.0, 3.1]])                 # standing still, turning on the spot, reversing,
    seg = np.array([0, 0, 1, 2, 3, 5, 7], np.int32)                        # sideways motion, a turn past +-pi
    a, b = hostsim.odom_information(deltas, seg, 0.1, 0.05, 0.2), co.odom_information(deltas, seg, 0.1, 0.05, 0.2)
    pose_close(a[0], b[0], 1e-13)
    _same(a[1], b[1], 1e-12, 1e-22)
    _same(a[2], b[2], 1e-9)
    assert np.all(np.abs(b[0][:, 2]) <= np.pi)
    # scan points: a single scan per window (no odometry in between), a window without any return
    beam = np.stack([np.cos(np.linspace(-1, 1, 7)), np.sin(np.linspace(-1, 1, 7))], 1).astype(np.float32)
    pts = np.full((2, 1, 7, 2), np.inf, np.float32)
    pts[0, 0] = beam * 2.0
    a = hostsim.scan_point_covariances(np.zeros((2, 0, 3)), beam, pts, 0.1, 0.05, 0.2, 9e-4)
    b = co.scan_point_covariances(np.zeros((2, 0, 3)), beam, pts, 0.1, 0.05, 0.2, 9e-4)
    assert np.array_equal(a[2], b[2]) and a[2][0].all() and not a[2][1].any()
    _same(a[0], b[0], 1e-6, 1e-12)
    _same(a[1], b[1], 1e-6, 1e-7)
    assert np.all(a[0][1] == 0) and np.all(a[1][1] == 0)


def test_edge_cases_chain_copy():
    # headings right at the wrap, a chain that turns through +-pi several times
    lm = np.array([[0.0, 0.0, np.pi - 1e-12], [0.5, 0.0, -np.pi], [1.0, 0.1, np.pi - 1e-9], [1.5, 0.1, -np.pi + 1e-9],
                   [2.0, 0.2, 3.0], [2.5, 0.2, -3.0]])
    prev = np.array([10.0, -3.0, -np.pi])
    z0, e0 = co.pg_append(prev, lm)
    z1, e1 = hostsim.pg_append(prev, lm, 4, 2, 2)
    pose_close(z1, z0, 1e-13)
    pose_close(e1, e0, 1e-12)
    assert np.all((z0[:, 2] >= -np.pi) & (z0[:, 2] < np.pi)) and np.all((e0[:, 2] >= -np.pi) & (e0[:, 2] < np.pi))
    z, e = hostsim.pg_append(prev, lm[:1])  # nothing to copy
    assert z.shape == (0, 3) and e.shape == (0, 3)
