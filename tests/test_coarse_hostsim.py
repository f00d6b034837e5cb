"""Two-level preconditioner of the resident solve (sparse-gslam_b200/csrc/sgb_coarse.h) on the host harness: the gather
lists of the coarse matrix (plan_coarse), the one-CTA factorisation (coarse_factor -- the very code k_setup_coarse runs,
with one "thread"), and its use inside the PCG. The reference has no counterpart (LinearSolverEigen factorises exactly):
what is checked is that the solve keeps its contract -- same x, same LM trajectory as the oracle -- in fewer iterations."""
import numpy as np
import pytest

import hostsim
from oracle.cpu_oracle import ALGO_LM, JAC_ANALYTIC, Oracle
from sparse_gslam_b200 import capi
from sparse_gslam_b200 import graphgen as gg


@pytest.fixture(autouse=True)
def _coarse_off_afterwards():
    yield
    hostsim.use_coarse(0)


def chain_prefix(g, n):
    return g.chain_prefix(n)


def hat_restriction(nP, h, nn):
    R = np.zeros((3 * nn, 3 * nP))
    for i in range(nP):
        wr = (i % h) / h
        for c in range(3):
            R[3 * (i // h) + c, 3 * i + c] = 1.0 - wr
            R[3 * (i // h + 1) + c, 3 * i + c] += wr
    return R


@pytest.mark.parametrize("P,lam", [(63, 1e-3), (63, 10.0), (82, 0.5)])
def test_coarse_inverse_is_the_inverse_of_the_galerkin_matrix(P, lam):
    """plan_coarse + coarse_factor against numpy: R S R^T from the oracle's dense Hessian, hat functions every 8 rows.
    P = 82 -> 81 free poses: row 80 sits exactly on node 10, node 11 is reached by no row and must come out as an
    identity block."""
    g = gg.make_small(seed=5, P=P, L=12, E_l=3 * P, n_closures=6)
    o = Oracle(g)
    assert o.initialize_optimization()
    st = o.structure()
    H = o.dense_hessian(o.linearize(JAC_ANALYTIC), st)
    nP = int((st["kind"] == 0).sum())
    n3 = 3 * nP
    Hd = H + lam * np.eye(H.shape[0])
    S = Hd[:n3, :n3] - Hd[:n3, n3:] @ np.linalg.inv(Hd[n3:, n3:]) @ Hd[n3:, :n3]
    hostsim.use_coarse(40)
    hs = hostsim.HostSim(g, jac_numeric=False)
    flag, x, iters, rel = hs.solve_once(lam)
    h, nn, failed, Ainv = hs.coarse()
    assert flag == 0 and h == 8 and nn == (nP + 7) // 8 + 1 and not failed
    R = hat_restriction(nP, h, nn)
    A = R @ S @ R.T
    dead = np.where(np.abs(R).sum(axis=1) == 0)[0]
    assert (len(dead) == 3) == (nP % 8 == 1)
    A[dead, dead] = 1.0
    np.testing.assert_allclose(Ainv, Ainv.T, rtol=0, atol=0)  # built as X^T X, mirrored
    assert np.linalg.eigvalsh(Ainv).min() > 0
    err = np.abs(A @ Ainv - np.eye(3 * nn)).max()
    assert err < 1e-7 * np.linalg.cond(A) ** 0.5, err
    # and the step is the one the exact solver finds
    ok, xo = o.solve_once(lam, JAC_ANALYTIC)
    assert ok
    np.testing.assert_allclose(x, xo, rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("n", [100, 300, 400])
def test_two_level_solve_same_step_in_fewer_iterations(n):
    g = chain_prefix(gg.make("c1"), n)
    out = {}
    for nodes in (0, 40):
        hostsim.use_coarse(nodes)
        hs = hostsim.HostSim(g, jac_numeric=False)
        out[nodes] = hs.solve_once(1e-2)
        assert out[nodes][0] == 0
    (_, x0, it0, _), (_, x1, it1, _) = out[0], out[40]
    np.testing.assert_allclose(x1, x0, rtol=1e-6, atol=1e-8 * np.abs(x0).max())
    assert it1 <= 0.5 * it0, (it0, it1)


def test_two_level_lm_matches_oracle_and_plain_solve():
    g = chain_prefix(gg.make("c1"), 200)
    o = Oracle(g)
    assert o.initialize_optimization()
    n0, s0 = o.optimize(15, ALGO_LM, JAC_ANALYTIC)
    runs = {}
    for nodes in (0, 40):
        hostsim.use_coarse(nodes)
        hs = hostsim.HostSim(g, jac_numeric=False)
        n, stats = hs.optimize(15, capi.ALGO_LM)
        runs[nodes] = (n, stats, hs.estimates())
    n, stats, (ph, lh) = runs[40]
    assert n == n0 == runs[0][0]
    prev = None
    for a, b in zip(s0, stats):
        if prev is None or prev - a["chi2"] > 1e-9 * prev:   # trials of a converged iteration are rounding noise
            assert a["trials"] == b["trials"]
        np.testing.assert_allclose(b["chi2"], a["chi2"], rtol=1e-8)
        prev = a["chi2"]
    po, lo = o.estimates()
    np.testing.assert_allclose(ph, po, atol=1e-6)
    np.testing.assert_allclose(lh, lo, atol=1e-6)
    it_plain = sum(s["pcg_iters"] for s in runs[0][1])
    it_two = sum(s["pcg_iters"] for s in stats)
    assert it_two <= 0.4 * it_plain, (it_plain, it_two)


def test_spacing_rule_and_switch_off():
    """coarse_spacing: the smallest of 8 / 16 / 32 / 64 rows per node that needs at most 40 nodes; graphs of a handful of
    rows and graphs too long for a dense coarse solve keep the plain block-Jacobi preconditioner."""
    c1 = gg.make("c1")
    for n, want in ((11, 0), (25, 8), (313, 8), (314, 16), (625, 16), (626, 32), (c1.P, 32)):
        hostsim.use_coarse(40)
        hs = hostsim.HostSim(chain_prefix(c1, n) if n < c1.P else c1, jac_numeric=False)
        nP = int((hs.structure()["kind"] == 0).sum())
        assert nP == n - 1
        h, nn = hs.coarse()[:2]
        assert h == want and (h == 0 or nn == (nP + h - 1) // h + 1 <= 40), (n, h, nn)
    hostsim.use_coarse(24)
    assert hostsim.HostSim(chain_prefix(c1, 313), jac_numeric=False).coarse()[0] == 16
    hostsim.use_coarse(40)
    assert hostsim.HostSim(gg.make("c2"), jac_numeric=False).coarse()[0] == 0   # 5488 poses
    hostsim.use_coarse(0)
    assert hostsim.HostSim(chain_prefix(c1, 200), jac_numeric=False).coarse()[0] == 0


@pytest.mark.parametrize("seed", range(16))
def test_two_level_solve_on_collision_fuzz(seed):
    """The coarse space is defined over the ROW ORDER of the reduced system, whatever the graph: on random graphs with
    shuffled ids (the rows are then no chain at all), duplicate and reversed edges, fixed and inactive vertices and DCS
    edges the two-level operator is still symmetric positive definite and the solve still returns the oracle's step."""
    from test_structure_hostsim import _random_graph
    g = _random_graph(np.random.default_rng(5000 + seed), p_range=(40, 130))
    o = Oracle(g)
    if not o.initialize_optimization():
        return
    hostsim.use_coarse(40)
    hs = hostsim.HostSim(g, jac_numeric=False, tol=1e-12)
    assert hs.status == capi.OK, hs.error
    lam = 10.0
    ok, xo = o.solve_once(lam, JAC_ANALYTIC)
    f, xh, it, rel = hs.solve_once(lam)
    h, nn, failed, Ainv = hs.coarse()
    n_free_poses = int((hs.structure()["kind"] == 0).sum())
    assert (h > 0) == (n_free_poses >= 24) and not failed
    assert ok and f == 0, (ok, f)
    assert np.abs(xh - xo).max() <= 1e-8 * max(1e-3, np.abs(xo).max())
    if h:
        assert np.linalg.eigvalsh(Ainv).min() > 0


def test_two_level_gauss_newton_on_a_pose_graph():
    """lambda = 0 and no landmarks at all (the reference's pose graph: GN + DCS closures, submap_loop_closer.cpp:286-288):
    the coarse matrix is R Hpp R^T alone, no G lists, and the loop closures put blocks far off the chain's diagonal."""
    from oracle.cpu_oracle import ALGO_GN
    g = gg.make("c1").pose_only(phi=1.0)
    keep = (g.pp_i < 300) & (g.pp_j < 300)
    import dataclasses
    g = dataclasses.replace(g, pose_id=g.pose_id[:300], pose_est=g.pose_est[:300].copy(), pose_fixed=g.pose_fixed[:300],
                            pose_gt=g.pose_gt[:300], pp_i=g.pp_i[keep], pp_j=g.pp_j[keep], pp_z=g.pp_z[keep],
                            pp_info=g.pp_info[keep], pp_phi=g.pp_phi[keep], pp_seq=g.pp_seq[keep])
    o = Oracle(g)
    assert o.initialize_optimization()
    assert o.optimize(6, ALGO_GN)[0] == 6
    its = {}
    for nodes in (0, 40):
        hostsim.use_coarse(nodes)
        hs = hostsim.HostSim(g, tol=1e-12)
        n, stats = hs.optimize(6, capi.ALGO_GN)
        assert n == 6
        np.testing.assert_allclose(hs.estimates()[0], o.estimates()[0], atol=1e-8)
        its[nodes] = sum(s["pcg_iters"] for s in stats)
        assert (hs.coarse()[0] > 0) == (nodes > 0)
    assert its[40] < its[0], its
