"""Two-level preconditioner of the resident solve on the GPU (sgb_coarse.h, k_setup_coarse, k_pcg_res4<.., true>): the
contract of the solve is unchanged -- same damped step, same LM trajectory as the oracle, same key-frame decisions -- in
fewer PCG iterations. The CPU tier checks the same code on the host harness (tests/test_coarse_hostsim.py)."""
import numpy as np
import pytest

from oracle.cpu_oracle import ALGO_LM, JAC_ANALYTIC, Oracle
from sparse_gslam_b200 import capi
from sparse_gslam_b200 import graphgen as gg

pytestmark = pytest.mark.gpu


def _opt(g, nodes, **kw):
    from sparse_gslam_b200 import SparseOptimizerB200
    o = SparseOptimizerB200(capi.ALGO_LM, jacobian_mode=capi.JAC_ANALYTIC, coarse_nodes=nodes, **kw)
    assert o.initialize_optimization(g)
    return o


# 60 rows: one small CTA; 200: one CTA of 1024 threads; 400 / 700: clusters, spacing 16 / 32; c1: 1226 rows, spacing 32
@pytest.mark.parametrize("n", [60, 200, 400, 700, 0])
@pytest.mark.parametrize("lam", [1e-3, 5.0])
def test_damped_step_equals_the_plain_solve(n, lam):
    c1 = gg.make("c1")
    g = c1.chain_prefix(n) if n else c1
    plain, two = _opt(g, -1), _opt(g, 40)
    assert plain.structure_info()["coarse_nodes"] == 0
    nodes = two.structure_info()["coarse_nodes"]
    ok0, x0, it0, rel0 = plain.solve_once(lam)
    ok1, x1, it1, rel1 = two.solve_once(lam)
    assert ok0 and ok1 and rel1 <= 1e-10
    np.testing.assert_allclose(x1, x0, rtol=1e-6, atol=1e-8 * np.abs(x0).max())
    if n and n <= 700:
        assert nodes == (n - 1 + (8 if n <= 313 else 16 if n <= 625 else 32) - 1) // (8 if n <= 313 else 16 if n <= 625 else 32) + 1
    if nodes == 0:   # the coarse inverse did not fit next to the matrices in shared memory: the plain solve, bit for bit
        assert it1 == it0 and np.array_equal(x0, x1)
        return
    assert it1 < it0, (it0, it1, nodes)
    if n >= 200:
        assert it1 <= 0.6 * it0, (it0, it1, nodes)


@pytest.mark.parametrize("n", [120, 300])
def test_lm15_matches_the_oracle(n):
    g = gg.make("c1").chain_prefix(n)
    o = Oracle(g)
    assert o.initialize_optimization()
    n0, s0 = o.optimize(15, ALGO_LM, JAC_ANALYTIC)
    po, lo = o.estimates()
    plain = _opt(g, -1)
    n_plain, s_plain = plain.optimize(15)
    two = _opt(g, 40)
    n1, s1 = two.optimize(15)
    assert n1 == n0 == n_plain
    prev = None
    for a, b in zip(s0, s1):
        # once chi2 has stopped moving the gain ratio is rounding noise and the number of trials of an iteration is
        # anybody's (the oracle and the plain solve differ there too): trials are compared while LM still makes progress
        if prev is None or prev - a["chi2"] > 1e-9 * prev:
            assert a["trials"] == b["trials"]
        np.testing.assert_allclose(b["chi2"], a["chi2"], rtol=1e-8)
        prev = a["chi2"]
    p1, l1 = two.estimates()
    np.testing.assert_allclose(p1, po, atol=1e-6)
    np.testing.assert_allclose(l1, lo, atol=1e-6)
    it_plain, it_two = sum(s["pcg_iters"] for s in s_plain), sum(s["pcg_iters"] for s in s1)
    assert it_two <= 0.5 * it_plain, (it_plain, it_two)


def test_keyframe_stream_same_decisions_fewer_iterations():
    """The reference's per-key-frame protocol (drone.cpp:146-190) with and without the coarse term: the same accepted /
    rejected key-frames (one key-frame carries three wrong data associations; whether the gate rejects it is the
    protocol's business, test_session.py covers that) and the same final estimates; the online path re-plans the coarse
    lists with every key-frame."""
    from sparse_gslam_b200.session import GpuBackend, LandmarkGraphSession, stream_from_graph
    g = gg.make("c1")
    frames = stream_from_graph(g, corrupt_at={90: 3})[:140]
    runs = {}
    for nodes in (-1, 40):
        be = GpuBackend(jacobian_mode=capi.JAC_ANALYTIC, coarse_nodes=nodes)
        s = LandmarkGraphSession(be)
        log = s.run(frames)
        runs[nodes] = (log, np.array(s.pose_est), np.array(s.lm_est), be.prof["pcg_iters"])
    la, pa, ma, ia = runs[-1]
    lb, pb, mb, ib = runs[40]
    assert [r.accepted for r in la] == [r.accepted for r in lb]
    np.testing.assert_allclose(pb, pa, atol=1e-6)
    np.testing.assert_allclose(mb, ma, atol=1e-6)
    assert ib <= 0.6 * ia, (ia, ib)
