"""Caller protocol: the per-key-frame sequence of drone.cpp:146-190 -- initialise / push / optimize(15)
/ chi2 gate at the 0.99 quantile / pop or discardTop -- against the oracle (CPU tier) and the GPU backend (gpu tier)."""
import numpy as np
import pytest

from oracle.cpu_oracle import ALGO_LM, JAC_ANALYTIC, JAC_G2O_NUMERIC, Oracle
from sparse_gslam_b200 import capi
from sparse_gslam_b200 import graphgen as gg
from sparse_gslam_b200.session import GpuBackend, LandmarkGraphSession, chi2_quantile, stream_from_graph


class OracleBackend:
    """The same protocol on the CPU oracle (test infrastructure): push/pop on host copies, as g2o does."""

    def __init__(self, jac=JAC_G2O_NUMERIC):
        self.jac, self.o, self.stack = jac, None, []

    def initialize(self, g):
        self.o = Oracle(g)
        return self.o.initialize_optimization()

    def push(self):
        self.stack.append(self.o.estimates())

    def pop(self):
        self.o.set_estimates(*self.stack.pop())

    def discard_top(self):
        self.stack.pop()

    def optimize(self, iters, online):
        return self.o.optimize(iters, ALGO_LM, self.jac)[0]

    def active_chi2(self):
        return self.o.chi2()[0]

    def estimates(self):
        return self.o.estimates()


def _stream(corrupt=True):
    g = gg.make_small(seed=11, P=40, L=10, E_l=110, n_closures=0)
    return g, stream_from_graph(g, corrupt_at={25: 3} if corrupt else None)


def test_chi2_quantile_known_values():
    assert abs(chi2_quantile(0.99, 1) - 6.6348966) < 1e-6        # tables: chi2_{0.99}(1), (10)
    assert abs(chi2_quantile(0.99, 10) - 23.2092512) < 1e-6


def test_protocol_on_the_oracle_rejects_the_bad_association():
    g, frames = _stream()
    s = LandmarkGraphSession(OracleBackend(JAC_ANALYTIC))
    log = s.run(frames)
    assert [r.frame for r in log if not r.accepted] == [25]
    bad = log[25]
    assert bad.chi2 > bad.gate and log[24].chi2 <= log[24].gate
    # the rejected key-frame keeps its pose and odometry edge but none of its line observations (drone.cpp:159,169-178)
    assert len(s.pose_est) == g.P
    assert int(np.sum(g.pl_pose != 25)) == len(s.pl["p"])
    assert log[26].accepted and log[-1].accepted
    # the optimised trajectory stays close to ground truth (no loop closures in the landmark graph: some drift)
    assert np.abs(s.pose_est[:, :2] - g.pose_gt[:, :2]).max() < 0.6


@pytest.mark.gpu
@pytest.mark.parametrize("jac_gpu,jac_cpu,tol", [(capi.JAC_ANALYTIC, JAC_ANALYTIC, 1e-6),
                                                 (capi.JAC_G2O_NUMERIC, JAC_G2O_NUMERIC, 5e-6)])
def test_gpu_session_matches_oracle_session(jac_gpu, jac_cpu, tol):
    g, frames = _stream()
    ref = LandmarkGraphSession(OracleBackend(jac_cpu))
    ref.run(frames)
    gpu = LandmarkGraphSession(GpuBackend(jacobian_mode=jac_gpu))
    gpu.run(frames)
    assert [r.accepted for r in gpu.log] == [r.accepted for r in ref.log]
    assert [r.dof for r in gpu.log] == [r.dof for r in ref.log]
    for a, b in zip(gpu.log, ref.log):
        np.testing.assert_allclose(a.chi2, b.chi2, rtol=1e-5, atol=1e-9)
    scale = max(1.0, float(np.abs(ref.pose_est[:, :2]).max()))
    d = gpu.pose_est - ref.pose_est
    d[:, 2] = gg.wrap(d[:, 2])
    assert np.abs(d).max() / scale < tol
    assert np.abs(gpu.lm_est - ref.lm_est).max() / max(1.0, float(np.abs(ref.lm_est).max())) < tol


@pytest.mark.gpu
def test_online_update_equals_full_reinitialisation():
    """sgb_update_graph (updateInitialization, drone.cpp:152-153): extending the device-resident graph by the new
    key-frames' vertices and edges gives exactly what a full initializeOptimization of the extended graph gives when it
    starts from the same estimates -- same structure, same values, the same converged state -- and the session takes
    that path for every accepted key-frame."""
    from sparse_gslam_b200 import SparseOptimizerB200
    g, frames = _stream(corrupt=False)
    s = LandmarkGraphSession(OracleBackend(JAC_ANALYTIC))   # only used to assemble the graphs of key-frames 0..k
    s.run(frames[:20])
    ga = s.graph()
    counts_a = (len(ga.pose_est), len(ga.lm_est), len(ga.pp_i), len(ga.pl_pose))
    a = SparseOptimizerB200(capi.ALGO_LM, jacobian_mode=capi.JAC_ANALYTIC, incremental=True)
    assert a.initialize_optimization(ga)
    a.optimize(15)
    pa, la = a.estimates()
    # five more key-frames appended on the host side without optimising in between
    s2 = LandmarkGraphSession(OracleBackend(JAC_ANALYTIC))
    s2.run(frames[:20])
    class _NoOpt(OracleBackend):
        def initialize(self, g):
            self.g = g
            return False
    s2.backend = _NoOpt()
    s2._backend_online = False
    for kf in frames[20:25]:
        s2.add_keyframe(kf)
    gb = s2.graph()
    nb = (len(gb.pose_est), len(gb.lm_est), len(gb.pp_i), len(gb.pl_pose))
    new = tuple(y - x for x, y in zip(counts_a, nb))
    assert all(n > 0 for n in (new[0], new[2], new[3]))
    # the extended graph starts from the optimised estimates of the old vertices
    gb.pose_est[:counts_a[0]] = pa
    gb.lm_est[:counts_a[1]] = la
    assert a.update_initialization(gb, *new)
    a.push()
    n_on, st_on = a.optimize(15, online=True)
    p_on, l_on = a.estimates()
    full = SparseOptimizerB200(capi.ALGO_LM, jacobian_mode=capi.JAC_ANALYTIC)
    assert full.initialize_optimization(gb)
    n_full, st_full = full.optimize(15)
    p_full, l_full = full.estimates()
    # same structure and the same values up to the last bit of the cached inverse measurements (formed on the device on
    # the online path, on the host otherwise): the same converged state; LM may stop an iteration earlier or later once
    # chi2 has stopped moving
    assert n_on >= 1 and n_full >= 1
    np.testing.assert_allclose(p_on, p_full, rtol=0, atol=1e-7)
    np.testing.assert_allclose(l_on, l_full, rtol=0, atol=1e-7)
    np.testing.assert_allclose(a.active_chi2()[0], full.active_chi2()[0], rtol=1e-8)
    assert np.array_equal(a.structure()["row"], full.structure()["row"])
    a.pop()                                   # the caller protocol keeps working on the extended graph
    p_back, _ = a.estimates()
    np.testing.assert_array_equal(p_back[:counts_a[0]], pa)
    # a handle without the option refuses, with a message
    from sparse_gslam_b200 import SgbError
    with pytest.raises(SgbError):
        full.update_initialization(gb, 0, 0, 0, 0)
    # the session uses the online path for every accepted key-frame after the first
    be = GpuBackend(jacobian_mode=capi.JAC_ANALYTIC)
    LandmarkGraphSession(be).run(frames[:12])
    assert be.prof.get("online_updates", 0) >= 9
