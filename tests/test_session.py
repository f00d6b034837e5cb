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
