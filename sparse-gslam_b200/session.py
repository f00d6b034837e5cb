"""Caller protocol of the landmark graph: what `Drone::msgCallback` does around the optimiser once
per key-frame (reference src/sparse_gslam/src/drone.cpp:111-190), driven by a synthetic key-frame stream.

Per key-frame the reference
  1. adds the new pose vertex + odometry edge and the pose-line edges of the extracted segments (new landmarks get
     their first observation mapped to the world frame as estimate, drone.cpp:246),
  2. `initializeOptimization()` (after a rejection / the first time) or `updateInitialization(new vertices, new edges)`,
     `push()`, `optimize(15, online)`                                                        (drone.cpp:146-156),
  3. `computeActiveErrors()`, `chi2 = activeChi2()`, dof = sum of the active edges' dimensions, and gates on the 0.99
     quantile of the chi-square distribution with dof degrees of freedom                      (drone.cpp:161-167),
  4. rejects the key-frame's data association: removes its pose-line edges (and landmarks left without edges),
     `pop()`, re-initialise next time (drone.cpp:168-181) -- or accepts: `discardTop()`         (drone.cpp:183-184).

This module keeps the host-side graph (the deques of `LandmarkGraph`, graphs.h:15-27, as growing SoA arrays) and runs
that protocol against a backend object: `GpuBackend` (this repo's optimiser through the C ABI; `updateInitialization`
is a re-initialisation) or, in the tests and the bench's CPU arm, the oracle. The landmark end-point
bookkeeping (`updateEndpoints`, vertex_rhotheta.cpp:9-26) is visualisation state and stays with the caller.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

from . import capi
from . import graphgen as gg

GATE_QUANTILE = 0.99  # drone.cpp:167 (the paper's eq. 19 says 0.95; the code wins)
LM_ITERS = 15         # drone.cpp:150,155


def chi2_quantile(q: float, dof: int) -> float:
    """boost::math::quantile(chi_squared(dof), q)"""
    from scipy.stats import chi2
    return float(chi2.ppf(q, dof))


@dataclass
class KeyFrame:
    """One key-frame of the stream: the new pose, its odometry edge to the previous pose and its line observations.
    `obs_lm` are stream-level landmark keys; a key seen for the first time creates the landmark with `obs_init`."""
    pose_init: np.ndarray                 # [3] dead-reckoned initial estimate
    odom_z: np.ndarray | None             # [3] relative pose from the previous key-frame (None for the first)
    odom_info: np.ndarray | None          # [6]
    obs_lm: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    obs_z: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))
    obs_info: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    obs_init: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))  # world-frame (rho, theta) if new


def stream_from_graph(g: gg.Graph, corrupt_at: dict[int, int] | None = None) -> list[KeyFrame]:
    """Replays a generated pose + line-landmark graph in temporal order (pose ids are temporal, drone.cpp:121).
    Loop-closure pose-pose edges are not part of the landmark graph and are skipped. `corrupt_at[k] = n` attaches the
    first n observations of key-frame k to wrong landmarks (a bad data association the gate must reject)."""
    corrupt_at = corrupt_at or {}
    odo = {int(j): k for k, (i, j) in enumerate(zip(g.pp_i, g.pp_j)) if j - i == 1}
    by_pose = [[] for _ in range(g.P)]
    for k in np.argsort(g.pl_seq, kind="stable"):
        by_pose[int(g.pl_pose[k])].append(int(k))
    frames = []
    for p in range(g.P):
        ks = by_pose[p]
        lm = g.pl_lm[ks].astype(np.int64)
        if p in corrupt_at and len(ks):
            n = min(corrupt_at[p], len(ks))
            lm = lm.copy()
            lm[:n] = (lm[:n] + g.L // 2 + 1) % g.L   # some other, far-away wall
        k = odo.get(p)
        frames.append(KeyFrame(pose_init=g.pose_est[p].copy(), odom_z=None if k is None else g.pp_z[k].copy(),
                               odom_info=None if k is None else g.pp_info[k].copy(), obs_lm=lm, obs_z=g.pl_z[ks].copy(),
                               obs_info=g.pl_info[ks].copy(), obs_init=g.lm_est[lm].copy()))
    return frames


class GpuBackend:
    """The product path: one sgb handle, estimates resident on the device between the calls of one key-frame.
    `prof` accumulates the wall time of every protocol call and the device-side counters of optimize()."""

    def __init__(self, jacobian_mode=capi.JAC_G2O_NUMERIC, device=-1):
        from .optimizer import SparseOptimizerB200
        self.opt = SparseOptimizerB200(capi.ALGO_LM, jacobian_mode=jacobian_mode, device=device)
        self.prof = dict(initialize_s=0.0, push_s=0.0, optimize_s=0.0, chi2_s=0.0, pop_discard_s=0.0, estimates_s=0.0,
                         device_ms=0.0, pcg_iters=0, trials=0, kernel_launches=0, calls=0)

    def _timed(self, key, f, *a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        self.prof[key] += time.perf_counter() - t0
        return r

    def initialize(self, g, online=False, n_new_poses=0, n_new_landmarks=0, n_new_pp=0, n_new_pl=0) -> bool:
        return self._timed("initialize_s", self.opt.initialize_optimization, g)

    def push(self):
        self._timed("push_s", self.opt.push)

    def pop(self):
        self._timed("pop_discard_s", self.opt.pop)

    def discard_top(self):
        self._timed("pop_discard_s", self.opt.discard_top)

    def optimize(self, iters, online):
        n = self._timed("optimize_s", self.opt.optimize, iters, online=online)[0]
        t = self.opt.timings()
        self.prof["device_ms"] += t["total_ms"]
        for k in ("pcg_iters", "trials", "kernel_launches"):
            self.prof[k] += t[k]
        self.prof["calls"] += 1
        return n

    def active_chi2(self):
        return self._timed("chi2_s", self.opt.active_chi2)[0]

    def estimates(self):
        return self._timed("estimates_s", self.opt.estimates)


@dataclass
class FrameLog:
    frame: int
    accepted: bool
    iterations: int
    chi2: float
    dof: int
    gate: float
    n_poses: int
    n_landmarks: int
    n_edges: int
    seconds: float


class LandmarkGraphSession:
    """Host-side `LandmarkGraph` + the per-key-frame protocol of `Drone::msgCallback`."""

    def __init__(self, backend, quantile: float = GATE_QUANTILE, iters: int = LM_ITERS):
        self.backend = backend
        self.quantile, self.iters = quantile, iters
        self.pose_est = np.zeros((0, 3))
        self.lm_est = np.zeros((0, 2))
        self.lm_key: list[int] = []              # landmark array index -> stream key
        self.key_lm: dict[int, int] = {}
        self.pp = dict(i=[], j=[], z=[], info=[], seq=[])
        self.pl = dict(p=[], l=[], z=[], info=[], seq=[])
        self.seq = 0
        self.need_reinit = True
        self.log: list[FrameLog] = []

    # -- graph assembly (drone.cpp:115-143, 193-246)
    def graph(self) -> gg.Graph:
        P, L = len(self.pose_est), len(self.lm_est)
        fixed = np.zeros(P, np.uint8)
        if P:
            fixed[0] = 1                          # first pose fixed (drone.cpp:66)
        f64 = np.float64
        return gg.Graph(
            name="landmark-graph", pose_id=np.arange(P, dtype=np.int32), pose_est=self.pose_est.copy(), pose_fixed=fixed,
            pose_gt=np.full((P, 3), np.nan), lm_id=(10_000_000 + np.arange(L)).astype(np.int32), lm_est=self.lm_est.copy(),
            lm_fixed=np.zeros(L, np.uint8), lm_gt=np.full((L, 2), np.nan),
            pp_i=np.asarray(self.pp["i"], np.int32), pp_j=np.asarray(self.pp["j"], np.int32),
            pp_z=np.asarray(self.pp["z"], f64).reshape(-1, 3), pp_info=np.asarray(self.pp["info"], f64).reshape(-1, 6),
            pp_phi=np.zeros(len(self.pp["i"])), pp_seq=np.asarray(self.pp["seq"], np.int64),
            pl_pose=np.asarray(self.pl["p"], np.int32), pl_lm=np.asarray(self.pl["l"], np.int32),
            pl_z=np.asarray(self.pl["z"], f64).reshape(-1, 2), pl_info=np.asarray(self.pl["info"], f64).reshape(-1, 3),
            pl_seq=np.asarray(self.pl["seq"], np.int64))

    def add_keyframe(self, kf: KeyFrame) -> FrameLog:
        t0 = time.perf_counter()
        p = len(self.pose_est)
        self.pose_est = np.vstack([self.pose_est, kf.pose_init[None, :]])
        if kf.odom_z is not None and p > 0:
            for key, v in zip(("i", "j", "z", "info", "seq"), (p - 1, p, kf.odom_z, kf.odom_info, self.seq)):
                self.pp[key].append(v)
            self.seq += 1
        first_new_edge, first_new_lm = len(self.pl["p"]), len(self.lm_est)
        for key_, z, info, init in zip(kf.obs_lm, kf.obs_z, kf.obs_info, kf.obs_init):
            key_ = int(key_)
            if key_ not in self.key_lm:               # mergeLine found no match: new landmark (drone.cpp:236-247)
                self.key_lm[key_] = len(self.lm_est)
                self.lm_key.append(key_)
                self.lm_est = np.vstack([self.lm_est, init[None, :]])
            for k2, v in zip(("p", "l", "z", "info", "seq"), (p, self.key_lm[key_], z, info, self.seq)):
                self.pl[k2].append(v)
            self.seq += 1
        n_edges = len(self.pp["i"]) + len(self.pl["p"])
        if n_edges == 0:                              # nothing to optimise yet (first key-frame without segments)
            rec = FrameLog(p, True, 0, 0.0, 0, 0.0, p + 1, len(self.lm_est), 0, time.perf_counter() - t0)
            self.log.append(rec)
            return rec
        g = self.graph()
        # initializeOptimization() / updateInitialization(new_vset, new_eset); push(); optimize(15, online)
        online = not self.need_reinit
        ok = self.backend.initialize(g)
        iters = 0
        if ok:
            self.backend.push()
            iters = self.backend.optimize(self.iters, online)
        self.need_reinit = False
        dof = 3 * len(self.pp["i"]) + 2 * len(self.pl["p"])   # sum of the active edges' dimensions
        chi2 = self.backend.active_chi2() if ok else 0.0
        gate = chi2_quantile(self.quantile, dof)
        accepted = not (chi2 > gate)
        if not accepted:
            # reject the data association of this key-frame: drop its pose-line edges and the landmarks they created,
            # restore the estimates, re-initialise next time (drone.cpp:168-181)
            for k2 in self.pl:
                del self.pl[k2][first_new_edge:]
            for key_ in self.lm_key[first_new_lm:]:
                del self.key_lm[key_]
            del self.lm_key[first_new_lm:]
            if ok:
                self.backend.pop()
                pe, le = self.backend.estimates()
                self.pose_est = pe
                self.lm_est = le[:first_new_lm]
            else:
                self.lm_est = self.lm_est[:first_new_lm]
            self.need_reinit = True
        elif ok:
            self.backend.discard_top()
            self.pose_est, self.lm_est = self.backend.estimates()
        rec = FrameLog(p, accepted, iters, chi2, dof, gate, len(self.pose_est), len(self.lm_est),
                       len(self.pp["i"]) + len(self.pl["p"]), time.perf_counter() - t0)
        self.log.append(rec)
        return rec

    def run(self, frames):
        for kf in frames:
            self.add_keyframe(kf)
        return self.log
