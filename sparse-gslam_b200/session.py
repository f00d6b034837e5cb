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
    from scipy.special import chdtri   # the inverse survival function scipy.stats.chi2.ppf wraps, without its argument checks
    return float(chdtri(dof, 1.0 - q))


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

    def __init__(self, jacobian_mode=capi.JAC_G2O_NUMERIC, device=-1, incremental=True, coarse_nodes=0):
        import os
        from .optimizer import SparseOptimizerB200
        if os.environ.get("SGB_SESSION_FULL") == "1":   # tuning runs: every key-frame re-initialises from host buffers
            incremental = False
        self.incremental = incremental
        self.opt = SparseOptimizerB200(capi.ALGO_LM, jacobian_mode=jacobian_mode, device=device, incremental=incremental,
                                       coarse_nodes=coarse_nodes)
        self.prof = dict(initialize_s=0.0, push_s=0.0, optimize_s=0.0, chi2_s=0.0, pop_discard_s=0.0, estimates_s=0.0,
                         device_ms=0.0, pcg_iters=0, trials=0, kernel_launches=0, calls=0)

    def _timed(self, key, f, *a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        self.prof[key] += time.perf_counter() - t0
        return r

    def initialize(self, g, online=False, n_new_poses=0, n_new_landmarks=0, n_new_pp=0, n_new_pl=0) -> bool:
        """initializeOptimization() (online == False: the whole graph from host buffers) or updateInitialization(new
        vertices, new edges) (online: only the delta crosses the boundary, drone.cpp:152-153)."""
        if online and self.incremental:
            self.prof["online_updates"] = self.prof.get("online_updates", 0) + 1
            return self._timed("initialize_s", self.opt.update_initialization, g, n_new_poses, n_new_landmarks, n_new_pp, n_new_pl)
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


class _Grow:
    """Append-only 2-D array with amortised O(1) appends (the deques of `LandmarkGraph`, graphs.h:15-27)."""

    def __init__(self, cols, dtype):
        self.a = np.zeros((64, cols), dtype)
        self.n = 0

    def append(self, row):
        if self.n == len(self.a):
            self.a = np.concatenate([self.a, np.zeros_like(self.a)])
        self.a[self.n] = row
        self.n += 1

    def truncate(self, n):
        self.n = n

    def view(self):
        return self.a[:self.n]


class LandmarkGraphSession:
    """Host-side `LandmarkGraph` + the per-key-frame protocol of `Drone::msgCallback`."""

    def __init__(self, backend, quantile: float = GATE_QUANTILE, iters: int = LM_ITERS):
        self.backend = backend
        self.quantile, self.iters = quantile, iters
        self._pose = _Grow(3, np.float64)
        self._lm = _Grow(2, np.float64)
        self.lm_key: list[int] = []              # landmark array index -> stream key
        self.key_lm: dict[int, int] = {}
        # edges: index columns, values, insertion rank
        self._pp_ij, self._pp_z, self._pp_info, self._pp_seq = _Grow(2, np.int32), _Grow(3, np.float64), _Grow(6, np.float64), _Grow(1, np.int64)
        self._pl_pl, self._pl_z, self._pl_info, self._pl_seq = _Grow(2, np.int32), _Grow(2, np.float64), _Grow(3, np.float64), _Grow(1, np.int64)
        self.seq = 0
        self.need_reinit = True
        self._planned = (0, 0, 0, 0)             # vertices / edges of the graph the backend holds
        import inspect
        self._backend_online = "online" in inspect.signature(backend.initialize).parameters
        self.log: list[FrameLog] = []

    @property
    def pp(self):
        """pose-pose (odometry) edges so far, insertion order"""
        ij = self._pp_ij.view()
        return dict(i=ij[:, 0], j=ij[:, 1], z=self._pp_z.view(), info=self._pp_info.view(), seq=self._pp_seq.view()[:, 0])

    @property
    def pl(self):
        """pose-line edges so far, insertion order"""
        pl = self._pl_pl.view()
        return dict(p=pl[:, 0], l=pl[:, 1], z=self._pl_z.view(), info=self._pl_info.view(), seq=self._pl_seq.view()[:, 0])

    @property
    def pose_est(self):
        return self._pose.view()

    @property
    def lm_est(self):
        return self._lm.view()

    # -- graph assembly (drone.cpp:115-143, 193-246)
    def graph(self) -> gg.Graph:
        P, L = self._pose.n, self._lm.n
        fixed = np.zeros(P, np.uint8)
        if P:
            fixed[0] = 1                          # first pose fixed (drone.cpp:66)
        ij, pl = self._pp_ij.view(), self._pl_pl.view()
        c = np.ascontiguousarray
        return gg.Graph(
            name="landmark-graph", pose_id=np.arange(P, dtype=np.int32), pose_est=self._pose.view().copy(), pose_fixed=fixed,
            pose_gt=None, lm_id=np.arange(10_000_000, 10_000_000 + L, dtype=np.int32), lm_est=self._lm.view().copy(),
            lm_fixed=np.zeros(L, np.uint8), lm_gt=None,
            pp_i=c(ij[:, 0]), pp_j=c(ij[:, 1]), pp_z=self._pp_z.view(), pp_info=self._pp_info.view(),
            pp_phi=np.zeros(len(ij)), pp_seq=self._pp_seq.view()[:, 0],
            pl_pose=c(pl[:, 0]), pl_lm=c(pl[:, 1]), pl_z=self._pl_z.view(), pl_info=self._pl_info.view(),
            pl_seq=self._pl_seq.view()[:, 0])

    def add_keyframe(self, kf: KeyFrame) -> FrameLog:
        t0 = time.perf_counter()
        p = self._pose.n
        self._pose.append(kf.pose_init)
        if kf.odom_z is not None and p > 0:
            self._pp_ij.append((p - 1, p))
            self._pp_z.append(kf.odom_z)
            self._pp_info.append(kf.odom_info)
            self._pp_seq.append(self.seq)
            self.seq += 1
        first_new_edge, first_new_lm = self._pl_pl.n, self._lm.n
        for key_, z, info, init in zip(kf.obs_lm, kf.obs_z, kf.obs_info, kf.obs_init):
            key_ = int(key_)
            if key_ not in self.key_lm:               # mergeLine found no match: new landmark (drone.cpp:236-247)
                self.key_lm[key_] = self._lm.n
                self.lm_key.append(key_)
                self._lm.append(init)
            self._pl_pl.append((p, self.key_lm[key_]))
            self._pl_z.append(z)
            self._pl_info.append(info)
            self._pl_seq.append(self.seq)
            self.seq += 1
        n_pp, n_pl = self._pp_ij.n, self._pl_pl.n
        if n_pp + n_pl == 0:                          # nothing to optimise yet (first key-frame without segments)
            rec = FrameLog(p, True, 0, 0.0, 0, 0.0, p + 1, self._lm.n, 0, time.perf_counter() - t0)
            self.log.append(rec)
            return rec
        g = self.graph()
        # initializeOptimization() / updateInitialization(new_vset, new_eset); push(); optimize(15, online)
        online = not self.need_reinit
        if self._backend_online:
            ok = self.backend.initialize(g, online=online, n_new_poses=self._pose.n - self._planned[0],
                                         n_new_landmarks=self._lm.n - self._planned[1], n_new_pp=n_pp - self._planned[2],
                                         n_new_pl=n_pl - self._planned[3])
        else:   # a backend that only knows initializeOptimization (the oracle arm of the tests / bench)
            ok = self.backend.initialize(g)
        self._planned = (self._pose.n, self._lm.n, n_pp, n_pl)   # what the backend's graph holds now
        iters = 0
        if ok:
            self.backend.push()
            iters = self.backend.optimize(self.iters, online)
        self.need_reinit = False
        dof = 3 * n_pp + 2 * n_pl                     # sum of the active edges' dimensions
        chi2 = self.backend.active_chi2() if ok else 0.0
        gate = chi2_quantile(self.quantile, dof)
        accepted = not (chi2 > gate)
        if not accepted:
            # reject the data association of this key-frame: drop its pose-line edges and the landmarks they created,
            # restore the estimates, re-initialise next time (drone.cpp:168-181)
            for a in (self._pl_pl, self._pl_z, self._pl_info, self._pl_seq):
                a.truncate(first_new_edge)
            for key_ in self.lm_key[first_new_lm:]:
                del self.key_lm[key_]
            del self.lm_key[first_new_lm:]
            if ok:
                self.backend.pop()
                pe, le = self.backend.estimates()
                self._pose.view()[:] = pe
                self._lm.view()[:first_new_lm] = le[:first_new_lm]
            self._lm.truncate(first_new_lm)
            self.need_reinit = True
        elif ok:
            self.backend.discard_top()
            pe, le = self.backend.estimates()
            self._pose.view()[:] = pe
            self._lm.view()[:] = le
        rec = FrameLog(p, accepted, iters, chi2, dof, gate, self._pose.n, self._lm.n, self._pp_ij.n + self._pl_pl.n,
                       time.perf_counter() - t0)
        self.log.append(rec)
        return rec

    def run(self, frames):
        for kf in frames:
            self.add_keyframe(kf)
        return self.log
