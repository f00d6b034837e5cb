"""Host-side mirror of the reference's PoseGraph (include/graphs.h:27-40) and of the edits made around its optimiser,
on top of the C ABI (sgb_pg_*). Method names follow what the reference does to `drone.pose_graph`:

    append_from_lm   submap_loop_closer.cpp:206-223   copy the newly optimised landmark-graph poses (re-measured odometry)
    add_closure      submap_loop_closer.cpp:272-285   EdgeSE2 + DCS kernel between two vertices
    optimize         submap_loop_closer.cpp:286-288   initializeOptimization(); optimize(20)   (GN, setup_pose_opt)
    prune_closures   log_runner.cpp:182-190           chi2() > 11.345 -> removeEdge

The values stay in device memory between calls; only index arrays are mirrored on the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .optimizer import SgbError, SparseOptimizerB200, _p

CHI2_REJECT = 11.345  # log_runner.cpp:184


class PoseGraphB200:
    def __init__(self, device=-1, solver: SparseOptimizerB200 | None = None):
        self.L = capi.load()
        if self.L.sgb_device_count() <= 0:
            raise SgbError(capi.ERR_NO_DEVICE, "no CUDA device: the backend has no CPU path")
        self.pg = C.c_void_p()
        st = self.L.sgb_pg_create(device, C.byref(self.pg))
        if st != capi.OK:
            raise SgbError(st, "sgb_pg_create failed")
        # setup_pose_opt (graphs.cpp:17-23): Gauss-Newton
        self.solver = solver or SparseOptimizerB200(capi.ALGO_GN, jacobian_mode=capi.JAC_ANALYTIC, device=device)

    def close(self):
        if getattr(self, "pg", None) and self.pg.value:
            self.L.sgb_pg_destroy(self.pg)
            self.pg = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != capi.OK:
            raise SgbError(st, self.L.sgb_pg_last_error(self.pg).decode())

    def reset(self, first_est, first_id=0):
        e = np.ascontiguousarray(first_est, np.float64)
        self._check(self.L.sgb_pg_reset(self.pg, int(first_id), _p(e)))

    def append_from_lm(self, lm: SparseOptimizerB200, lm_first: int, count: int, info, ids=None):
        info = np.ascontiguousarray(info, np.float64)
        assert info.size == 6 * count
        ids = None if ids is None else np.ascontiguousarray(ids, np.int32)
        self._check(self.L.sgb_pg_append_from_lm(self.pg, lm.h, int(lm_first), int(count), _p(ids), _p(info)))

    def append_from_host(self, lm_est, info, ids=None):
        lm_est = np.ascontiguousarray(lm_est, np.float64).reshape(-1, 3)
        count = lm_est.shape[0] - 1
        info = np.ascontiguousarray(info, np.float64)
        assert info.size == 6 * count
        ids = None if ids is None else np.ascontiguousarray(ids, np.int32)
        self._check(self.L.sgb_pg_append_from_host(self.pg, _p(lm_est), count, _p(ids), _p(info)))

    def add_closure(self, i, j, z, info, dcs_phi) -> int:
        z = np.ascontiguousarray(z, np.float64)
        info = np.ascontiguousarray(info, np.float64)
        idx = C.c_int32(-1)
        self._check(self.L.sgb_pg_add_closure(self.pg, int(i), int(j), _p(z), _p(info), float(dcs_phi), C.byref(idx)))
        return idx.value

    def optimize(self, iters=20):
        stats = (capi.IterStat * max(1, iters))()
        done = C.c_int32(-1)
        st = self.L.sgb_pg_optimize(self.pg, self.solver.h, self.solver.algo, iters, C.byref(done), C.cast(stats, C.c_void_p))
        self._check(st)
        self.solver.P, self.solver.Lm = self.info()["n_poses"], 0  # the solver now holds this graph
        n = done.value
        return n, [stats[i].as_dict() for i in range(max(n, 0))]

    def prune_closures(self, threshold=CHI2_REJECT):
        info = self.info()
        chi = np.zeros(info["n_closures"])
        act = np.zeros(info["n_closures"], np.uint8)
        removed = C.c_int32(0)
        self._check(self.L.sgb_pg_prune_closures(self.pg, float(threshold), C.byref(removed), _p(chi), _p(act)))
        return removed.value, chi, act.astype(bool)

    def info(self):
        o = capi.PgInfo()
        self._check(self.L.sgb_pg_get_info(self.pg, C.byref(o)))
        return o.as_dict()

    def download(self):
        n = self.info()
        P, E = n["n_poses"], n["n_edges"]
        out = dict(pose_id=np.zeros(P, np.int32), pose_est=np.zeros((P, 3)), pp_i=np.zeros(E, np.int32),
                   pp_j=np.zeros(E, np.int32), pp_z=np.zeros((E, 3)), pp_info=np.zeros((E, 6)), pp_phi=np.zeros(E),
                   pp_active=np.zeros(E, np.uint8), pp_is_closure=np.zeros(E, np.uint8))
        self._check(self.L.sgb_pg_download(self.pg, *[_p(out[k]) for k in
                                                     ("pose_id", "pose_est", "pp_i", "pp_j", "pp_z", "pp_info", "pp_phi",
                                                      "pp_active", "pp_is_closure")]))
        return out
