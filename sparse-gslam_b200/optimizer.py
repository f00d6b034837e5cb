"""Host-side mirror of the g2o calls sparse-gslam makes on its optimisers, on top of the C ABI.

Method names follow `g2o::SparseOptimizer` as used by the reference (drone.cpp:146-190,
submap_loop_closer.cpp:286-288, log_runner.cpp:203-204): initializeOptimization, optimize, push, pop, discardTop,
computeActiveErrors + activeChi2. The graph is handed over as SoA arrays (graphgen.Graph or anything with the same
attributes); the real C++ drop-in is sparse-gslam_b200/adapter/sgb_g2o_adapter.h.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class SgbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"sgb status {status}: {msg}")
        self.status = status


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pack_graph(g):
    keep = []

    def arr(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return _p(a)

    s = capi.GraphSoA()
    s.n_poses = int(g.pose_est.shape[0])
    s.pose_id = arr(g.pose_id, np.int32)
    s.pose_est = arr(g.pose_est, np.float64)
    s.pose_fixed = arr(g.pose_fixed, np.uint8)
    s.n_landmarks = int(g.lm_est.shape[0])
    s.lm_id = arr(g.lm_id, np.int32)
    s.lm_est = arr(g.lm_est, np.float64)
    s.lm_fixed = arr(g.lm_fixed, np.uint8)
    s.n_pp = int(g.pp_i.shape[0])
    s.pp_i = arr(g.pp_i, np.int32)
    s.pp_j = arr(g.pp_j, np.int32)
    s.pp_z = arr(g.pp_z, np.float64)
    s.pp_info = arr(g.pp_info, np.float64)
    s.pp_phi = arr(g.pp_phi, np.float64)
    s.pp_seq = arr(g.pp_seq, np.int64)
    s.n_pl = int(g.pl_pose.shape[0])
    s.pl_pose = arr(g.pl_pose, np.int32)
    s.pl_lm = arr(g.pl_lm, np.int32)
    s.pl_z = arr(g.pl_z, np.float64)
    s.pl_info = arr(g.pl_info, np.float64)
    s.pl_seq = arr(g.pl_seq, np.int64)
    return s, keep


class SparseOptimizerB200:
    """One optimiser handle = one of the reference's two graphs (LandmarkGraph.opt / PoseGraph.opt, graphs.h:19-40)."""

    def __init__(self, algo=capi.ALGO_LM, jacobian_mode=capi.JAC_G2O_NUMERIC, pcg_tolerance=1e-10, pcg_max_iters=0,
                 device=-1, lm_user_lambda=0.0, incremental=False, coarse_nodes=0):
        self.L = capi.load()
        if self.L.sgb_device_count() <= 0:
            raise SgbError(capi.ERR_NO_DEVICE, "no CUDA device: the backend has no CPU path")
        opt = capi.Options()
        self.L.sgb_default_options(C.byref(opt))
        opt.device = device
        opt.jacobian_mode = jacobian_mode
        opt.pcg_tolerance = pcg_tolerance
        opt.pcg_max_iters = pcg_max_iters
        opt.lm_user_lambda = lm_user_lambda
        opt.incremental = int(bool(incremental))
        opt.coarse_nodes = int(coarse_nodes)   # two-level preconditioner of the resident solve: >0 on, <0 off, 0 default
        self.algo = algo
        self.h = C.c_void_p()
        st = self.L.sgb_create(C.byref(opt), C.byref(self.h))
        if st != capi.OK:
            raise SgbError(st, "sgb_create failed")
        self.g = None
        self.P = self.Lm = 0

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.sgb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != capi.OK:
            raise SgbError(st, self.L.sgb_last_error(self.h).decode())

    # g2o: addVertex/addEdge for the whole graph + initializeOptimization()
    def initialize_optimization(self, g) -> bool:
        s, keep = pack_graph(g)
        st = self.L.sgb_set_graph(self.h, C.byref(s))
        if st == capi.ERR_NOT_INITIALIZED:
            return False
        self._check(st)
        self.g = g
        self.P, self.Lm = s.n_poses, s.n_landmarks
        return True

    def update_initialization(self, g, n_new_poses, n_new_landmarks, n_new_pp, n_new_pl) -> bool:
        """g2o: updateInitialization(new vertices, new edges) (drone.cpp:152-153). `g` is the EXTENDED graph whose last
        n_new_* vertices / edges are new; only those cross the boundary (sgb_update_graph), the estimates of the
        existing vertices stay what they are on the device. Needs incremental=True at construction."""
        def tail(a, n, dt):
            a = np.ascontiguousarray(a[len(a) - n:], dtype=dt)
            return a

        P, Lm, E, F = len(g.pose_est), len(g.lm_est), len(g.pp_i), len(g.pl_pose)
        keep = [tail(g.pose_id, n_new_poses, np.int32), tail(g.pose_est, n_new_poses, np.float64), tail(g.pose_fixed, n_new_poses, np.uint8),
                tail(g.lm_id, n_new_landmarks, np.int32), tail(g.lm_est, n_new_landmarks, np.float64), tail(g.lm_fixed, n_new_landmarks, np.uint8),
                tail(g.pp_i, n_new_pp, np.int32), tail(g.pp_j, n_new_pp, np.int32), tail(g.pp_z, n_new_pp, np.float64),
                tail(g.pp_info, n_new_pp, np.float64), tail(g.pp_phi, n_new_pp, np.float64), tail(g.pp_seq, n_new_pp, np.int64),
                tail(g.pl_pose, n_new_pl, np.int32), tail(g.pl_lm, n_new_pl, np.int32), tail(g.pl_z, n_new_pl, np.float64),
                tail(g.pl_info, n_new_pl, np.float64), tail(g.pl_seq, n_new_pl, np.int64)]
        d = capi.GraphDelta()
        d.n_new_poses, d.n_new_landmarks, d.n_new_pp, d.n_new_pl = n_new_poses, n_new_landmarks, n_new_pp, n_new_pl
        (d.pose_id, d.pose_est, d.pose_fixed, d.lm_id, d.lm_est, d.lm_fixed, d.pp_i, d.pp_j, d.pp_z, d.pp_info, d.pp_phi,
         d.pp_seq, d.pl_pose, d.pl_lm, d.pl_z, d.pl_info, d.pl_seq) = [_p(a) for a in keep]
        st = self.L.sgb_update_graph(self.h, C.byref(d))
        if st == capi.ERR_NOT_INITIALIZED:
            return False
        self._check(st)
        self.g = g
        self.P, self.Lm = P, Lm
        return True

    def initialize_optimization_device(self, g, dev, pp_slot=None, pl_slot=None, has_robust=None) -> bool:
        """sgb_set_graph_device: the index arrays of `g` on the host, the values already in device memory. `dev` maps
        pose_est / lm_est / pp_z / pp_info / pp_phi / pl_z / pl_info to device pointers (ints; e.g. torch tensors'
        data_ptr()); pp_slot / pl_slot (host int arrays) give the device slot of every edge of `g`."""
        s, keep = pack_graph(g)
        dv = capi.DeviceValues()
        for k in ("pose_est", "lm_est", "pp_z", "pp_info", "pp_phi", "pl_z", "pl_info"):
            setattr(dv, k, dev.get(k) or None)
        if pp_slot is not None:
            pp_slot = np.ascontiguousarray(pp_slot, np.int32)
            dv.pp_slot = _p(pp_slot)
        if pl_slot is not None:
            pl_slot = np.ascontiguousarray(pl_slot, np.int32)
            dv.pl_slot = _p(pl_slot)
        dv.has_robust = int(bool(np.any(g.pp_phi > 0)) if has_robust is None else has_robust)
        st = self.L.sgb_set_graph_device(self.h, C.byref(s), C.byref(dv))
        if st == capi.ERR_NOT_INITIALIZED:
            return False
        self._check(st)
        self.g = g
        self.P, self.Lm = s.n_poses, s.n_landmarks
        return True

    def initialize_partitioned(self, g, world, rank, exchange):
        """Row-block partition across `world` GPUs (one process per GPU). `exchange(blob: bytes) -> list[bytes]` must
        return every rank's 64-byte blob in rank order (e.g. torch.distributed.all_gather_object)."""
        s, keep = pack_graph(g)
        st = self.L.sgb_set_graph_partitioned(self.h, C.byref(s), world, rank)
        if st == capi.ERR_NOT_INITIALIZED:
            return False
        self._check(st)
        self.g = g
        self.P, self.Lm = s.n_poses, s.n_landmarks
        if world > 1:
            blob = C.create_string_buffer(64)
            self._check(self.L.sgb_comm_get_handle(self.h, blob))
            blobs = exchange(bytes(blob.raw))
            assert len(blobs) == world and all(len(b) == 64 for b in blobs)
            allb = C.create_string_buffer(b"".join(blobs), 64 * world)
            self._check(self.L.sgb_comm_connect(self.h, allb, world))
        return True

    def partition_info(self):
        info = capi.PartitionInfo()
        self._check(self.L.sgb_get_partition_info(self.h, C.byref(info)))
        return info.as_dict()

    def structure(self):
        info = capi.StructureInfo()
        self._check(self.L.sgb_get_structure_info(self.h, C.byref(info)))
        nf, nb = info.n_free, info.n_blocks
        kind, index, off = (np.zeros(nf, np.int32) for _ in range(3))
        row, col, nr, nc = (np.zeros(nb, np.int32) for _ in range(4))
        ph, lh = np.zeros(self.P, np.int32), np.zeros(self.Lm, np.int32)
        self._check(self.L.sgb_get_structure(self.h, _p(kind), _p(index), _p(off), _p(row), _p(col), _p(nr), _p(nc),
                                             _p(ph), _p(lh)))
        return dict(n_free=nf, n_blocks=nb, dim=info.scalar_dim, kind=kind, index=index, offset=off, row=row, col=col,
                    nrows=nr, ncols=nc, pose_hidx=ph, lm_hidx=lh, block_values=info.block_values,
                    n_free_poses=info.n_free_poses, n_free_landmarks=info.n_free_landmarks,
                    n_active_pp=info.n_active_pp, n_active_pl=info.n_active_pl)

    def structure_info(self):
        info = capi.StructureInfo()
        self._check(self.L.sgb_get_structure_info(self.h, C.byref(info)))
        return dict(n_free=info.n_free, n_blocks=info.n_blocks, dim=info.scalar_dim, block_values=info.block_values,
                    n_free_poses=info.n_free_poses, n_free_landmarks=info.n_free_landmarks, coarse_nodes=info.coarse_nodes)

    def linearize(self, hessian=True):
        """hessian=False skips the block export (a rank-filtered multi-GPU handle only holds its own share of the
        structure and cannot lay its blocks out in the reference's global order)."""
        st = self.structure_info()
        b = np.zeros(st["dim"])
        H = np.zeros(st["block_values"]) if hessian else None
        chi = np.zeros(2)
        self._check(self.L.sgb_linearize(self.h, _p(b), _p(H), _p(chi)))
        return dict(b=b, H=H, chi2=chi)

    def solve_once(self, lam):
        st = self.structure_info()
        x = np.zeros(st["dim"])
        it = C.c_int32()
        rel = C.c_double()
        rc = self.L.sgb_solve_once(self.h, float(lam), _p(x), C.byref(it), C.byref(rel))
        if rc not in (capi.OK, capi.ERR_SOLVE_FAILED):
            self._check(rc)
        return rc == capi.OK, x, it.value, rel.value

    def compute_marginals(self, block_pairs):
        """g2o::SparseOptimizer::computeMarginals: blocks (row, col) -- Hessian indices -- of the inverse Hessian at the
        current estimates. Returns (ok, list of numpy blocks); ok == False where g2o returns false (H not SPD)."""
        st = self.structure()
        nfp = st["n_free_poses"]
        rows = np.ascontiguousarray([p[0] for p in block_pairs], np.int32)
        cols = np.ascontiguousarray([p[1] for p in block_pairs], np.int32)
        dims = [(3 if r < nfp else 2, 3 if c < nfp else 2) for r, c in zip(rows, cols)]
        out = np.zeros(max(1, sum(a * b for a, b in dims)))
        rc = self.L.sgb_compute_marginals(self.h, len(rows), _p(rows), _p(cols), _p(out))
        if rc == capi.ERR_SOLVE_FAILED:
            return False, []
        self._check(rc)
        blocks, o = [], 0
        for a, b in dims:
            blocks.append(out[o:o + a * b].reshape(b, a).T.copy())  # column-major
            o += a * b
        return True, blocks

    def optimize(self, iters, online=False, resident=False):
        """Returns (g2o return value, per-iteration stats)."""
        stats = (capi.IterStat * max(1, iters))()
        done = C.c_int32(-1)
        if resident:
            st = self.L.sgb_optimize_resident(self.h, self.algo, iters, C.byref(done), C.cast(stats, C.c_void_p))
        else:
            st = self.L.sgb_optimize(self.h, self.algo, iters, int(online), C.byref(done), C.cast(stats, C.c_void_p))
        if st == capi.ERR_NOT_INITIALIZED:
            return -1, []
        self._check(st)
        n = done.value
        return n, [stats[i].as_dict() for i in range(max(n, 0))]

    def estimates(self):
        p = np.zeros((self.P, 3))
        l = np.zeros((self.Lm, 2))
        self._check(self.L.sgb_get_estimates(self.h, _p(p), _p(l)))
        return p, l

    def set_estimates(self, poses, lms):
        p = np.ascontiguousarray(poses, np.float64)
        l = np.ascontiguousarray(lms, np.float64)
        self._check(self.L.sgb_set_estimates(self.h, _p(p), _p(l)))

    def push(self):
        self._check(self.L.sgb_push(self.h))

    def pop(self):
        self._check(self.L.sgb_pop(self.h))

    def discard_top(self):
        self._check(self.L.sgb_discard_top(self.h))

    def active_chi2(self):
        """computeActiveErrors(); returns (activeChi2, activeRobustChi2)."""
        c = np.zeros(2)
        self._check(self.L.sgb_chi2(self.h, _p(c)))
        return float(c[0]), float(c[1])

    def timings(self):
        t = capi.Timings()
        self._check(self.L.sgb_get_timings(self.h, C.byref(t)))
        return t.as_dict()


def optimize_batch(optimizers, iters, resident=False):
    """One launch for many independent optimisers (sgb_optimize_batch): one thread block per graph, the whole LM / GN
    loop on the device. All optimisers must share the algorithm and the device. Returns (g2o return values, last
    iteration's stats) per optimiser."""
    n = len(optimizers)
    if n == 0:
        return [], []
    L = optimizers[0].L
    algo = optimizers[0].algo
    assert all(o.algo == algo for o in optimizers), "one algorithm per batch"
    hs = (C.c_void_p * n)(*[o.h for o in optimizers])
    done = (C.c_int32 * n)()
    stats = (capi.IterStat * n)()
    f = L.sgb_optimize_batch_resident if resident else L.sgb_optimize_batch
    st = f(C.cast(hs, C.c_void_p), n, algo, iters, C.cast(done, C.c_void_p), C.cast(stats, C.c_void_p))
    if st != capi.OK:
        raise SgbError(st, L.sgb_last_error(optimizers[0].h).decode())
    return [done[i] for i in range(n)], [stats[i].as_dict() for i in range(n)]


class LinearSolverB200:
    """Mirror of g2o::LinearSolver<MatrixType> as the reference instantiates it (LinearSolverEigen, graphs.cpp:11,19):
    init() forgets the pattern, solve(A, b) analyses the pattern on the first call after init() and solves. A is the
    upper triangle in g2o's SparseBlockMatrix order: (block_dim [n], col_ptr [n+1], row_idx [nnzb], values)."""

    def __init__(self, device=-1, pcg_tolerance=1e-10, pcg_max_iters=0):
        self.opt = SparseOptimizerB200(capi.ALGO_GN, device=device, pcg_tolerance=pcg_tolerance, pcg_max_iters=pcg_max_iters)
        self._have_pattern = False
        self.last = {}

    def init(self) -> bool:
        self._have_pattern = False
        return True

    def solve(self, block_dim, col_ptr, row_idx, values, b):
        """Returns (ok, x); ok == False is LinearSolver::solve() == false (matrix not positive definite)."""
        L, h = self.opt.L, self.opt.h
        if not self._have_pattern:
            bd = np.ascontiguousarray(block_dim, np.int32)
            cp = np.ascontiguousarray(col_ptr, np.int32)
            ri = np.ascontiguousarray(row_idx, np.int32)
            A = capi.BlockMatrix(len(bd), _p(bd), _p(cp), _p(ri))
            self.opt._check(L.sgb_linear_set_pattern(h, C.byref(A)))
            self._have_pattern = True
        values = np.ascontiguousarray(values, np.float64)
        b = np.ascontiguousarray(b, np.float64)
        x = np.zeros_like(b)
        it, rel = C.c_int32(0), C.c_double(0.0)
        st = L.sgb_linear_solve(h, _p(values), _p(b), _p(x), C.byref(it), C.byref(rel))
        self.last = dict(pcg_iters=it.value, rel_residual=rel.value)
        if st == capi.ERR_SOLVE_FAILED:
            return False, x
        self.opt._check(st)
        return True, x
