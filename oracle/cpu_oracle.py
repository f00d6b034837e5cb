"""ctypes binding of the C++ CPU oracle (oracle/sgo_oracle.cpp). TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
PARITY UNPINNED: see the header of oracle/sgo_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ALGO_LM, ALGO_GN = 0, 1
JAC_G2O_NUMERIC, JAC_ANALYTIC = 0, 1


class _Graph(C.Structure):
    _fields_ = [
        ("n_poses", C.c_int32), ("pose_id", C.c_void_p), ("pose_est", C.c_void_p), ("pose_fixed", C.c_void_p),
        ("n_landmarks", C.c_int32), ("lm_id", C.c_void_p), ("lm_est", C.c_void_p), ("lm_fixed", C.c_void_p),
        ("n_pp", C.c_int32), ("pp_i", C.c_void_p), ("pp_j", C.c_void_p), ("pp_z", C.c_void_p), ("pp_info", C.c_void_p),
        ("pp_phi", C.c_void_p), ("pp_seq", C.c_void_p),
        ("n_pl", C.c_int32), ("pl_pose", C.c_void_p), ("pl_lm", C.c_void_p), ("pl_z", C.c_void_p),
        ("pl_info", C.c_void_p), ("pl_seq", C.c_void_p),
    ]


class IterStat(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("trials", C.c_int32), ("result", C.c_int32), ("pad", C.c_int32),
                ("chi2", C.c_double), ("lambda_", C.c_double), ("rho", C.c_double), ("chi2_before", C.c_double)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libsgo_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("sgo_oracle.cpp", "sgo_frontend.cpp", "sgo_oracle.h")]
    def stale():
        return not os.path.exists(so) or any(os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(so) for f in srcs)

    if force or stale():
        import fcntl
        with open(so + ".lock", "w") as lk:  # several test processes may get here at once
            fcntl.flock(lk, fcntl.LOCK_EX)
            if force or stale():
                subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.sgo_create.restype = C.c_void_p
        L.sgo_destroy.argtypes = [C.c_void_p]
        L.sgo_set_graph.argtypes = [C.c_void_p, C.POINTER(_Graph)]
        L.sgo_initialize.argtypes = [C.c_void_p]
        for f in ("sgo_num_free", "sgo_num_blocks", "sgo_scalar_dim"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.sgo_block_values_size.argtypes = [C.c_void_p]
        L.sgo_block_values_size.restype = C.c_int64
        L.sgo_get_order.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.sgo_get_blocks.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.sgo_get_hessian_index.argtypes = [C.c_void_p] + [C.c_void_p] * 2
        L.sgo_linearize.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 9
        L.sgo_optimize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.sgo_get_estimates.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.sgo_set_estimates.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.sgo_chi2.argtypes = [C.c_void_p, C.c_void_p]
        L.sgo_solve_once.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        L.sgo_last_profile.argtypes = [C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pack_graph(g, struct_cls=_Graph):
    """Graph (graphgen.Graph) -> (ctypes struct, keep-alive list). Shared with the product binding's tests."""
    keep = []

    def arr(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return _p(a)

    s = struct_cls()
    s.n_poses = g.P
    s.pose_id = arr(g.pose_id, np.int32)
    s.pose_est = arr(g.pose_est, np.float64)
    s.pose_fixed = arr(g.pose_fixed, np.uint8)
    s.n_landmarks = g.L
    s.lm_id = arr(g.lm_id, np.int32)
    s.lm_est = arr(g.lm_est, np.float64)
    s.lm_fixed = arr(g.lm_fixed, np.uint8)
    s.n_pp = g.n_pp
    s.pp_i = arr(g.pp_i, np.int32)
    s.pp_j = arr(g.pp_j, np.int32)
    s.pp_z = arr(g.pp_z, np.float64)
    s.pp_info = arr(g.pp_info, np.float64)
    s.pp_phi = arr(g.pp_phi, np.float64)
    s.pp_seq = arr(g.pp_seq, np.int64)
    s.n_pl = g.n_pl
    s.pl_pose = arr(g.pl_pose, np.int32)
    s.pl_lm = arr(g.pl_lm, np.int32)
    s.pl_z = arr(g.pl_z, np.float64)
    s.pl_info = arr(g.pl_info, np.float64)
    s.pl_seq = arr(g.pl_seq, np.int64)
    return s, keep


class Oracle:
    """Mirror of the g2o calls the reference makes: initialize_optimization(), optimize(n), chi2(), estimates."""

    def __init__(self, g):
        self.L = lib()
        self.h = self.L.sgo_create()
        self.g = g
        s, keep = pack_graph(g)
        rc = self.L.sgo_set_graph(self.h, C.byref(s))
        if rc != 0:
            raise ValueError("sgo_set_graph failed (bad vertex index)")
        self.initialized = False

    def __del__(self):
        try:
            self.L.sgo_destroy(self.h)
        except Exception:
            pass

    def initialize_optimization(self) -> bool:
        self.initialized = bool(self.L.sgo_initialize(self.h))
        return self.initialized

    def structure(self):
        nf, nb = self.L.sgo_num_free(self.h), self.L.sgo_num_blocks(self.h)
        kind, index, off = (np.zeros(nf, np.int32) for _ in range(3))
        self.L.sgo_get_order(self.h, _p(kind), _p(index), _p(off))
        row, col, nr, nc = (np.zeros(nb, np.int32) for _ in range(4))
        self.L.sgo_get_blocks(self.h, _p(row), _p(col), _p(nr), _p(nc))
        ph = np.zeros(self.g.P, np.int32)
        lh = np.zeros(self.g.L, np.int32)
        self.L.sgo_get_hessian_index(self.h, _p(ph), _p(lh))
        return dict(n_free=nf, n_blocks=nb, dim=self.L.sgo_scalar_dim(self.h), kind=kind, index=index, offset=off,
                    row=row, col=col, nrows=nr, ncols=nc, pose_hidx=ph, lm_hidx=lh)

    def linearize(self, jac_mode=JAC_G2O_NUMERIC):
        g = self.g
        out = dict(pp_err=np.zeros((g.n_pp, 3)), pp_A=np.zeros((g.n_pp, 3, 3)), pp_B=np.zeros((g.n_pp, 3, 3)),
                   pl_err=np.zeros((g.n_pl, 2)), pl_A=np.zeros((g.n_pl, 2, 3)), pl_B=np.zeros((g.n_pl, 2, 2)),
                   b=np.zeros(self.L.sgo_scalar_dim(self.h)), H=np.zeros(self.L.sgo_block_values_size(self.h)),
                   chi2=np.zeros(2))
        rc = self.L.sgo_linearize(self.h, jac_mode, _p(out["pp_err"]), _p(out["pp_A"]), _p(out["pp_B"]),
                                  _p(out["pl_err"]), _p(out["pl_A"]), _p(out["pl_B"]), _p(out["b"]), _p(out["H"]),
                                  _p(out["chi2"]))
        if rc != 0:
            raise RuntimeError("not initialised")
        return out

    def dense_hessian(self, lin, st=None):
        """Symmetric dense H from the block list (small graphs only)."""
        st = st or self.structure()
        n = st["dim"]
        H = np.zeros((n, n))
        o = 0
        for r, c, nr, nc in zip(st["row"], st["col"], st["nrows"], st["ncols"]):
            blk = lin["H"][o:o + nr * nc].reshape(nc, nr).T  # column-major
            o += nr * nc
            ro, co = st["offset"][r], st["offset"][c]
            H[ro:ro + nr, co:co + nc] = blk
            if r != c:
                H[co:co + nc, ro:ro + nr] = blk.T
        return H

    def optimize(self, iters, algo=ALGO_LM, jac_mode=JAC_G2O_NUMERIC):
        stats = (IterStat * max(1, iters))()
        n = self.L.sgo_optimize(self.h, algo, iters, jac_mode, C.cast(stats, C.c_void_p))
        return n, [dict(iteration=s.iteration, trials=s.trials, result=s.result, chi2=s.chi2, lambda_=s.lambda_,
                        rho=s.rho, chi2_before=s.chi2_before) for s in stats[:max(0, n if n > 0 else 0)]]

    def solve_once(self, lam, jac_mode=JAC_G2O_NUMERIC):
        x = np.zeros(self.L.sgo_scalar_dim(self.h))
        rc = self.L.sgo_solve_once(self.h, jac_mode, float(lam), _p(x))
        return rc == 0, x

    def estimates(self):
        p = np.zeros((self.g.P, 3))
        l = np.zeros((self.g.L, 2))
        self.L.sgo_get_estimates(self.h, _p(p), _p(l))
        return p, l

    def set_estimates(self, poses, lms):
        p = np.ascontiguousarray(poses, np.float64)
        l = np.ascontiguousarray(lms, np.float64)
        self.L.sgo_set_estimates(self.h, _p(p), _p(l))

    def chi2(self):
        c = np.zeros(2)
        self.L.sgo_chi2(self.h, _p(c))
        return float(c[0]), float(c[1])

    def profile(self):
        o = np.zeros(4)
        self.L.sgo_last_profile(self.h, _p(o))
        return dict(nnzL=o[0], t_lin=o[1], t_solve=o[2], t_total=o[3])


# ---------------------------------------------------------------------------------------------------------------
# rows either side of the optimiser (oracle/sgo_frontend.cpp)
def _f64(a):
    return np.ascontiguousarray(a, np.float64)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _i32(a):
    return np.ascontiguousarray(a, np.int32)


def pg_append(prev_pg_est, lm_est):
    """submap_loop_closer.cpp:206-223. lm_est [(count+1), 3], predecessor first -> (z [count,3], est [count,3])."""
    L = lib()
    lm = _f64(lm_est).reshape(-1, 3)
    count = lm.shape[0] - 1
    z, est = np.zeros((count, 3)), np.zeros((count, 3))
    prev = _f64(prev_pg_est)
    L.sgo_pg_append(_p(prev), _p(lm), C.c_int32(count), _p(z), _p(est))
    return z, est


def closure_chi2(est, ei, ej, z, info6):
    """log_runner.cpp:182-184: un-robustified chi2 of EdgeSE2 edges at `est`."""
    L = lib()
    est, ei, ej, z, info6 = _f64(est), _i32(ei), _i32(ej), _f64(z), _f64(info6)
    out = np.zeros(len(ei))
    L.sgo_closure_chi2(_p(est), _p(ei), _p(ej), _p(z), _p(info6), C.c_int32(len(ei)), _p(out))
    return out


def odom_information(deltas, seg_ptr, std_x, std_y, std_w):
    L = lib()
    deltas, seg_ptr = _f64(deltas), _i32(seg_ptr)
    n = len(seg_ptr) - 1
    z, cov, info = np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros((n, 6))
    L.sgo_odom_information(_p(deltas), _p(seg_ptr), C.c_int32(n), C.c_double(std_x), C.c_double(std_y), C.c_double(std_w),
                           _p(z), _p(cov), _p(info))
    return z, cov, info


def scan_point_covariances(deltas, beam_cos_sin, pts, std_x, std_y, std_w, var_r):
    """pts [n_windows, n_scans, scan_size, 2] float32; deltas [n_windows, n_scans-1, 3]."""
    L = lib()
    pts = _f32(pts)
    nw, ns, sz = pts.shape[:3]
    deltas, beam = _f64(deltas), _f32(beam_cos_sin)
    cov = np.zeros((nw, ns, sz, 4), np.float32)
    rt = np.zeros((nw, ns, sz, 2), np.float32)
    valid = np.zeros((nw, ns, sz), np.uint8)
    L.sgo_scan_point_covariances(_p(deltas), C.c_int32(nw), C.c_int32(ns), C.c_int32(sz), _p(beam), _p(pts),
                                 C.c_float(std_x), C.c_float(std_y), C.c_float(std_w), C.c_float(var_r), _p(cov), _p(rt),
                                 _p(valid))
    return cov, rt, valid


def line_fit_information(pts, pcov, seg_ptr):
    L = lib()
    pts, pcov, seg_ptr = _f32(pts), _f32(pcov), _i32(seg_ptr)
    n = len(seg_ptr) - 1
    rt, cov, info = np.zeros((n, 2), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 3))
    L.sgo_line_fit_information(_p(pts), _p(pcov), _p(seg_ptr), C.c_int32(n), _p(rt), _p(cov), _p(info))
    return rt, cov, info
