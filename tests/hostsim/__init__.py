"""ctypes binding of the host-side test harness (tests/hostsim/hostsim.cpp). TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from sparse_gslam_b200 import capi
from sparse_gslam_b200.optimizer import _p, pack_graph

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libhostsim.so")
_lib = None


def build():
    csrc = os.path.join(ROOT, "sparse-gslam_b200", "csrc")
    srcs = [os.path.join(HERE, "hostsim.cpp"), os.path.join(HERE, "hostsim_edits.cpp"),
            os.path.join(csrc, "sgb_structure.cpp"), os.path.join(csrc, "sgb_partition.cpp")]
    deps = srcs + [os.path.join(csrc, f) for f in
                   ("sgb_rows.h", "sgb_math.h", "sgb_types.h", "sgb_structure.h", "sgb_partition.h", "sgb_edits.h",
                    "sgb_coarse.h")]
    def fresh():
        return os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps)

    if fresh():
        return SO
    from sparse_gslam_b200._buildlock import build_lock
    with build_lock(SO) as tmp:  # the gloo world-2 test builds from two processes at once
        if not fresh():
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-pthread",
                                   "-o", tmp] + srcs)
    return SO


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.hs_create.restype = vp
        L.hs_create.argtypes = [C.POINTER(capi.GraphSoA), C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.hs_destroy.argtypes = [vp]
        L.hs_error.argtypes = [vp]
        L.hs_error.restype = C.c_char_p
        L.hs_info.argtypes = [vp, C.POINTER(capi.StructureInfo)]
        L.hs_structure.argtypes = [vp] + [vp] * 9
        L.hs_sell_stats.argtypes = [vp, vp]
        L.hs_linearize.argtypes = [vp, vp, vp, vp]
        L.hs_solve_once.argtypes = [vp, C.c_double, vp, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.hs_optimize.argtypes = [vp, C.c_int, C.c_int, vp]
        L.hs_get_estimates.argtypes = [vp, C.c_int, vp, vp]
        L.hs_partition_stats.argtypes = [vp, vp]
        L.hs_check_hlp.argtypes = [vp]
        L.hs_check_hlp.restype = C.c_double
        L.hs_chi2.argtypes = [vp, vp]
        L.hs_use_ghost_landmarks.argtypes = [C.c_int]
        L.hs_use_filtered_structure.argtypes = [C.c_int]
        L.hs_time_symbolic.argtypes = [C.POINTER(capi.GraphSoA), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.hs_time_symbolic.restype = C.c_double
        L.hs_preconditioner.argtypes = [vp, C.c_double, vp]
        L.hs_use_coarse.argtypes = [C.c_int]
        L.hs_coarse.argtypes = [vp, vp, vp]
        L.hs_pg_append.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
        L.hs_closure_chi2.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp]
        L.hs_odom_information.argtypes = [vp, vp, C.c_int, C.c_double, C.c_double, C.c_double, vp, vp, vp]
        L.hs_scan_point_covariances.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_float, C.c_float, C.c_float,
                                                C.c_float, vp, vp, vp]
        L.hs_line_fit_information.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp]
        _lib = L
    return _lib


def use_ghost_landmarks(on: bool):
    """Planner switch (sgb_partition.h): ghost copies of the landmark rows other ranks own. Process-global."""
    lib().hs_use_ghost_landmarks(int(bool(on)))


def use_coarse(max_nodes: int):
    """Two-level preconditioner of the resident solve (sgb_coarse.h): at most `max_nodes` coarse nodes, 0 = off."""
    lib().hs_use_coarse(int(max_nodes))


def use_filtered_structure(on: bool):
    """Every virtual rank builds its own rank-filtered structure (what libsgb does on > 1 GPUs with ghost rows)."""
    lib().hs_use_filtered_structure(int(bool(on)))


def time_symbolic(g, world, rank, filtered):
    """(ms, local pose-pose edges, local pose-line edges) of one rank's host symbolic phase."""
    s, keep = pack_graph(g)
    a, b = C.c_int(), C.c_int()
    ms = lib().hs_time_symbolic(C.byref(s), world, rank, int(filtered), C.byref(a), C.byref(b))
    return ms, a.value, b.value


class HostSim:
    def __init__(self, g, jac_numeric=True, tol=1e-10, maxit=0, world=1):
        self.L = lib()
        s, keep = pack_graph(g)
        st = C.c_int()
        self.world = world
        self.h = C.c_void_p(self.L.hs_create(C.byref(s), int(jac_numeric), tol, maxit, world, C.byref(st)))
        self.status = st.value
        self.error = self.L.hs_error(self.h).decode()
        self.P, self.Lm = s.n_poses, s.n_landmarks

    def __del__(self):
        try:
            self.L.hs_destroy(self.h)
        except Exception:
            pass

    def structure(self):
        info = capi.StructureInfo()
        self.L.hs_info(self.h, C.byref(info))
        nf, nb = info.n_free, info.n_blocks
        kind, index, off = (np.zeros(nf, np.int32) for _ in range(3))
        row, col, nr, nc = (np.zeros(nb, np.int32) for _ in range(4))
        ph, lh = np.zeros(self.P, np.int32), np.zeros(self.Lm, np.int32)
        self.L.hs_structure(self.h, _p(kind), _p(index), _p(off), _p(row), _p(col), _p(nr), _p(nc), _p(ph), _p(lh))
        return dict(n_free=nf, n_blocks=nb, dim=info.scalar_dim, kind=kind, index=index, offset=off, row=row, col=col,
                    nrows=nr, ncols=nc, pose_hidx=ph, lm_hidx=lh, block_values=info.block_values)

    def sell_stats(self):
        o = np.zeros(6, np.int64)
        self.L.hs_sell_stats(self.h, _p(o))
        return dict(hpp=(int(o[0]), int(o[1])), hpl=(int(o[2]), int(o[3])), hlp=(int(o[4]), int(o[5])))

    def linearize(self):
        st = self.structure()
        b, H, chi = np.zeros(st["dim"]), np.zeros(st["block_values"]), np.zeros(2)
        self.L.hs_linearize(self.h, _p(b), _p(H), _p(chi))
        return dict(b=b, H=H, chi2=chi)

    def solve_once(self, lam):
        st = self.structure()
        x = np.zeros(st["dim"])
        it, rel = C.c_int(), C.c_double()
        flag = self.L.hs_solve_once(self.h, float(lam), _p(x), C.byref(it), C.byref(rel))
        return flag, x, it.value, rel.value

    def optimize(self, iters, algo):
        stats = (capi.IterStat * max(1, iters))()
        n = self.L.hs_optimize(self.h, algo, iters, C.cast(stats, C.c_void_p))
        return n, [stats[i].as_dict() for i in range(max(n, 0))]

    def estimates(self, rank=0):
        p, l = np.zeros((self.P, 3)), np.zeros((self.Lm, 2))
        self.L.hs_get_estimates(self.h, rank, _p(p), _p(l))
        return p, l

    def partition_stats(self):
        o = np.zeros((self.world, 11), np.int64)
        self.L.hs_partition_stats(self.h, _p(o))
        keys = ("nP", "nL", "n_pp", "n_pl", "n_pp_owned", "n_pl_owned", "halo_p", "halo_t", "nL_owned", "remote_cols", "nH")
        return [dict(zip(keys, map(int, row))) for row in o]

    def coarse(self):
        """(node spacing h, nodes, failed, inverse coarse matrix of the last solve or None): sgb_coarse.h"""
        info = np.zeros(3, np.int32)
        self.L.hs_coarse(self.h, _p(info), None)
        if info[0] == 0:
            return 0, 0, 0, None
        out = np.zeros((3 * int(info[1]), 3 * int(info[1])))
        self.L.hs_coarse(self.h, _p(info), _p(out))
        return int(info[0]), int(info[1]), int(info[2]), out

    def preconditioner(self, lam, n_free_poses):
        """Rows of the inverse 12x12 Schur diagonal blocks as the kernels store them: [nP, 3, 12] float32 (rank 0)."""
        out = np.zeros((n_free_poses, 3, 12), np.float32)
        rc = self.L.hs_preconditioner(self.h, float(lam), _p(out))
        return rc == 0, out

    def check_hlp(self):
        return float(self.L.hs_check_hlp(self.h))

    def chi2(self):
        c = np.zeros(2)
        self.L.hs_chi2(self.h, _p(c))
        return float(c[0]), float(c[1])


# ---- bodies of sgb_edits.h with the kernels' work decomposition (hostsim_edits.cpp)
def _c(a, dt):
    return np.ascontiguousarray(a, dt)


def pg_append(prev, lm_est, threads=256, items=8, scan_threads=512):
    lm = _c(lm_est, np.float64).reshape(-1, 3)
    n = lm.shape[0] - 1
    z, est = np.zeros((n, 3)), np.zeros((n, 3))
    prev = _c(prev, np.float64)
    lib().hs_pg_append(_p(prev), _p(lm), n, threads, items, scan_threads, _p(z), _p(est))
    return z, est


def closure_chi2(est, ei, ej, z, info6):
    est, ei, ej, z, info6 = _c(est, np.float64), _c(ei, np.int32), _c(ej, np.int32), _c(z, np.float64), _c(info6, np.float64)
    out = np.zeros(len(ei))
    lib().hs_closure_chi2(_p(est), _p(ei), _p(ej), _p(z), _p(info6), len(ei), _p(out))
    return out


def odom_information(deltas, seg_ptr, std_x, std_y, std_w):
    deltas, seg_ptr = _c(deltas, np.float64), _c(seg_ptr, np.int32)
    n = len(seg_ptr) - 1
    z, cov, info = np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros((n, 6))
    lib().hs_odom_information(_p(deltas), _p(seg_ptr), n, std_x, std_y, std_w, _p(z), _p(cov), _p(info))
    return z, cov, info


def scan_point_covariances(deltas, beam, pts, std_x, std_y, std_w, var_r):
    pts = _c(pts, np.float32)
    nw, ns, sz = pts.shape[:3]
    deltas, beam = _c(deltas, np.float64), _c(beam, np.float32)
    cov, rt = np.zeros((nw, ns, sz, 4), np.float32), np.zeros((nw, ns, sz, 2), np.float32)
    valid = np.zeros((nw, ns, sz), np.uint8)
    lib().hs_scan_point_covariances(_p(deltas), nw, ns, sz, _p(beam), _p(pts), std_x, std_y, std_w, var_r, _p(cov), _p(rt),
                                    _p(valid))
    return cov, rt, valid


def line_fit_information(pts, pcov, seg_ptr):
    pts, pcov, seg_ptr = _c(pts, np.float32), _c(pcov, np.float32), _c(seg_ptr, np.int32)
    n = len(seg_ptr) - 1
    rt, cov, info = np.zeros((n, 2), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 3))
    lib().hs_line_fit_information(_p(pts), _p(pcov), _p(seg_ptr), n, _p(rt), _p(cov), _p(info))
    return rt, cov, info
