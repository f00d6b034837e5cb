"""Information-matrix producers on top of the C ABI: batch forms of what the reference computes one
key-frame / one segment at a time before it adds an edge.

    odom_information         include/odom_error_propagator.h:18-46, drone.cpp:84,127-128,143
    scan_point_covariances   src/multicloud2.cpp:56-83
    line_fit_information     src/ls_extractor/src/impl/smc.cpp:30-68, drone.cpp:203

Host arrays in and out; each call also returns the device time of its kernels (ms, CUDA events).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .optimizer import SgbError, _p


def _check(L, st):
    if st != capi.OK:
        raise SgbError(st, L.sgb_frontend_last_error().decode())


def odom_information(deltas, seg_ptr, std_x, std_y, std_w, device=-1, want_cov=True):
    L = capi.load()
    deltas = np.ascontiguousarray(deltas, np.float64)
    seg_ptr = np.ascontiguousarray(seg_ptr, np.int32)
    n = len(seg_ptr) - 1
    z, info = np.zeros((n, 3)), np.zeros((n, 6))
    cov = np.zeros((n, 3, 3)) if want_cov else None
    ms = C.c_double(0.0)
    _check(L, L.sgb_odom_information(device, _p(deltas), _p(seg_ptr), n, float(std_x), float(std_y), float(std_w), _p(z),
                                     _p(cov), _p(info), C.byref(ms)))
    return z, cov, info, ms.value


def scan_point_covariances(deltas, beam_cos_sin, pts, std_x, std_y, std_w, var_r, device=-1):
    L = capi.load()
    pts = np.ascontiguousarray(pts, np.float32)
    nw, ns, sz = pts.shape[:3]
    deltas = np.ascontiguousarray(deltas, np.float64)
    beam = np.ascontiguousarray(beam_cos_sin, np.float32)
    assert beam.size == 2 * sz and deltas.size == 3 * nw * (ns - 1)
    cov = np.zeros((nw, ns, sz, 4), np.float32)
    rt = np.zeros((nw, ns, sz, 2), np.float32)
    valid = np.zeros((nw, ns, sz), np.uint8)
    ms = C.c_double(0.0)
    _check(L, L.sgb_scan_point_covariances(device, _p(deltas), nw, ns, sz, _p(beam), _p(pts), float(std_x), float(std_y),
                                           float(std_w), float(var_r), _p(cov), _p(rt), _p(valid), C.byref(ms)))
    return cov, rt, valid, ms.value


def line_fit_information(pts, pcov, seg_ptr, device=-1):
    L = capi.load()
    pts = np.ascontiguousarray(pts, np.float32)
    pcov = np.ascontiguousarray(pcov, np.float32)
    seg_ptr = np.ascontiguousarray(seg_ptr, np.int32)
    n = len(seg_ptr) - 1
    rt, cov, info = np.zeros((n, 2), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 3))
    ms = C.c_double(0.0)
    _check(L, L.sgb_line_fit_information(device, _p(pts), _p(pcov), _p(seg_ptr), n, _p(rt), _p(cov), _p(info), C.byref(ms)))
    return rt, cov, info, ms.value
