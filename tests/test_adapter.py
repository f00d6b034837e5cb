"""The g2o-facing adapter (sparse-gslam_b200/adapter/sgb_g2o_adapter.h): compile check on CPU, end-to-end run on GPU."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "sparse-gslam_b200", "adapter")


def _build():
    from sparse_gslam_b200 import build
    build.build()
    exe = os.path.join(ADAPTER, "example_graphs")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", os.path.join(ADAPTER, "example_graphs.cpp"),
                           "-I" + os.path.join(ROOT, "include"), "-L" + os.path.join(ROOT, "sparse-gslam_b200"), "-lsgb",
                           "-Wl,-rpath," + os.path.join(ROOT, "sparse-gslam_b200"), "-o", exe])
    return exe


def test_adapter_compiles_and_fails_loudly_without_gpu():
    exe = _build()
    from sparse_gslam_b200 import capi
    if capi.load().sgb_device_count() > 0:
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0                      # no silent CPU fallback
    assert "no CPU fallback" in r.stderr


def _parse_result(path):
    out = {"LM": {"POSE": {}, "LINE": {}}, "GN": {"POSE": {}, "LINE": {}}, "ON": {"POSE": {}, "LINE": {}}, "MARGINAL": []}
    for ln in open(path):
        f = ln.split()
        if f[0] == "MARGINAL":
            out["MARGINAL"].append((int(f[1]), int(f[2]), np.array(f[3:], float)))
        elif f[1] in ("ITERATIONS", "CHI2"):
            out[f[0]][f[1]] = float(f[2])
        else:
            out[f[0]][f[1]][int(f[2])] = np.array(f[3:], float)
    return out


@pytest.mark.gpu
def test_adapter_matches_oracle_on_the_reference_call_sequence(tmp_path):
    """graphs.cpp-style set-up through the g2o plugin surface, both optimisers: setup_lm_opt + the drone.cpp:146-165
    sequence (LM-15, g2o-numeric Jacobians = the adapter's default, like the reference) and setup_pose_opt +
    optimize(20) with DCS closures (submap_loop_closer.cpp:272-288). The example dumps the graphs it built; the CPU
    oracle optimises the same graphs; final estimates and chi2 agree to 1e-6 relative. computeMarginals (pure virtual
    in g2o's OptimizationAlgorithm) against the dense inverse of the oracle's Hessian."""
    from oracle.cpu_oracle import ALGO_GN, ALGO_LM, JAC_G2O_NUMERIC, Oracle
    from sparse_gslam_b200 import graphgen as gg
    exe = _build()
    prefix = str(tmp_path / "ex")
    r = subprocess.run([exe, prefix], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "optimize returned" in r.stdout and "pose graph: optimize returned 20" in r.stdout
    res = _parse_result(prefix + "_result.txt")
    # ---- landmark graph, LM-15
    g = gg.load_g2o(prefix + "_lm.g2o")
    o = Oracle(g)
    assert o.initialize_optimization()
    n, _ = o.optimize(15, ALGO_LM, JAC_G2O_NUMERIC)
    po, lo = o.estimates()
    assert res["LM"]["ITERATIONS"] >= 1 and n >= 1
    pa = np.array([res["LM"]["POSE"][int(i)] for i in g.pose_id])
    la = np.array([res["LM"]["LINE"][int(i)] for i in g.lm_id])
    scale = max(1.0, float(np.abs(po[:, :2]).max()))
    d = pa - po
    d[:, 2] = gg.wrap(d[:, 2])
    assert np.abs(d).max() / scale < 1e-6, np.abs(d).max() / scale
    assert np.abs(la - lo).max() / max(1.0, float(np.abs(lo).max())) < 1e-6
    assert abs(res["LM"]["CHI2"] - o.chi2()[0]) <= 1e-6 * o.chi2()[0]
    # ---- marginals: blocks of H^-1 at the final estimates (H from the oracle, dense inverse)
    o.set_estimates(pa, la)
    H = o.dense_hessian(o.linearize(JAC_G2O_NUMERIC))
    Hinv = np.linalg.inv(H)
    st = o.structure()
    off = st["offset"]
    dim = np.where(st["kind"] == 0, 3, 2)
    assert len(res["MARGINAL"]) == 3
    for (r_, c_, vals) in res["MARGINAL"]:
        blk = vals.reshape(dim[c_], dim[r_]).T   # column-major
        ref = Hinv[off[r_]:off[r_] + dim[r_], off[c_]:off[c_] + dim[c_]]
        assert np.abs(blk - ref).max() <= 1e-5 * np.abs(ref).max(), (r_, c_, blk, ref)
    # ---- the online path (updateInitialization + optimize(15, true), drone.cpp:152-156): the oracle optimises the dumped
    # extended graph from the same estimates with a full initializeOptimization
    assert "online key-frame" in r.stdout
    gon = gg.load_g2o(prefix + "_online.g2o")
    assert gon.P == g.P + 2 and len(gon.pl_pose) == len(g.pl_pose) + 4
    oo = Oracle(gon)
    assert oo.initialize_optimization()
    n_on, _ = oo.optimize(15, ALGO_LM, JAC_G2O_NUMERIC)
    assert n_on >= 1 and res["ON"]["ITERATIONS"] >= 1
    pon, lon = oo.estimates()
    pa2 = np.array([res["ON"]["POSE"][int(i)] for i in gon.pose_id])
    la2 = np.array([res["ON"]["LINE"][int(i)] for i in gon.lm_id])
    d = pa2 - pon
    d[:, 2] = gg.wrap(d[:, 2])
    assert np.abs(d).max() / max(1.0, float(np.abs(pon[:, :2]).max())) < 1e-6, np.abs(d).max()
    assert np.abs(la2 - lon).max() / max(1.0, float(np.abs(lon).max())) < 1e-6
    assert abs(res["ON"]["CHI2"] - oo.chi2()[0]) <= 1e-6 * oo.chi2()[0]
    # ---- pose graph, GN-20 + DCS
    gp = gg.load_g2o(prefix + "_pose.g2o")
    assert (gp.pp_phi > 0).sum() == 3 and float(gp.pp_phi.max()) == 0.75
    op = Oracle(gp)
    assert op.initialize_optimization()
    assert op.optimize(20, ALGO_GN)[0] == 20 == res["GN"]["ITERATIONS"]
    pp_, _ = op.estimates()
    pg = np.array([res["GN"]["POSE"][int(i)] for i in gp.pose_id])
    d = pg - pp_
    d[:, 2] = gg.wrap(d[:, 2])
    assert np.abs(d).max() / max(1.0, float(np.abs(pp_[:, :2]).max())) < 1e-6
    assert abs(res["GN"]["CHI2"] - op.chi2()[0]) <= 1e-6 * max(1e-12, op.chi2()[0])
