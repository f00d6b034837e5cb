"""Short runs of the hot path for compute-sanitizer (memcheck / racecheck): a few LM iterations of one workload through
the C ABI. usage: python tools/sanitize_run.py <small|c1|c5s|stream|coarse> [iterations]
(coarse = 150 / 300 key-frames of C1 with the two-level preconditioner on: one CTA, then a cluster)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sparse_gslam_b200 import SparseOptimizerB200, capi  # noqa: E402
from sparse_gslam_b200 import graphgen as gg  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "small"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if name == "stream":
    from sparse_gslam_b200.session import GpuBackend, LandmarkGraphSession, stream_from_graph
    frames = stream_from_graph(gg.make("c1"))[:40]
    s = LandmarkGraphSession(GpuBackend(jacobian_mode=capi.JAC_ANALYTIC), iters=iters)
    s.run(frames)
    print("stream ok:", len(s.log), "key-frames, last chi2", s.log[-1].chi2)
elif name == "coarse":
    for n in ([int(a) for a in sys.argv[3:]] or [150, 300]):   # 150 / 300: clusters of two CTAs; <= 100: one CTA
        g = gg.make("c1").chain_prefix(n)
        opt = SparseOptimizerB200(capi.ALGO_LM, jacobian_mode=capi.JAC_ANALYTIC, coarse_nodes=40)
        assert opt.initialize_optimization(g)
        k, st = opt.optimize(iters)
        print("coarse", n, "nodes", opt.structure_info()["coarse_nodes"], "ok: iterations", k, "chi2", st[-1]["chi2"], "pcg", [s["pcg_iters"] for s in st])
else:
    g = gg.make_c5(rows=60, cols=60) if name == "c5s" else gg.make(name)
    opt = SparseOptimizerB200(capi.ALGO_LM, jacobian_mode=capi.JAC_ANALYTIC)
    assert opt.initialize_optimization(g)
    n, st = opt.optimize(iters)
    print(name, "ok: iterations", n, "chi2", st[-1]["chi2"], "pcg", [s["pcg_iters"] for s in st])
    gp = g.pose_only(phi=1.0)
    gn = SparseOptimizerB200(capi.ALGO_GN)
    assert gn.initialize_optimization(gp)
    print(name, "pose graph GN:", gn.optimize(2)[0])
