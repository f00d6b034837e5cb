"""Diagnostic: the key-frame stream with the two-level preconditioner on ONE handle, several passes (the bench's warm-up
+ timed passes re-initialise the same handle from a tiny graph after a 400-pose one). usage: coarse_repro.py <frames> <passes>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparse_gslam_b200 import capi  # noqa: E402
from sparse_gslam_b200 import graphgen as gg  # noqa: E402
from sparse_gslam_b200.session import GpuBackend, LandmarkGraphSession, stream_from_graph  # noqa: E402

n, passes = int(sys.argv[1]), int(sys.argv[2])
frames = stream_from_graph(gg.make("c1"))[:n]
be = GpuBackend(jacobian_mode=capi.JAC_ANALYTIC, coarse_nodes=int(os.environ.get("NODES", "40")))
for p in range(passes):
    s = LandmarkGraphSession(be)
    for k, kf in enumerate(frames):
        try:
            s.add_keyframe(kf)
        except Exception as e:
            print("pass", p, "key-frame", k, "FAILED:", e, "| graph:", len(s.pose_est), "poses", len(s.lm_est), "landmarks",
                  "| planned coarse nodes", None, flush=True)
            raise
    print("pass", p, "ok:", len(s.log), "key-frames, pcg iterations", be.prof["pcg_iters"], "coarse nodes at the end",
          be.opt.structure_info()["coarse_nodes"], flush=True)
