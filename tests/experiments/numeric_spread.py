"""How reproducible is the reference's own LM-15 result under g2o's numeric Jacobians (delta = 1e-9)?

The CPU oracle is run twice on every BASELINE config: from the initial estimates, and from the initial estimates
perturbed by a few ulp (relative 1e-15 * N(0,1) on every coordinate of every free vertex). The spread of the two final
states is the sensitivity of the REFERENCE ALGORITHM to last-bit input differences -- the band inside which any other
realisation of the same algorithm (another compiler, another libm, the GPU path) can land. The same experiment in
analytic mode shows the band collapses once the central differences are out of the loop.

Also runs the host simulation of the product's row bodies + PCG (tests/hostsim) against the oracle, which is the CPU
preview of the GPU parity tests. Test infrastructure (uses oracle/): python tests/experiments/numeric_spread.py [names]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.cpu_oracle import ALGO_LM, JAC_ANALYTIC, JAC_G2O_NUMERIC, Oracle  # noqa: E402
from sparse_gslam_b200 import graphgen as gg  # noqa: E402


def pose_err(a, b):
    d = a - b
    d[:, 2] = gg.wrap(d[:, 2])
    return float(np.abs(d).max()) / max(1.0, float(np.abs(b[:, :2]).max()))


def rel_err(a, b):
    return float(np.abs(a - b).max()) / max(1.0, float(np.abs(b).max()))


def run_oracle(g, jac, p0=None, l0=None, iters=15):
    o = Oracle(g)
    o.initialize_optimization()
    if p0 is not None:
        o.set_estimates(p0, l0)
    n, st = o.optimize(iters, ALGO_LM, jac)
    p, l = o.estimates()
    return p, l, o.chi2()[0], n, st


def spread(g, jac, rel=1e-15, seed=0, iters=15):
    rng = np.random.default_rng(seed)
    pa, la, ca, _, _ = run_oracle(g, jac, iters=iters)
    p0 = g.pose_est * (1.0 + rel * rng.normal(size=g.pose_est.shape))
    p0[g.pose_fixed != 0] = g.pose_est[g.pose_fixed != 0]
    l0 = g.lm_est * (1.0 + rel * rng.normal(size=g.lm_est.shape))
    pb, lb, cb, _, _ = run_oracle(g, jac, p0, l0, iters=iters)
    return pose_err(pb, pa), rel_err(lb, la), abs(cb - ca) / ca


if __name__ == "__main__":
    names = sys.argv[1:] or ["c1", "c2", "c3", "c5s"]
    for name in names:
        g = gg.make_c5(rows=60, cols=60) if name == "c5s" else gg.make(name)
        for jac, jn in ((JAC_G2O_NUMERIC, "numeric"), (JAC_ANALYTIC, "analytic")):
            t0 = time.time()
            worst = [0.0, 0.0, 0.0]
            for seed in range(3):
                s = spread(g, jac, seed=seed)
                worst = [max(a, b) for a, b in zip(worst, s)]
            print(f"{name} {jn:8s} oracle self-spread (3 perturbations of 1e-15 relative): poses {worst[0]:.2e} "
                  f"landmarks {worst[1]:.2e} chi2 {worst[2]:.2e}   [{time.time() - t0:.1f} s]", flush=True)
