"""sparse-gslam_b200: B200-native pose-graph optimisation backend for sparse-gslam's 2D graph SLAM.

The product is libsgb.so (CUDA kernels for sm_100a behind the C ABI of include/sgb_capi.h); this package holds the
ctypes binding, the g2o-call mirror used by the tests and the bench, the synthetic graph generators and the build
recipe. Importing the package does not need a GPU; creating an optimiser does, and fails loudly without one.
"""
from . import capi, graphgen  # noqa: F401
from .capi import ALGO_GN, ALGO_LM, JAC_ANALYTIC, JAC_G2O_NUMERIC  # noqa: F401
from .optimizer import SgbError, SparseOptimizerB200  # noqa: F401
from .posegraph import PoseGraphB200  # noqa: F401
from . import frontend  # noqa: F401

__all__ = ["capi", "graphgen", "frontend", "SparseOptimizerB200", "PoseGraphB200", "SgbError", "ALGO_LM", "ALGO_GN", "JAC_G2O_NUMERIC",
           "JAC_ANALYTIC"]
