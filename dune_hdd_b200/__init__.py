"""dune_hdd_b200 - B200-native SWIPDG assembly / CG / estimator path behind dune-hdd's discretization API.

Layout: ``csrc/`` holds the sm_100a CUDA kernels and the C-ABI (include/hdd_b200.h); the Python modules mirror
the reference interface for this path (grids, problems, discretizations, estimators, testcases).
"""
from . import capi, discretizations, estimators, grids, parallel, problems, studies, testcases  # noqa: F401
from .discretizations import SWIPDG, BlockSWIPDG  # noqa: F401

__all__ = ["capi", "grids", "problems", "discretizations", "estimators", "testcases", "studies", "parallel", "SWIPDG", "BlockSWIPDG"]
