"""dune_hdd_b200 - B200-native SWIPDG assembly / CG / estimator path behind dune-hdd's discretization API.

Layout: ``csrc/`` holds the sm_100a CUDA kernels and the C-ABI (include/hdd_b200.h); the Python modules mirror
the reference interface for this path (grids, problems, discretizations, estimators, testcases, studies) and the
configuration-file driver / VTK output around it (discreteproblem, vtk).
"""
from . import capi, discreteproblem, discretizations, estimators, grids, parallel, problems, studies, testcases, vtk  # noqa: F401
from .discretizations import SWIPDG, BlockSWIPDG  # noqa: F401

__all__ = ["capi", "grids", "problems", "discretizations", "estimators", "testcases", "studies", "parallel", "discreteproblem", "vtk",
           "SWIPDG", "BlockSWIPDG"]
