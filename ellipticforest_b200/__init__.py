"""ellipticforest_b200: B200-native Hierarchical Poincare-Steklov hot path behind EllipticForest's API.

The product is the CUDA library (csrc/ -> libefgpu.so, C-ABI in include/efgpu.h).  This package is
the thin host-side mirror of the reference interface used by the tests and the benchmark.
"""
from .hps import (CACHE_OPERATORS, HOMOGENEOUS_RHS, KEEP_X, FiniteVolumeGrid, FiniteVolumeSolver, HPSAlgorithm, Mesh)  # noqa: F401
from ._lib import EfgpuError, LIB_PATH, load  # noqa: F401
