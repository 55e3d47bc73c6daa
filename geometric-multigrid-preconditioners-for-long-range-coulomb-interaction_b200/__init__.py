"""B200-native MG-PCG + Gaussian-charge RHS path (drop-in for `LaplaceProblem::solve()` and the RHS
assembly of vinayak-gholap1993/Geometric-Multigrid-preconditioners-for-long-range-Coulomb-interaction).

Layout: `csrc/` hand-written sm_100a CUDA kernels + the C ABI (`include/gmg_b200.h`),
`host/` the C++ host side mirroring `Step50::LaplaceProblem`, `capi.py` the ctypes binding used by
tests and bench.  No CPU fallback: everything here needs the compiled library and a B200.
"""
from . import capi, hostapi, lattice  # noqa: F401
from .capi import Gmg, GmgError, NoConvergence, load_library  # noqa: F401
