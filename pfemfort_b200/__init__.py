"""pfemfort_b200: B200-native (sm_100a CUDA + NCCL) implementation of PFEMFort's implicit hot path.

Product = ``libpfemb200.so`` (CUDA kernels behind the C ABI of ``include/pfem_b200.h``).  This package is
the host-side mirror of the reference interface: ``solver.SolverB200`` (TYPE PetscSolver), ``driver`` (the
*parallelimpl1 PROGRAM bodies) and ``mesh`` (the text formats and generator recipes).
"""
from . import driver, mesh, solver  # noqa: F401
from .solver import SolverB200, PfemError  # noqa: F401

__all__ = ["driver", "mesh", "solver", "SolverB200", "PfemError"]
