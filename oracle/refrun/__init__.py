"""oracle/refrun -- execute the reference's own Fortran.  TEST INFRASTRUCTURE ONLY (nothing under pfemfort_b200/ imports it).

    fortran_to_py.py   Fortran (fixed form .F, free form .f90) -> Python: program units, TYPEs, INTERFACE / BIND(C), the
                       statements and intrinsics the reference's hot path uses; refuses everything else
    runtime.py         Fortran semantics at run time: typed scalars, integer division, MATMUL order, list-directed I/O,
                       by-reference scalars, NaN / sentinel fill, ctypes binding of BIND(C) interfaces
    mocks.py           what is external to the reference: MPI (P ranks = P threads), PETSc Vec / Mat / KSP, METIS, VTK writer
    run_reference.py   run a `*parallelimpl1` / explicit PROGRAM from /root/reference/src on P simulated ranks
    dropin.py          the reference's PROGRAM text + the INTEGRATION.md diff, executed against the SolverB200 interface or
                       through include/pfem_b200.f90 itself
    __main__.py        python -m oracle.refrun <driver.F> <inputs...> [--ranks P]

Outputs derived from the reference's sources (translated programs) go to oracle/_ref/ (git-ignored); what the runs
PRODUCE is committed as tests/golden/ref_* by tests/golden/make_reference_vectors.py.  DESIGN.md section 2 has the account.
"""
