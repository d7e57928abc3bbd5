! pfem_b200.f90 -- ISO_C_BINDING interface to libpfemb200.so for the PFEMFort *parallelimpl1 drivers.
!
! A driver switches from PETSc to the B200 path by replacing
!     USE Module_SolverPetsc            (solverpetsc.F:28)
! with
!     USE Module_SolverB200
! and TYPE(PetscSolver) with TYPE(B200Solver).  The type-bound procedures keep the names of solverpetsc.F:94-103.
! The element loop of the value pass (tetrapoissonparallelimpl1.F:828-884) becomes one call of
! solver%assemble(elemData, timeData); see INTEGRATION.md for the full diff of tetrapoissonparallelimpl1.F.
!
! NOTE: no Fortran compiler exists in the build image, so this file is kept deliberately thin (pure interface blocks +
! one-line wrappers, names mirrored 1:1 from include/pfem_b200.h) and has not been compiled.  What has been done instead:
! it is parsed and EXECUTED by the Fortran front end of oracle/refrun (tests/test_fortran_module.py: every BIND(C) interface
! is checked against the prototype of the same name in pfem_b200.h -- argument count, VALUE vs address, C kind -- and against
! the symbols libpfemb200.so exports; tests/test_refrun_dropin.py: the reference's own tetrapoissonparallelimpl1.F with the
! INTEGRATION.md diff runs through this module).  Whoever has a toolchain:
! `gfortran -std=f2008 -fsyntax-only include/pfem_b200.f90` is still the first check (INTEGRATION.md).
      MODULE Module_SolverB200
      USE, INTRINSIC :: ISO_C_BINDING
      IMPLICIT NONE

      INTEGER, PARAMETER :: PFEM_POISSON_TRIA=0, PFEM_POISSON_TETRA=1
      INTEGER, PARAMETER :: PFEM_ELASTICITY_TRIA=2, PFEM_ELASTICITY_TETRA=3
      INTEGER, PARAMETER :: PFEM_PC_NONE=0, PFEM_PC_JACOBI=1
      INTEGER, PARAMETER :: PFEM_PC_BJACOBI_ILU0=2     ! the default, like solverpetsc.F:206 PCSetType(PCBJACOBI)
      INTEGER, PARAMETER :: PFEM_ERR_PATTERN=8         ! slow-path add outside the pattern (never dropped silently)

      INTERFACE
        INTEGER(C_INT) FUNCTION pfem_comm_unique_id(id128) BIND(C)
          IMPORT; CHARACTER(KIND=C_CHAR) :: id128(128)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_create(h, device, rank, nranks, id128) BIND(C)
          IMPORT; TYPE(C_PTR) :: h
          INTEGER(C_INT), VALUE :: device, rank, nranks
          CHARACTER(KIND=C_CHAR) :: id128(128)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_free(h) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_initialise(h, size_local, size_global, diag_nnz, offdiag_nnz) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          INTEGER(C_INT), VALUE :: size_local, size_global
          INTEGER(C_INT) :: diag_nnz(*), offdiag_nnz(*)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_set_options(h, rtol, abstol, dtol, max_it, pc_type) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          REAL(C_DOUBLE), VALUE :: rtol, abstol, dtol
          INTEGER(C_INT), VALUE :: max_it, pc_type
        END FUNCTION
        ! PetscInitialize(..., "petsc_options.dat") + KSPSetFromOptions/PCSetFromOptions; path is NUL-terminated
        INTEGER(C_INT) FUNCTION pfem_solver_set_options_from_file(h, path) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          CHARACTER(KIND=C_CHAR) :: path(*)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_set_mesh(h, kind, nElem, conn, nNode, coords, node_map_get_old) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          INTEGER(C_INT), VALUE :: kind, nElem, nNode
          INTEGER(C_INT) :: conn(nElem,*), node_map_get_old(*)      ! elemNodeConn(nElem,npElem): column-major = SoA
          REAL(C_DOUBLE) :: coords(nNode,*)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_set_pattern(h, nElem, nsize, elemDof) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          INTEGER(C_INT), VALUE :: nElem, nsize
          INTEGER(C_INT) :: elemDof(nElem,*)                        ! ElemDofArray(nElem,nsize)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_set_pattern_nodal(h, ndof, NodeDofArrayNew) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          INTEGER(C_INT), VALUE :: ndof
          INTEGER(C_INT) :: NodeDofArrayNew(*)                      ! NodeDofArrayNew(nNode,ndof)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_set_zero(h) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_set_applied(h, solnApplied, n) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          REAL(C_DOUBLE) :: solnApplied(*)
          INTEGER(C_INT), VALUE :: n
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_assemble(h, elemData, timeData, n_negative_jac) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          REAL(C_DOUBLE) :: elemData(*), timeData(*)
          INTEGER(C_INT) :: n_negative_jac
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_add_matrix(h, n, rows, cols, Klocal) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          INTEGER(C_INT), VALUE :: n
          INTEGER(C_INT) :: rows(*), cols(*)
          REAL(C_DOUBLE) :: Klocal(n,*)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_add_vector(h, n, rows, F) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          INTEGER(C_INT), VALUE :: n
          INTEGER(C_INT) :: rows(*)
          REAL(C_DOUBLE) :: F(*)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_add_value(h, row, val) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          INTEGER(C_INT), VALUE :: row
          REAL(C_DOUBLE), VALUE :: val
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_assemble_matrix(h, n, rindices, cindices, KLOCAL) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          INTEGER(C_INT), VALUE :: n
          INTEGER(C_INT) :: rindices(*), cindices(*)
          REAL(C_DOUBLE) :: KLOCAL(n,*)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_assemble_vector(h, n, rindices, FLOCAL) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          INTEGER(C_INT), VALUE :: n
          INTEGER(C_INT) :: rindices(*)
          REAL(C_DOUBLE) :: FLOCAL(*)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_assemble_matrix_and_vector(h, n, rindices, cindices, KLOCAL, FLOCAL) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          INTEGER(C_INT), VALUE :: n
          INTEGER(C_INT) :: rindices(*), cindices(*)
          REAL(C_DOUBLE) :: KLOCAL(n,*), FLOCAL(*)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_factorise(h) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_solve(h) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_factorise_and_solve(h) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_get_solution(h, x_global) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          REAL(C_DOUBLE) :: x_global(*)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_get_info(h, its, reason, rnorm, t_assemble, t_solve) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
          INTEGER(C_INT) :: its, reason
          REAL(C_DOUBLE) :: rnorm, t_assemble, t_solve
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_solver_print_info(h) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: h
        END FUNCTION
        ! element routines (elementutilitiespoisson.F:23,107; elasticity2D.F:23; elasticity3D.F:248)
        INTEGER(C_INT) FUNCTION pfem_poisson_tria_ke(x, y, elemData, timeData, valC, valDotC, K, F) BIND(C)
          IMPORT; REAL(C_DOUBLE) :: x(3), y(3), elemData(*), timeData(*), valC(3), valDotC(3), K(3,3), F(3)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_poisson_tetra_ke(x, y, z, elemData, timeData, valC, valDotC, K, F) BIND(C)
          IMPORT; REAL(C_DOUBLE) :: x(4), y(4), z(4), elemData(*), timeData(*), valC(4), valDotC(4), K(4,4), F(4)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_elasticity_tria_ke(x, y, elemData, timeData, valC, valDotC, K, F) BIND(C)
          IMPORT; REAL(C_DOUBLE) :: x(3), y(3), elemData(*), timeData(*), valC(6), valDotC(6), K(6,6), F(6)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_elasticity_tetra_ke(x, y, z, elemData, timeData, valC, valDotC, K, F) BIND(C)
          IMPORT; REAL(C_DOUBLE) :: x(4), y(4), z(4), elemData(*), timeData(*), valC(12), valDotC(12), K(12,12), F(12)
        END FUNCTION
        ! explicit dynamics: elementutilitieselasticity2D.F:158,283; elasticity3D.F:575,401; triaelasticityexplicit.F:881-1121
        INTEGER(C_INT) FUNCTION pfem_residual_elasticity_linear_tria(x, y, elemData, timeData, dispC, veloC, Flocal) BIND(C)
          IMPORT; REAL(C_DOUBLE) :: x(3), y(3), elemData(*), timeData(*), dispC(6), veloC(6), Flocal(6)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_mass_matrix_linear_tria(x, y, elemData, Mlocal) BIND(C)
          IMPORT; REAL(C_DOUBLE) :: x(3), y(3), elemData(*), Mlocal(6)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_residual_elasticity_linear_tetra(x, y, z, elemData, timeData, valC, valDotC, Flocal) BIND(C)
          IMPORT; REAL(C_DOUBLE) :: x(4), y(4), z(4), elemData(*), timeData(*), valC(12), valDotC(12), Flocal(12)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_mass_matrix_linear_tetra(x, y, z, elemData, Mlocal) BIND(C)
          IMPORT; REAL(C_DOUBLE) :: x(4), y(4), z(4), elemData(*), Mlocal(12)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_explicit_create(ex, device) BIND(C)
          IMPORT; TYPE(C_PTR) :: ex
          INTEGER(C_INT), VALUE :: device
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_explicit_free(ex) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: ex
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_explicit_set_mesh(ex, kind, nElem, conn, nNode, coords) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: ex
          INTEGER(C_INT), VALUE :: kind, nElem, nNode
          INTEGER(C_INT) :: conn(nElem,*)
          REAL(C_DOUBLE) :: coords(nNode,*)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_explicit_set_free_dofs(ex, size_global, assyForSoln) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: ex
          INTEGER(C_INT), VALUE :: size_global
          INTEGER(C_INT) :: assyForSoln(*)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_explicit_lumped_mass(ex, elemData) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: ex
          REAL(C_DOUBLE) :: elemData(*)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_explicit_advance(ex, nsteps, dt, elemData, timeData) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: ex
          INTEGER(C_INT), VALUE :: nsteps
          REAL(C_DOUBLE), VALUE :: dt
          REAL(C_DOUBLE) :: elemData(*), timeData(*)
        END FUNCTION
        ! disp / dispPrev2 / velo / acce / mass: node-slot arrays (nNode*ndof), like the driver's plain arrays
        INTEGER(C_INT) FUNCTION pfem_explicit_get_state(ex, disp, dispPrev2, velo, acce, mass) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: ex
          REAL(C_DOUBLE) :: disp(*), dispPrev2(*), velo(*), acce(*), mass(*)
        END FUNCTION
        INTEGER(C_INT) FUNCTION pfem_explicit_set_state(ex, disp, dispPrev2) BIND(C)
          IMPORT; TYPE(C_PTR), VALUE :: ex
          REAL(C_DOUBLE) :: disp(*), dispPrev2(*)
        END FUNCTION
      END INTERFACE

      TYPE B200Solver
        TYPE(C_PTR) :: h = C_NULL_PTR
        INTEGER :: ierr = 0
      CONTAINS
        PROCEDURE :: create
        PROCEDURE :: setOptionsFromFile
        PROCEDURE :: initialise
        PROCEDURE :: setZero
        PROCEDURE :: free
        PROCEDURE :: printInfo
        PROCEDURE :: assembleMatrix
        PROCEDURE :: assembleVector
        PROCEDURE :: assembleMatrixAndVector
        PROCEDURE :: factorise
        PROCEDURE :: solve
        PROCEDURE :: factoriseAndSolve
        PROCEDURE :: assemble
      END TYPE B200Solver

      CONTAINS

      SUBROUTINE check(ierr, where)
        INTEGER, INTENT(IN) :: ierr
        CHARACTER(LEN=*), INTENT(IN) :: where
        IF (ierr /= 0) THEN
          WRITE(*,*) " libpfemb200 error ", ierr, " in ", where
          STOP " Aborting... in Module_SolverB200 "
        END IF
      END SUBROUTINE check

      ! id128: the 128-byte communicator id of rank 0 (pfem_comm_unique_id, broadcast by the driver); not needed for nranks = 1
      SUBROUTINE create(this, device, rank, nranks, id128)
        CLASS(B200Solver) :: this
        INTEGER, INTENT(IN) :: device, rank, nranks
        CHARACTER(KIND=C_CHAR), OPTIONAL :: id128(128)
        CHARACTER(KIND=C_CHAR) :: no_id(128)
        IF (PRESENT(id128)) THEN
          call check(pfem_solver_create(this%h, device, rank, nranks, id128), "create")
        ELSE
          no_id = C_NULL_CHAR
          call check(pfem_solver_create(this%h, device, rank, nranks, no_id), "create")
        END IF
      END SUBROUTINE create

      ! after initialise: the options file PetscInitialize would read (tetrapoissonparallelimpl1.F:168)
      SUBROUTINE setOptionsFromFile(this, path)
        CLASS(B200Solver) :: this
        CHARACTER(LEN=*), INTENT(IN) :: path
        call check(pfem_solver_set_options_from_file(this%h, TRIM(path)//C_NULL_CHAR), "setOptionsFromFile")
      END SUBROUTINE setOptionsFromFile

      SUBROUTINE initialise(this, size_local, size_global, diag_nnz, offdiag_nnz)
        CLASS(B200Solver) :: this
        INTEGER, INTENT(IN) :: size_global, size_local
        INTEGER, DIMENSION(:) :: diag_nnz, offdiag_nnz
        call check(pfem_solver_initialise(this%h, size_local, size_global, diag_nnz, offdiag_nnz), "initialise")
      END SUBROUTINE initialise

      SUBROUTINE setZero(this)
        CLASS(B200Solver) :: this
        call check(pfem_solver_set_zero(this%h), "setZero")
      END SUBROUTINE setZero

      SUBROUTINE free(this)
        CLASS(B200Solver) :: this
        call check(pfem_solver_free(this%h), "free")
        this%h = C_NULL_PTR
      END SUBROUTINE free

      SUBROUTINE printInfo(this)
        CLASS(B200Solver) :: this
        call check(pfem_solver_print_info(this%h), "printInfo")
      END SUBROUTINE printInfo

      SUBROUTINE assembleMatrix(this, RINDICES, CINDICES, KLOCAL)
        CLASS(B200Solver) :: this
        INTEGER, DIMENSION(:) :: RINDICES, CINDICES
        DOUBLE PRECISION, DIMENSION(:,:) :: KLOCAL
        call check(pfem_solver_assemble_matrix(this%h, size(RINDICES), RINDICES, CINDICES, KLOCAL), "assembleMatrix")
      END SUBROUTINE assembleMatrix

      SUBROUTINE assembleVector(this, RINDICES, FLOCAL)
        CLASS(B200Solver) :: this
        INTEGER, DIMENSION(:) :: RINDICES
        DOUBLE PRECISION, DIMENSION(:) :: FLOCAL
        call check(pfem_solver_assemble_vector(this%h, size(RINDICES), RINDICES, FLOCAL), "assembleVector")
      END SUBROUTINE assembleVector

      SUBROUTINE assembleMatrixAndVector(this, RINDICES, CINDICES, KLOCAL, FLOCAL)
        CLASS(B200Solver) :: this
        INTEGER, DIMENSION(:) :: RINDICES, CINDICES
        DOUBLE PRECISION, DIMENSION(:,:) :: KLOCAL
        DOUBLE PRECISION, DIMENSION(:) :: FLOCAL
        call check(pfem_solver_assemble_matrix_and_vector(this%h, size(RINDICES), RINDICES, CINDICES, KLOCAL, FLOCAL), &
                   "assembleMatrixAndVector")
      END SUBROUTINE assembleMatrixAndVector

      SUBROUTINE factorise(this)
        CLASS(B200Solver) :: this
        call check(pfem_solver_factorise(this%h), "factorise")
      END SUBROUTINE factorise

      SUBROUTINE solve(this)
        CLASS(B200Solver) :: this
        call check(pfem_solver_solve(this%h), "solve")
      END SUBROUTINE solve

      SUBROUTINE factoriseAndSolve(this)
        CLASS(B200Solver) :: this
        call check(pfem_solver_factorise_and_solve(this%h), "factoriseAndSolve")
      END SUBROUTINE factoriseAndSolve

      ! the whole value pass of the drivers in one call
      SUBROUTINE assemble(this, elemData, timeData)
        CLASS(B200Solver) :: this
        DOUBLE PRECISION, DIMENSION(:) :: elemData, timeData
        INTEGER(C_INT) :: nneg
        call check(pfem_solver_assemble(this%h, elemData, timeData, nneg), "assemble")
      END SUBROUTINE assemble

      END MODULE Module_SolverB200
