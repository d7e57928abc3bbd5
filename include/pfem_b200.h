/*
 * pfem_b200.h -- C ABI of libpfemb200.so: the B200 (sm_100a) drop-in for the implicit hot path of
 * chennachaos/PFEMFort (element Ke/Fe -> sparse assembly + Dirichlet lifting -> Jacobi-CG).
 *
 * Every entry point replaces one piece of the reference's Fortran call surface; the file:line it
 * replaces is cited beside it (paths are relative to the reference's src/).  All functions are
 * extern "C", take plain host pointers and sizes, and return an int status (0 = PFEM_OK).  The
 * library owns all device memory.  One process drives one GPU (exactly like one MPI rank drives
 * one PETSc sub-domain); ranks are joined by an NCCL communicator over NVLink.  A handle is not
 * re-entrant.  There is no CPU fallback: every compute entry point fails with PFEM_ERR_CUDA when
 * no sm_100 device is usable.
 *
 * Array conventions are the Fortran drivers' own ("SoA on the wire", column-major):
 *   conn[i*nElem + e]    = elemNodeConn(e+1, i+1), 1-based NEW node ids
 *   coords[c*nNode + n]  = coords(n+1, c+1), OLD node numbering
 *   elemDof[k*nElem + e] = ElemDofArray(e+1, k+1), 0-based global dof ids, -1 = Dirichlet
 *   K[i + nsize*j]       = Klocal(i+1, j+1)
 */
#ifndef PFEM_B200_H
#define PFEM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pfem_solver pfem_solver_t;

/* status codes */
enum {
    PFEM_OK = 0,
    PFEM_ERR_CUDA = 1,          /* CUDA runtime error or no usable device */
    PFEM_ERR_ARG = 2,           /* bad argument */
    PFEM_ERR_STATE = 3,         /* solver state machine violated (the reference STOPs, solverpetsc.F:415-419,441-445) */
    PFEM_ERR_NEG_JACOBIAN = 4,  /* the reference STOPs (elementutilitiespoisson.F:71,157; elasticity2D.F:90; elasticity3D.F:320) */
    PFEM_ERR_NCCL = 5,
    PFEM_ERR_SIZE = 6,          /* 32-bit index range exceeded */
    PFEM_ERR_NUMBERING = 7,     /* inconsistent dof numbering (tetrapoissonparallelimpl1.F:614-616,650-655) */
    PFEM_ERR_PATTERN = 8        /* a slow-path add hit a matrix location the pattern pass did not create (PETSc would insert it under
                                   MAT_NEW_NONZERO_LOCATIONS, solverpetsc.F:171; here the structure is fixed): reported, never dropped silently */
};

/* solver states, solverpetsc.F:64-68 */
enum { PFEM_SOLVER_EMPTY = 1, PFEM_PATTERN_OK = 2, PFEM_INIT_OK = 3, PFEM_ASSEMBLY_OK = 4, PFEM_FACTORISE_OK = 5 };

/* preconditioners (PCSetFromOptions, solverpetsc.F:202-210); the north-star run uses -pc_type jacobi */
enum { PFEM_PC_NONE = 0, PFEM_PC_JACOBI = 1,
       PFEM_PC_BJACOBI_ILU0 = 2 /* the reference's default: PCBJACOBI (solverpetsc.F:206), one block per rank, sub-PC ILU(0) */ };

/* value-pass kernel selection (pfem_solver_set_assembly_mode; the environment variable PFEM_ASM overrides it) */
enum { PFEM_ASM_AUTO = 0, PFEM_ASM_ROWS = 1, PFEM_ASM_FAST = 2 };
/* FP64 instructions per element visit of the tile kernel, counted in the SASS of this build (profiles/r02_sass_value_pass.txt) */
#define PFEM_FP64_PER_VISIT_CTILE_TET 0.0
#define PFEM_FP64_PER_VISIT_CTILE_TRIA 0.0
/* FP64 instructions per (row, element) incidence of the default (FMA) row-gather kernel, same source */
#define PFEM_FP64_PER_INCIDENCE_FAST_TET 0.0
#define PFEM_FP64_PER_INCIDENCE_FAST_TRIA 0.0

/* element kinds */
enum { PFEM_POISSON_TRIA = 0, PFEM_POISSON_TETRA = 1, PFEM_ELASTICITY_TRIA = 2, PFEM_ELASTICITY_TETRA = 3 };

/* KSPConvergedReason values reported by pfem_solver_get_info (PETSc numbering) */
enum {
    PFEM_CONVERGED_RTOL = 2, PFEM_CONVERGED_ATOL = 3, PFEM_CONVERGED_ITS = 4,
    PFEM_DIVERGED_ITS = -3, PFEM_DIVERGED_DTOL = -4, PFEM_DIVERGED_INDEFINITE_PC = -8,
    PFEM_DIVERGED_NANORINF = -9, PFEM_DIVERGED_INDEFINITE_MAT = -10,
    PFEM_DIVERGED_PCSETUP_FAILED = -11 /* zero pivot in ILU(0) */
};

const char *pfem_last_error(void);
int pfem_device_count(int *count);

/* ---------------------------------------------------------------------------------------------
 * Element routines (elementutilitiespoisson.F, elementutilitieselasticity2D/3D.F), computed on the
 * GPU.  The *_batch forms take n elements in SoA layout (x[i*n+e], K[(i+nsize*j)*n+e], F[i*n+e],
 * valC[k*n+e] or NULL for zero); jac_neg[e] (may be NULL) is set to 1 where Jac < 0.  The
 * single-element forms are a batch of one and return PFEM_ERR_NEG_JACOBIAN where the reference STOPs.
 * ------------------------------------------------------------------------------------------- */
/* StiffnessResidualPoissonLinearTria, elementutilitiespoisson.F:23-101 */
int pfem_poisson_tria_ke(const double x[3], const double y[3], const double *elemData, const double *timeData,
                         const double valC[3], const double valDotC[3], double K[9], double F[3]);
/* StiffnessResidualPoissonLinearTetra, elementutilitiespoisson.F:107-193 */
int pfem_poisson_tetra_ke(const double x[4], const double y[4], const double z[4], const double *elemData,
                          const double *timeData, const double valC[4], const double valDotC[4], double K[16],
                          double F[4]);
/* StiffnessResidualElasticityLinearTria, elementutilitieselasticity2D.F:23-153 */
int pfem_elasticity_tria_ke(const double x[3], const double y[3], const double *elemData, const double *timeData,
                            const double valC[6], const double valDotC[6], double K[36], double F[6]);
/* StiffnessResidualElasticityLinearTetra, elementutilitieselasticity3D.F:248-393 (ETYPE=4, 1 Gauss point) */
int pfem_elasticity_tetra_ke(const double x[4], const double y[4], const double z[4], const double *elemData,
                             const double *timeData, const double valC[12], const double valDotC[12],
                             double K[144], double F[12]);
/* kind = PFEM_POISSON_TRIA ... PFEM_ELASTICITY_TETRA; z is ignored for the 2-D kinds */
int pfem_element_ke_batch(int kind, int n, const double *x, const double *y, const double *z,
                          const double *elemData, const double *timeData, const double *valC, double *K,
                          double *F, int *jac_neg);

/* ---------------------------------------------------------------------------------------------
 * Communicator.  Rank 0 calls pfem_comm_unique_id and the host program broadcasts the 128 bytes
 * (the Fortran drivers would MPI_Bcast it next to elem_proc_id, tetrapoissonparallelimpl1.F:480).
 * ------------------------------------------------------------------------------------------- */
int pfem_comm_unique_id(void *id128);

/* TYPE PetscSolver, solverpetsc.F:72-105.  nccl_id may be NULL when nranks == 1. */
int pfem_solver_create(pfem_solver_t **h, int device, int rank, int nranks, const void *nccl_id128);
/* PetscSolver%free, solverpetsc.F:254-278 */
int pfem_solver_free(pfem_solver_t *h);

/* PetscSolver%initialise, solverpetsc.F:116-214.  size_local = rows owned by this rank (contiguous
 * block, in rank order, as PETSc lays them out); the nnz hints are accepted and ignored (the pattern
 * is built exactly).  Resets the state to PFEM_SOLVER_EMPTY.  CG + PETSc default tolerances. */
int pfem_solver_initialise(pfem_solver_t *h, int size_local, int size_global, const int *diag_nnz,
                           const int *offdiag_nnz);
/* KSPSetFromOptions / PCSetFromOptions, solverpetsc.F:190-210 (-ksp_rtol -ksp_atol -ksp_divtol -ksp_max_it -pc_type) */
int pfem_solver_set_options(pfem_solver_t *h, double rtol, double abstol, double dtol, int max_it, int pc_type);
/* PetscInitialize(PETSC_NULL_CHARACTER "petsc_options.dat") + KSPSetFromOptions/PCSetFromOptions
   (tetrapoissonparallelimpl1.F:168, solverpetsc.F:190-210): -ksp_rtol -ksp_atol -ksp_divtol -ksp_max_it -ksp_type cg
   -pc_type none|jacobi|bjacobi [-sub_pc_type ilu]; a missing file leaves the coded defaults (CG + PCBJACOBI/ILU(0),
   rtol 1e-5), an unsupported KSP/PC type is PFEM_ERR_ARG.  In pfem_solver_set_options negative values mean "keep". */
int pfem_solver_set_options_from_file(pfem_solver_t *h, const char *path);

/* The mesh arrays the element loop reads (tetrapoissonparallelimpl1.F:832-838): uploaded once and kept
 * resident in HBM.  conn holds NEW node ids; coords stay in OLD numbering and are reached through
 * node_map_get_old (NULL = identity).  In a multi-rank run a rank may pass only the elements that touch
 * its rows (owned + overlap elements, ascending global id); elements that touch none are ignored. */
int pfem_solver_set_mesh(pfem_solver_t *h, int kind, int nElem, const int *conn, int nNode, const double *coords,
                         const int *node_map_get_old);
/* Pattern pass + setZero: the MatSetValues(INSERT_VALUES, zeros) loop at tetrapoissonparallelimpl1.F:791-802
 * followed by MatAssembly (solverpetsc.F:228-231).  Builds, on the GPU, the CSR pattern of the owned rows
 * (sorted unique global columns, explicit zeros kept, negative dofs dropped) and the row -> element
 * incidence lists the value pass gathers from.  State -> PFEM_PATTERN_OK. */
int pfem_solver_set_pattern(pfem_solver_t *h, int nElem, int nsize, const int *elemDof);
/* Same pattern pass, but the element dof lists are formed on the GPU from the nodal numbering:
 * ElemDofArray(e, ndof*(i-1)+j) = NodeDofArrayNew(conn(e,i), j) - 1 (tetrapoissonparallelimpl1.F:698-713).
 * NodeDofArrayNew is the driver's array (column-major nNode x ndof, 1-based dof id, 0 = Dirichlet); the host then
 * uploads nNode*ndof ints instead of nElem*nsize. */
int pfem_solver_set_pattern_nodal(pfem_solver_t *h, int ndof, const int *NodeDofArrayNew);
/* PetscSolver%setZero, solverpetsc.F:222-246 */
int pfem_solver_set_zero(pfem_solver_t *h);
/* solnApplied (tetrapoissonparallelimpl1.F:352,676): applied Dirichlet values indexed (node-1)*ndof+dof, NEW ids */
int pfem_solver_set_applied(pfem_solver_t *h, const double *solnApplied, int n);

/* The whole value pass (tetrapoissonparallelimpl1.F:828-884; tetraelasticityparallelimpl1.F:901-968):
 * Ke/Fe + MatSetValues(ADD) + Dirichlet lifting + VecSetValues(ADD) for every element, in one fused
 * kernel.  Contributions are summed per matrix entry in ascending element order (the reference's np=1
 * order), without atomics.  elemData/timeData as in the element routines.  n_negative_jac (may be NULL)
 * receives the number of negative-Jacobian elements seen; if it is non-zero the call returns
 * PFEM_ERR_NEG_JACOBIAN.  State -> PFEM_ASSEMBLY_OK. */
int pfem_solver_assemble(pfem_solver_t *h, const double *elemData, const double *timeData, int *n_negative_jac);

/* MatSetValues(mtx, n, rows, n, cols, Klocal, ADD_VALUES) as the drivers call it
 * (tetrapoissonparallelimpl1.F:851): Klocal is column-major and PETSc reads it row-major, so entry
 * (rows[i], cols[j]) receives Klocal[j + n*i]; negative indices and rows of other ranks are skipped. */
int pfem_solver_add_matrix(pfem_solver_t *h, int n, const int *rows, const int *cols, const double *Klocal);
/* VecSetValues(rhsVec, n, rows, F, ADD_VALUES), tetrapoissonparallelimpl1.F:880 */
int pfem_solver_add_vector(pfem_solver_t *h, int n, const int *rows, const double *F);
/* VecSetValue(rhsVec, row, val, ADD_VALUES): the ForceBC add, tetraelasticityparallelimpl1.F:971-982 */
int pfem_solver_add_value(pfem_solver_t *h, int row, double val);
/* PetscSolver%assembleMatrix / assembleVector / assembleMatrixAndVector, solverpetsc.F:328-401:
 * entry (R(ii), C(jj)) += KLOCAL(ii,jj) (per-entry MatSetValue: NOT transposed) */
int pfem_solver_assemble_matrix(pfem_solver_t *h, int n, const int *rindices, const int *cindices,
                                const double *KLOCAL);
int pfem_solver_assemble_vector(pfem_solver_t *h, int n, const int *rindices, const double *FLOCAL);
int pfem_solver_assemble_matrix_and_vector(pfem_solver_t *h, int n, const int *rindices, const int *cindices,
                                           const double *KLOCAL, const double *FLOCAL);

/* PetscSolver%factorise / solve / factoriseAndSolve, solverpetsc.F:409-509.  Same state machine
 * (PFEM_ERR_STATE where the reference STOPs).  solve zeroes the initial guess (solverpetsc.F:459) and runs
 * KSPSolve_CG semantics: left preconditioning, preconditioned-norm test, PETSc reason codes. */
int pfem_solver_factorise(pfem_solver_t *h);
int pfem_solver_solve(pfem_solver_t *h);
int pfem_solver_factorise_and_solve(pfem_solver_t *h);

/* VecScatterCreateToAll + VecGetArray, tetrapoissonparallelimpl1.F:922-933: the full solution on every rank */
int pfem_solver_get_solution(pfem_solver_t *h, double *x_global);
int pfem_solver_get_solution_local(pfem_solver_t *h, double *x_local);
int pfem_solver_get_rhs(pfem_solver_t *h, double *rhs_local);
int pfem_solver_set_rhs(pfem_solver_t *h, const double *rhs_local);
/* owned rows of the assembled matrix: rowptr[size_local+1], col[nnz] (global), val[nnz]; any may be NULL */
int pfem_solver_get_nnz(pfem_solver_t *h, long long *nnz);
int pfem_solver_get_csr(pfem_solver_t *h, int *rowptr, int *col, double *val);
/* diagnostics: ILU(0) factor of the last PCBJACOBI solve (PCSetUp_ILU inside KSPSolve, solverpetsc.F:476) on the CSR slots
   of the local rows (unit-lower multipliers below the diagonal, U on and above it, slots outside the diagonal block keep
   the matrix value) and the inverted pivots [size_local] */
int pfem_solver_get_ilu_factor(pfem_solver_t *h, double *fval, double *invdiag);
/* KSPGetIterationNumber / KSPGetConvergedReason (solverpetsc.F:479-488) and the two timers the drivers
 * print (tetrapoissonparallelimpl1.F:893-905), measured with CUDA events on the library's stream */
int pfem_solver_get_info(pfem_solver_t *h, int *its, int *reason, double *rnorm, double *t_assemble_s,
                         double *t_solve_s);
int pfem_solver_get_state(pfem_solver_t *h, int *state, int *row_start, int *row_end, int *size_global);
/* exchange path in use for nranks > 1: 0 = single rank, 1 = NCCL send/recv + all-reduce, 2 = peer-memory kernels (NVLink, CUDA IPC) */
int pfem_solver_comm_mode(pfem_solver_t *h, int *mode);
/* kernel used by the last value pass: 0 = generic row gather (binary-search slots, reference arithmetic), 1 = streamed row
 * gather with the reference-order no-FMA arithmetic (default), 2 = round-1 tiled compute-once kernel (opt-in,
 * PFEM_ASM=tiled|tiled2), 3 = colour-scheduled tile kernel (opt-in, PFEM_ASM=ctile), 4 = streamed row gather with the FMA
 * element operators (opt-in, PFEM_ASM=fast); for mode 2 also the number of row tiles and the mean number of times an element is computed */
int pfem_solver_assembly_mode(pfem_solver_t *h, int *mode, int *ntiles, double *visits_per_element);
/* Which arithmetic the value pass uses (no reference counterpart: PETSc's MatSetValues has one code path).  Both sum every
 * matrix entry in the reference's sequential element order (tetrapoissonparallelimpl1.F:828-884), deterministically.
 * PFEM_ASM_ROWS (= PFEM_ASM_AUTO, the default): the reference's no-FMA evaluation order of Ke/Fe -- bit-identical to a
 * sequential np=1 CPU run.
 * PFEM_ASM_FAST: FMA-contracted element operators (cofactor form for the Poisson kinds) -- values agree with the reference
 * evaluation order to rounding (<= 1e-12 relative, the north-star contract), not bit for bit; 5 % faster on C5. */
int pfem_solver_set_assembly_mode(pfem_solver_t *h, int mode);
/* name of the kernel of the last value pass, its FP64 instructions per element visit, visits per pass, arithmetic (0 no-FMA, 1 FMA) */
int pfem_solver_assembly_info(pfem_solver_t *h, char *kernel, int cap, double *fp64_per_visit, long long *visits, int *arith);
/* kernel launches issued on this handle since the last call with reset != 0 */
int pfem_solver_launch_count(pfem_solver_t *h, long long *launches, int reset);
/* Run the CG SpMV (w = A p on an internal vector) `reps` times and report the mean device time per
 * launch in seconds: the roofline probe used by bench.py. */
int pfem_solver_time_spmv(pfem_solver_t *h, int reps, double *seconds_per_launch);
/* Optional: bracket every SpMV launch of the next solves with CUDA events (on the launching stream) and
 * report their summed device time and count: the live per-launch duration bench.py's roofline uses. */
int pfem_solver_set_profiling(pfem_solver_t *h, int on);
int pfem_solver_get_profile(pfem_solver_t *h, double *spmv_seconds_total, long long *spmv_launches);
/* PetscSolver%printInfo, solverpetsc.F:286-320 */
int pfem_solver_print_info(pfem_solver_t *h);

/* ---------------------------------------------------------------------------------------------------------------------
 * Explicit dynamics: the matrix-free element residual / lumped-mass routines and the
 * central-difference time loop of the *elasticityexplicit drivers.  No matrix, no solver object.
 * Single-element entry points mirror the Fortran routines argument for argument (status PFEM_ERR_NEG_JACOBIAN where
 * they STOP):
 *   ResidualElasticityLinearTria   elementutilitieselasticity2D.F:158-275  (plane strain; Flocal[6])
 *   MassMatrixLinearTria           elementutilitieselasticity2D.F:283-362  (Mlocal[6], row sums = lumped mass)
 *   ResidualElasticityLinearTetra  elementutilitieselasticity3D.F:575-723  (Flocal[12]; ETYPE 4 / one Gauss point intent)
 *   MassMatrixLinearTetra          elementutilitieselasticity3D.F:401-482  (Mlocal[12])
 * elemData = (E, nu, density, bx, by[, bz]); timeData and the velocity argument are accepted and, like in the reference,
 * unused. */
int pfem_residual_elasticity_linear_tria(const double *x, const double *y, const double *elemData, const double *timeData,
                                         const double *dispC, const double *veloC, double *Flocal);
int pfem_mass_matrix_linear_tria(const double *x, const double *y, const double *elemData, double *Mlocal);
int pfem_residual_elasticity_linear_tetra(const double *x, const double *y, const double *z, const double *elemData,
                                          const double *timeData, const double *valC, const double *valDotC, double *Flocal);
int pfem_mass_matrix_linear_tetra(const double *x, const double *y, const double *z, const double *elemData, double *Mlocal);

/* Batched time loop (one call replaces the element loops of triaelasticityexplicit.F:881-921 and :972-1121).
 * Arrays as in the driver: conn SoA [npElem][nElem] 1-based, coords SoA [ndim][nNode]; displacement / velocity /
 * acceleration / mass are indexed by node slot (node-1)*ndof + dof like the driver's plain arrays; assyForSoln lists the
 * 1-based slots of the free dofs (:722-734), all other dofs keep their value (the reference never touches them).
 * One process, one GPU (the explicit drivers loop over ALL elements on every rank). */
typedef struct pfem_explicit pfem_explicit_t;
int pfem_explicit_create(pfem_explicit_t **ex, int device);
int pfem_explicit_free(pfem_explicit_t *ex);
int pfem_explicit_set_mesh(pfem_explicit_t *ex, int kind /* PFEM_ELASTICITY_TRIA | PFEM_ELASTICITY_TETRA */, int nElem, const int *conn,
                           int nNode, const double *coords);
int pfem_explicit_set_free_dofs(pfem_explicit_t *ex, int size_global, const int *assyForSoln);
/* globalM: triaelasticityexplicit.F:881-921 */
int pfem_explicit_lumped_mass(pfem_explicit_t *ex, const double *elemData);
/* nsteps passes of the time loop body (:972-1121) with constant elemData (the driver's load switch at timeNow <= 0.1 is
   host logic: call once per load phase) */
int pfem_explicit_advance(pfem_explicit_t *ex, int nsteps, double dt, const double *elemData, const double *timeData);
/* any output may be NULL; disp = current displacement (= dispPrev of the next step), dispPrev2 = the one before */
int pfem_explicit_get_state(pfem_explicit_t *ex, double *disp, double *dispPrev2, double *velo, double *acce, double *mass);
/* restart from a saved (disp, dispPrev2) pair: resume is bit-identical to an uninterrupted run */
int pfem_explicit_set_state(pfem_explicit_t *ex, const double *disp, const double *dispPrev2);
int pfem_explicit_get_info(pfem_explicit_t *ex, long long *steps, long long *launches, double *t_advance);

/* ---------------------------------------------------------------------------------------------------------------------
 * Driver set-up loops on the GPU.  Host arrays in, host arrays out, the work in between
 * as sorts / scans / gathers on `device`; outputs are bit-identical to the sequential loops of the drivers.
 * Negative return = -status.
 *   pfem_gpu_number_dofs    tetrapoissonparallelimpl1.F:357-367 (free dofs), :402-421 (np = 1), :500-677 (node renumbering by
 *                           partition, NodeDofArrayNew, row ranges, applied values re-keyed).  All node ids 1-based;
 *                           NodeDofArrayNew column-major nNode x ndof (0 = Dirichlet); part_info [nparts][5] = node_start,
 *                           node_end, row_start, row_end (1-based, inclusive), size_local.  Returns size_global.
 *   pfem_gpu_renumber_conn  :659-664, in place.
 *   pfem_gpu_elem_dof_array :698-713 ElemDofArray (SoA [nsize][nElem], 0-based, -1 = Dirichlet), :722-734 assyForSoln
 *                           (NULL to skip) and, with list != NULL, the owned + overlap elements of the row block
 *                           [row_lo, row_hi) in ascending id; returns their count.
 *   pfem_gpu_gen_tetra      genTetra.cpp:194-334 (nodes, 6 tets per cell) + :497-525 (Dirichlet rows) for the box mesh;
 *                           ax/ay/az are the accumulated axis coordinates; coords == NULL returns the number of Dirichlet rows. */
int pfem_gpu_number_dofs(int device, int nNode, int ndof, int nDBC, const int *dbc_node, const int *dbc_dof, const double *dbc_val,
                         int nparts, const int *node_proc_id, int *node_map_get_old, int *node_map_get_new, int *NodeDofArrayNew,
                         double *solnApplied, int *part_info);
int pfem_gpu_renumber_conn(int device, long long n_entries, int *conn, int nNode, const int *node_map_get_new);
int pfem_gpu_elem_dof_array(int device, int nElem, int npElem, int ndof, int nNode, const int *conn_new, const int *NodeDofArrayNew,
                            int size_global, int *elemDof, int *assyForSoln, int row_lo, int row_hi, int *list);
long long pfem_gpu_gen_tetra(int device, int nEx, int nEy, int nEz, const double *ax, const double *ay, const double *az, int dbc_mode,
                             int ndof, double *coords, int *conn, int *dbc_node, int *dbc_dof, double *dbc_val);

#ifdef __cplusplus
}
#endif
#endif /* PFEM_B200_H */
