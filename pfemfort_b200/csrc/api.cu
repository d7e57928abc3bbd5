// api.cu -- the extern "C" boundary of libpfemb200.so (see include/pfem_b200.h for the reference
// interface each entry point replaces).  Host code only: argument checks, the PetscSolver state machine
// (solverpetsc.F:64-68, 409-445, 498-509) and dispatch to the CUDA translation units.
#include <cstdarg>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <new>

#include "internal.cuh"

namespace pfem {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

double StageTimer::now()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
StageTimer::StageTimer(const char *n) : name(n), t0(0), on(false)
{
    const char *e = getenv("PFEM_TRACE");
    on = e && e[0] == '1';
    if (on) { cudaDeviceSynchronize(); t0 = now(); }
}
StageTimer::~StageTimer()
{
    if (on) { cudaDeviceSynchronize(); fprintf(stderr, "[pfem trace] %-28s %8.3f ms\n", name, 1e3 * (now() - t0)); }
}

static int need_handle(pfem_solver *h, const char *who)
{
    if (!h) { set_error("%s: NULL solver handle", who); return PFEM_ERR_ARG; }
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) { set_error("%s: cudaSetDevice(%d): %s", who, h->device, cudaGetErrorString(e)); return PFEM_ERR_CUDA; }
    return PFEM_OK;
}

static int require_gpu(const char *who)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("%s: no CUDA device (%s); libpfemb200 has no CPU fallback", who, e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
        return PFEM_ERR_CUDA;
    }
    return PFEM_OK;
}

}  // namespace pfem

using namespace pfem;

extern "C" {

const char *pfem_last_error(void) { return g_err; }

int pfem_device_count(int *count)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (count) *count = e == cudaSuccess ? n : 0;
    if (e != cudaSuccess) { set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); return PFEM_ERR_CUDA; }
    return PFEM_OK;
}

// ---- element routines -------------------------------------------------------------------------------------------

int pfem_element_ke_batch(int kind, int n, const double *x, const double *y, const double *z, const double *elemData,
                          const double *timeData, const double *valC, double *K, double *F, int *jac_neg)
{
    PFEM_TRY(require_gpu("pfem_element_ke_batch"));
    return element_ke_batch(kind, n, x, y, z, elemData, timeData, valC, K, F, jac_neg);
}

static int single_ke(int kind, const double *x, const double *y, const double *z, const double *elemData,
                     const double *timeData, const double *valC, double *K, double *F)
{
    int neg = 0;
    PFEM_TRY(pfem_element_ke_batch(kind, 1, x, y, z, elemData, timeData, valC, K, F, &neg));
    if (neg) { set_error("negative Jacobian (the reference STOPs)"); return PFEM_ERR_NEG_JACOBIAN; }
    return PFEM_OK;
}

int pfem_poisson_tria_ke(const double x[3], const double y[3], const double *elemData, const double *timeData,
                         const double valC[3], const double valDotC[3], double K[9], double F[3])
{
    (void)valDotC;
    return single_ke(PFEM_POISSON_TRIA, x, y, nullptr, elemData, timeData, valC, K, F);
}

int pfem_poisson_tetra_ke(const double x[4], const double y[4], const double z[4], const double *elemData,
                          const double *timeData, const double valC[4], const double valDotC[4], double K[16], double F[4])
{
    (void)valDotC;
    return single_ke(PFEM_POISSON_TETRA, x, y, z, elemData, timeData, valC, K, F);
}

int pfem_elasticity_tria_ke(const double x[3], const double y[3], const double *elemData, const double *timeData,
                            const double valC[6], const double valDotC[6], double K[36], double F[6])
{
    (void)valDotC;
    return single_ke(PFEM_ELASTICITY_TRIA, x, y, nullptr, elemData, timeData, valC, K, F);
}

int pfem_elasticity_tetra_ke(const double x[4], const double y[4], const double z[4], const double *elemData,
                             const double *timeData, const double valC[12], const double valDotC[12], double K[144],
                             double F[12])
{
    (void)valDotC;
    return single_ke(PFEM_ELASTICITY_TETRA, x, y, z, elemData, timeData, valC, K, F);
}

// ---- solver object ------------------------------------------------------------------------------------------------

int pfem_comm_unique_id(void *id128)
{
    if (!id128) { set_error("pfem_comm_unique_id: NULL"); return PFEM_ERR_ARG; }
    return comm_unique_id(id128);
}

int pfem_solver_create(pfem_solver_t **out, int device, int rank, int nranks, const void *nccl_id128)
{
    if (!out || nranks < 1 || rank < 0 || rank >= nranks) { set_error("pfem_solver_create: bad argument"); return PFEM_ERR_ARG; }
    *out = nullptr;
    PFEM_TRY(require_gpu("pfem_solver_create"));
    PFEM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PFEM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("pfem_solver_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return PFEM_ERR_CUDA;
    }
    pfem_solver *h = new (std::nothrow) pfem_solver;
    if (!h) { set_error("out of host memory"); return PFEM_ERR_ARG; }
    h->device = device; h->rank = rank; h->nranks = nranks;
    h->sm_count = prop.multiProcessorCount;
    // any failure from here on releases the half-built handle (streams / events already created) before returning
    auto make = [&]() -> int {
        PFEM_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        PFEM_CUDA(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
        PFEM_CUDA(cudaEventCreate(&h->ev0));
        PFEM_CUDA(cudaEventCreate(&h->ev1));
        PFEM_CUDA(cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
        PFEM_CUDA(cudaEventCreateWithFlags(&h->ev_pack, cudaEventDisableTiming));
        return comm_init(h, nccl_id128);
    };
    const int st = make();
    if (st != PFEM_OK) { pfem_solver_free(h); return st; }
    *out = h;
    return PFEM_OK;
}

int pfem_solver_free(pfem_solver_t *h)
{
    if (!h) return PFEM_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
    comm_p2p_teardown(h, true);
    comm_destroy(h);
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    if (h->cg_host) cudaFreeHost(h->cg_host);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_halo) cudaEventDestroy(h->ev_halo);
    if (h->ev_pack) cudaEventDestroy(h->ev_pack);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    delete h;
    return PFEM_OK;
}

int pfem_solver_initialise(pfem_solver_t *h, int size_local, int size_global, const int *diag_nnz, const int *offdiag_nnz)
{
    (void)diag_nnz; (void)offdiag_nnz;   // preallocation hints: the pattern pass builds the exact structure
    PFEM_TRY(need_handle(h, "pfem_solver_initialise"));
    if (size_local < 0 || size_global < 0 || size_local > size_global) { set_error("pfem_solver_initialise: bad sizes"); return PFEM_ERR_ARG; }
    std::vector<int> sizes;
    PFEM_TRY(comm_allgather_int(h, size_local, sizes));
    h->row_starts.assign(h->nranks + 1, 0);
    for (int q = 0; q < h->nranks; q++) h->row_starts[q + 1] = h->row_starts[q] + sizes[q];
    if (h->row_starts[h->nranks] != size_global) {      // tetrapoissonparallelimpl1.F:650-655
        set_error("Sum of local problem sizes (%d) is not equal to global size (%d)", h->row_starts[h->nranks], size_global);
        return PFEM_ERR_NUMBERING;
    }
    h->size_local = size_local; h->size_global = size_global;
    h->row_lo = h->row_starts[h->rank]; h->row_hi = h->row_starts[h->rank + 1];
    h->rtol = 1e-5; h->abstol = 1e-50; h->dtol = 1e4; h->max_it = 10000;
    h->pc_type = PFEM_PC_BJACOBI_ILU0;  // solverpetsc.F:206 PCSetType(PCBJACOBI): the reference's default unless the options say otherwise
    h->state = PFEM_SOLVER_EMPTY;       // solverpetsc.F:212
    h->initialised = true;
    h->have_dofs = false;
    h->its = 0; h->reason = 0; h->rnorm = 0.0;
    return PFEM_OK;
}

int pfem_solver_set_options(pfem_solver_t *h, double rtol, double abstol, double dtol, int max_it, int pc_type)
{
    PFEM_TRY(need_handle(h, "pfem_solver_set_options"));
    if (pc_type >= 0 && pc_type != PFEM_PC_NONE && pc_type != PFEM_PC_JACOBI && pc_type != PFEM_PC_BJACOBI_ILU0) {
        set_error("pc_type %d not supported (none|jacobi|bjacobi)", pc_type);
        return PFEM_ERR_ARG;
    }
    if (rtol >= 0) h->rtol = rtol;
    if (abstol >= 0) h->abstol = abstol;
    if (dtol >= 0) h->dtol = dtol;
    if (max_it >= 0) h->max_it = max_it;
    if (pc_type >= 0) h->pc_type = pc_type;           // negative: keep (the default is the reference's PCBJACOBI/ILU(0))
    return PFEM_OK;
}

// PetscInitialize(..., "petsc_options.dat") + KSPSetFromOptions/PCSetFromOptions (tetrapoissonparallelimpl1.F:168,
// solverpetsc.F:190-210): the subset of the PETSc options database this path understands.  Unknown options are ignored
// like PETSc ignores unused ones; a KSP/PC type this library does not implement is an error, never a silent substitute.
int pfem_solver_set_options_from_file(pfem_solver_t *h, const char *path)
{
    PFEM_TRY(need_handle(h, "pfem_solver_set_options_from_file"));
    FILE *f = path ? fopen(path, "r") : nullptr;
    if (!f) return PFEM_OK;                            // no options file: PETSc runs with the coded defaults
    char line[512];
    int rc = PFEM_OK;
    while (rc == PFEM_OK && fgets(line, sizeof line, f)) {
        char key[128] = "", val[256] = "";
        char *hash = strchr(line, '#');
        if (hash) *hash = 0;
        if (sscanf(line, " %127s %255s", key, val) < 1 || key[0] != '-') continue;
        if (!strcmp(key, "-ksp_rtol")) h->rtol = atof(val);
        else if (!strcmp(key, "-ksp_atol")) h->abstol = atof(val);
        else if (!strcmp(key, "-ksp_divtol")) h->dtol = atof(val);
        else if (!strcmp(key, "-ksp_max_it")) h->max_it = atoi(val);
        else if (!strcmp(key, "-ksp_type")) {
            if (strcmp(val, "cg")) { set_error("%s: -ksp_type %s not supported (cg)", path, val); rc = PFEM_ERR_ARG; }
        } else if (!strcmp(key, "-pc_type")) {
            if (!strcmp(val, "jacobi")) h->pc_type = PFEM_PC_JACOBI;
            else if (!strcmp(val, "none")) h->pc_type = PFEM_PC_NONE;
            else if (!strcmp(val, "bjacobi")) h->pc_type = PFEM_PC_BJACOBI_ILU0;
            else { set_error("%s: -pc_type %s not supported (none|jacobi|bjacobi)", path, val); rc = PFEM_ERR_ARG; }
        } else if (!strcmp(key, "-sub_pc_type")) {
            if (strcmp(val, "ilu")) { set_error("%s: -sub_pc_type %s not supported (ilu)", path, val); rc = PFEM_ERR_ARG; }
        }
    }
    fclose(f);
    return rc;
}

int pfem_solver_set_mesh(pfem_solver_t *h, int kind, int nElem, const int *conn, int nNode, const double *coords,
                         const int *node_map_get_old)
{
    PFEM_TRY(need_handle(h, "pfem_solver_set_mesh"));
    h->state = PFEM_SOLVER_EMPTY;        // a new mesh invalidates the pattern: the pattern pass must run again
    return upload_mesh(h, kind, nElem, conn, nNode, coords, node_map_get_old);
}

int pfem_solver_set_pattern_nodal(pfem_solver_t *h, int ndof, const int *NodeDofArrayNew)
{
    PFEM_TRY(need_handle(h, "pfem_solver_set_pattern_nodal"));
    if (!h->have_mesh) { set_error("pfem_solver_set_pattern_nodal: call pfem_solver_set_mesh first"); return PFEM_ERR_STATE; }
    if (ndof != h->ndof || !NodeDofArrayNew) { set_error("pfem_solver_set_pattern_nodal: bad argument"); return PFEM_ERR_ARG; }
    PFEM_TRY(build_pattern(h, h->nElem, h->nsize, nullptr, NodeDofArrayNew));
    {
        StageTimer tm("pattern: solver structures");
        PFEM_TRY(build_solver_structures(h));
    }
    h->state = PFEM_PATTERN_OK;
    return PFEM_OK;
}

int pfem_solver_set_pattern(pfem_solver_t *h, int nElem, int nsize, const int *elemDof)
{
    PFEM_TRY(need_handle(h, "pfem_solver_set_pattern"));
    PFEM_TRY(build_pattern(h, nElem, nsize, elemDof, nullptr));
    {
        StageTimer tm("pattern: solver structures");
        PFEM_TRY(build_solver_structures(h));
    }
    h->state = PFEM_PATTERN_OK;
    return PFEM_OK;
}

static int need_pattern(pfem_solver *h, const char *who)
{
    PFEM_TRY(need_handle(h, who));
    if (h->state < PFEM_PATTERN_OK) { set_error("%s: set the matrix pattern first", who); return PFEM_ERR_STATE; }
    return PFEM_OK;
}

int pfem_solver_set_zero(pfem_solver_t *h)
{
    PFEM_TRY(need_pattern(h, "pfem_solver_set_zero"));
    PFEM_CUDA(cudaMemsetAsync(h->val.p, 0, (size_t)(h->nnz > 0 ? h->nnz : 1) * sizeof(double), h->stream));
    PFEM_CUDA(cudaMemsetAsync(h->rhs.p, 0, (size_t)(h->size_local > 0 ? h->size_local : 1) * sizeof(double), h->stream));
    PFEM_CUDA(cudaStreamSynchronize(h->stream));
    h->values_zero = true; h->rhs_zero = true;
    for (auto &st : h->stash) st.clear();     // setZero assembles and then zeroes (solverpetsc.F:228-239): stashed adds vanish with it
    return PFEM_OK;
}

int pfem_solver_set_applied(pfem_solver_t *h, const double *solnApplied, int n)
{
    PFEM_TRY(need_handle(h, "pfem_solver_set_applied"));
    if (!h->have_mesh) { set_error("pfem_solver_set_applied: set the mesh first"); return PFEM_ERR_STATE; }
    if (!solnApplied || n != h->nNode * h->ndof) { set_error("pfem_solver_set_applied: expected %d values", h->nNode * h->ndof); return PFEM_ERR_ARG; }
    PFEM_CUDA(cudaMemcpyAsync(h->applied.p, solnApplied, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    PFEM_CUDA(cudaStreamSynchronize(h->stream));
    h->ct_ready = false; h->ct_tried = false;      // the tile kernel's node tables carry the applied values
    return PFEM_OK;
}

int pfem_solver_set_assembly_mode(pfem_solver_t *h, int mode)
{
    PFEM_TRY(need_handle(h, "pfem_solver_set_assembly_mode"));
    if (mode != PFEM_ASM_AUTO && mode != PFEM_ASM_ROWS && mode != PFEM_ASM_FAST) {
        set_error("pfem_solver_set_assembly_mode: mode %d (0 auto | 1 rows | 2 fast)", mode);
        return PFEM_ERR_ARG;
    }
    h->asm_mode_req = mode;
    return PFEM_OK;
}

int pfem_solver_assemble(pfem_solver_t *h, const double *elemData, const double *timeData, int *n_negative_jac)
{
    PFEM_TRY(need_pattern(h, "pfem_solver_assemble"));
    if (!elemData || !timeData) { set_error("pfem_solver_assemble: NULL elemData/timeData"); return PFEM_ERR_ARG; }
    if (n_negative_jac) *n_negative_jac = 0;
    PFEM_TRY(assemble_values(h, elemData, timeData, n_negative_jac));
    h->state = PFEM_ASSEMBLY_OK;
    return PFEM_OK;
}

int pfem_solver_add_matrix(pfem_solver_t *h, int n, const int *rows, const int *cols, const double *Klocal)
{
    PFEM_TRY(need_pattern(h, "pfem_solver_add_matrix"));
    if (!Klocal) { set_error("pfem_solver_add_matrix: NULL block"); return PFEM_ERR_ARG; }
    return add_entries(h, n, rows, cols, Klocal, /*transposed=*/true, nullptr);
}

int pfem_solver_add_vector(pfem_solver_t *h, int n, const int *rows, const double *F)
{
    PFEM_TRY(need_pattern(h, "pfem_solver_add_vector"));
    if (!F) { set_error("pfem_solver_add_vector: NULL vector"); return PFEM_ERR_ARG; }
    return add_entries(h, n, rows, nullptr, nullptr, false, F);
}

int pfem_solver_add_value(pfem_solver_t *h, int row, double val)
{
    return pfem_solver_add_vector(h, 1, &row, &val);
}

int pfem_solver_assemble_matrix(pfem_solver_t *h, int n, const int *rindices, const int *cindices, const double *KLOCAL)
{
    PFEM_TRY(need_pattern(h, "pfem_solver_assemble_matrix"));
    if (!KLOCAL) { set_error("pfem_solver_assemble_matrix: NULL block"); return PFEM_ERR_ARG; }
    return add_entries(h, n, rindices, cindices, KLOCAL, /*transposed=*/false, nullptr);
}

int pfem_solver_assemble_vector(pfem_solver_t *h, int n, const int *rindices, const double *FLOCAL)
{
    return pfem_solver_add_vector(h, n, rindices, FLOCAL);
}

int pfem_solver_assemble_matrix_and_vector(pfem_solver_t *h, int n, const int *rindices, const int *cindices,
                                           const double *KLOCAL, const double *FLOCAL)
{
    PFEM_TRY(need_pattern(h, "pfem_solver_assemble_matrix_and_vector"));
    if (!KLOCAL || !FLOCAL) { set_error("pfem_solver_assemble_matrix_and_vector: NULL argument"); return PFEM_ERR_ARG; }
    return add_entries(h, n, rindices, cindices, KLOCAL, /*transposed=*/false, FLOCAL);
}

int pfem_solver_factorise(pfem_solver_t *h)
{
    PFEM_TRY(need_handle(h, "pfem_solver_factorise"));
    if (h->state != PFEM_ASSEMBLY_OK) {                  // solverpetsc.F:415-419
        set_error("Assemble matrix first before solving it!");
        return PFEM_ERR_STATE;
    }
    h->state = PFEM_FACTORISE_OK;
    return PFEM_OK;
}

int pfem_solver_solve(pfem_solver_t *h)
{
    PFEM_TRY(need_handle(h, "pfem_solver_solve"));
    if (h->state != PFEM_FACTORISE_OK) {                 // solverpetsc.F:441-445
        set_error("Factorise matrix first before solving it!");
        return PFEM_ERR_STATE;
    }
    PFEM_TRY(stash_flush(h));                            // MatAssemblyBegin/End, VecAssemblyBegin/End (solverpetsc.F:447-468)
    return cg_solve(h);
}

int pfem_solver_factorise_and_solve(pfem_solver_t *h)
{
    PFEM_TRY(need_pattern(h, "pfem_solver_factorise_and_solve"));
    h->state = PFEM_ASSEMBLY_OK;                         // solverpetsc.F:504 force-sets the state
    PFEM_TRY(pfem_solver_factorise(h));
    return pfem_solver_solve(h);
}

int pfem_solver_get_solution_local(pfem_solver_t *h, double *x_local)
{
    PFEM_TRY(need_pattern(h, "pfem_solver_get_solution_local"));
    if (!x_local) { set_error("NULL output"); return PFEM_ERR_ARG; }
    PFEM_CUDA(cudaMemcpyAsync(x_local, h->x.p, (size_t)h->size_local * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    PFEM_CUDA(cudaStreamSynchronize(h->stream));
    return PFEM_OK;
}

int pfem_solver_get_solution(pfem_solver_t *h, double *x_global)
{
    PFEM_TRY(need_pattern(h, "pfem_solver_get_solution"));
    if (!x_global) { set_error("NULL output"); return PFEM_ERR_ARG; }
    if (h->nranks == 1) return pfem_solver_get_solution_local(h, x_global);
    DevBuf<double> g;
    PFEM_TRY(g.alloc((size_t)h->size_global));
    PFEM_TRY(comm_allgatherv_double(h, h->x.p, g.p, h->stream));
    PFEM_CUDA(cudaMemcpyAsync(x_global, g.p, (size_t)h->size_global * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    PFEM_CUDA(cudaStreamSynchronize(h->stream));
    return PFEM_OK;
}

int pfem_solver_get_rhs(pfem_solver_t *h, double *rhs_local)
{
    PFEM_TRY(need_pattern(h, "pfem_solver_get_rhs"));
    if (!rhs_local) { set_error("NULL output"); return PFEM_ERR_ARG; }
    PFEM_CUDA(cudaMemcpyAsync(rhs_local, h->rhs.p, (size_t)h->size_local * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    PFEM_CUDA(cudaStreamSynchronize(h->stream));
    return PFEM_OK;
}

int pfem_solver_set_rhs(pfem_solver_t *h, const double *rhs_local)
{
    PFEM_TRY(need_pattern(h, "pfem_solver_set_rhs"));
    if (!rhs_local) { set_error("NULL input"); return PFEM_ERR_ARG; }
    PFEM_CUDA(cudaMemcpyAsync(h->rhs.p, rhs_local, (size_t)h->size_local * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    PFEM_CUDA(cudaStreamSynchronize(h->stream));
    h->rhs_zero = false;
    return PFEM_OK;
}

int pfem_solver_get_nnz(pfem_solver_t *h, long long *nnz)
{
    PFEM_TRY(need_pattern(h, "pfem_solver_get_nnz"));
    if (nnz) *nnz = h->nnz;
    return PFEM_OK;
}

// diagnostics: the ILU(0) factor of the last solve with pc_type = PFEM_PC_BJACOBI_ILU0, on the CSR slots of the local rows
// (slots outside the diagonal block keep the matrix value) + the inverted pivots; parity tests compare them bit for bit
int pfem_solver_get_ilu_factor(pfem_solver_t *h, double *fval, double *invdiag)
{
    PFEM_TRY(need_pattern(h, "pfem_solver_get_ilu_factor"));
    if (!fval || !invdiag) { set_error("NULL output"); return PFEM_ERR_ARG; }
    if (!h->ilu_fval.p || h->ilu_fval.n < (size_t)h->nnz || !h->ilu_invd.p) { set_error("pfem_solver_get_ilu_factor: no ILU(0) factor yet"); return PFEM_ERR_STATE; }
    PFEM_CUDA(cudaMemcpyAsync(fval, h->ilu_fval.p, (size_t)h->nnz * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    PFEM_CUDA(cudaMemcpyAsync(invdiag, h->ilu_invd.p, (size_t)h->size_local * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    PFEM_CUDA(cudaStreamSynchronize(h->stream));
    return PFEM_OK;
}

int pfem_solver_get_csr(pfem_solver_t *h, int *rowptr, int *col, double *val)
{
    PFEM_TRY(need_pattern(h, "pfem_solver_get_csr"));
    cudaStream_t s = h->stream;
    if (rowptr) PFEM_CUDA(cudaMemcpyAsync(rowptr, h->rowptr.p, ((size_t)h->size_local + 1) * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (col && h->nnz) PFEM_CUDA(cudaMemcpyAsync(col, h->col.p, (size_t)h->nnz * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (val && h->nnz) PFEM_CUDA(cudaMemcpyAsync(val, h->val.p, (size_t)h->nnz * sizeof(double), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    return PFEM_OK;
}

int pfem_solver_get_info(pfem_solver_t *h, int *its, int *reason, double *rnorm, double *t_assemble_s, double *t_solve_s)
{
    if (!h) { set_error("NULL handle"); return PFEM_ERR_ARG; }
    if (its) *its = h->its;
    if (reason) *reason = h->reason;
    if (rnorm) *rnorm = h->rnorm;
    if (t_assemble_s) *t_assemble_s = h->t_assemble;
    if (t_solve_s) *t_solve_s = h->t_solve;
    return PFEM_OK;
}

int pfem_solver_get_state(pfem_solver_t *h, int *state, int *row_start, int *row_end, int *size_global)
{
    if (!h) { set_error("NULL handle"); return PFEM_ERR_ARG; }
    if (state) *state = h->state;
    if (row_start) *row_start = h->row_lo;
    if (row_end) *row_end = h->row_hi;
    if (size_global) *size_global = h->size_global;
    return PFEM_OK;
}

int pfem_solver_comm_mode(pfem_solver_t *h, int *mode)
{
    if (!h) { set_error("NULL handle"); return PFEM_ERR_ARG; }
    if (mode) *mode = h->nranks == 1 ? 0 : (h->p2p ? 2 : 1);
    return PFEM_OK;
}

int pfem_solver_assembly_mode(pfem_solver_t *h, int *mode, int *ntiles, double *visits_per_element)
{
    if (!h) { set_error("NULL handle"); return PFEM_ERR_ARG; }
    if (mode) *mode = h->last_asm_mode;
    if (ntiles) *ntiles = h->last_asm_mode == 2 ? h->ntiles : 0;
    if (visits_per_element)
        *visits_per_element = (h->last_asm_mode == 2 && h->tile_elems_touched) ? (double)h->tile_elem_visits / (double)h->tile_elems_touched : 0.0;
    return PFEM_OK;
}

int pfem_solver_assembly_info(pfem_solver_t *h, char *kernel, int cap, double *fp64_per_visit, long long *visits, int *arith)
{
    if (!h) { set_error("NULL handle"); return PFEM_ERR_ARG; }
    // FP64 instructions (DADD/DMUL/DFMA + the reciprocal sequence) per element visit, counted in the SASS of this build
    // (tools/sass_count.py -> profiles/r02_sass_value_pass.txt); rows kernels: per (row, element) incidence
    char buf[256];
    double f = 0.0;
    long long v = 0;
    int ar = 0;
    switch (h->last_asm_mode) {
    case 3:
        snprintf(buf, sizeof buf, "assemble_ctile_kernel (colour-scheduled tiles: %d tiles of <= %d rows, %d threads, %.3f visits/element, <= %d rounds)",
                 h->ct_ntiles, h->ct_TR, h->ct_threads, h->nElem ? (double)h->ct_visits / h->nElem : 0.0, h->ct_max_rounds);
        f = h->npe == 4 ? PFEM_FP64_PER_VISIT_CTILE_TET : PFEM_FP64_PER_VISIT_CTILE_TRIA; v = h->ct_visits; ar = 1;
        break;
    case 4:
        snprintf(buf, sizeof buf, "assemble_sell_kernel<FastOp> (streamed row gather, sequential order, FMA cofactor arithmetic)");
        f = h->kind == PFEM_POISSON_TETRA ? PFEM_FP64_PER_INCIDENCE_FAST_TET : (h->kind == PFEM_POISSON_TRIA ? PFEM_FP64_PER_INCIDENCE_FAST_TRIA : 0.0);
        v = h->ninc; ar = 1;
        break;
    case 2:
        snprintf(buf, sizeof buf, "assemble_tiled_kernel (round-1 compute-once tiles: %d tiles)", h->ntiles);
        f = 177.0; v = h->tile_elem_visits; ar = 0;
        break;
    case 1:
        snprintf(buf, sizeof buf, "assemble_sell_kernel (streamed row gather, sequential order)");
        f = h->kind == PFEM_POISSON_TETRA ? 114.0 : 0.0; v = h->ninc; ar = 0;
        break;
    default:
        snprintf(buf, sizeof buf, "assemble_kernel (row gather, binary-search slots)");
        f = 0.0; v = h->ninc; ar = 0;
    }
    if (kernel && cap > 0) { strncpy(kernel, buf, (size_t)cap - 1); kernel[cap - 1] = 0; }
    if (fp64_per_visit) *fp64_per_visit = f;
    if (visits) *visits = v;
    if (arith) *arith = ar;
    return PFEM_OK;
}

int pfem_solver_launch_count(pfem_solver_t *h, long long *launches, int reset)
{
    if (!h) { set_error("NULL handle"); return PFEM_ERR_ARG; }
    if (launches) *launches = h->launches;
    if (reset) h->launches = 0;
    return PFEM_OK;
}

int pfem_solver_time_spmv(pfem_solver_t *h, int reps, double *seconds_per_launch)
{
    PFEM_TRY(need_pattern(h, "pfem_solver_time_spmv"));
    if (!seconds_per_launch) { set_error("NULL output"); return PFEM_ERR_ARG; }
    return time_spmv(h, reps, seconds_per_launch);
}

int pfem_solver_set_profiling(pfem_solver_t *h, int on)
{
    if (!h) { set_error("NULL handle"); return PFEM_ERR_ARG; }
    h->profile = on != 0;
    h->prof_spmv_s = 0.0;
    h->prof_spmv_n = 0;
    return PFEM_OK;
}

int pfem_solver_get_profile(pfem_solver_t *h, double *spmv_seconds_total, long long *spmv_launches)
{
    if (!h) { set_error("NULL handle"); return PFEM_ERR_ARG; }
    if (spmv_seconds_total) *spmv_seconds_total = h->prof_spmv_s;
    if (spmv_launches) *spmv_launches = h->prof_spmv_n;
    return PFEM_OK;
}

int pfem_solver_print_info(pfem_solver_t *h)
{
    if (!h) { set_error("NULL handle"); return PFEM_ERR_ARG; }
    printf(" pfem_b200 solver: rank %d/%d device %d  rows [%d,%d) of %d  nnz(local) %lld  ghosts %d\n", h->rank, h->nranks,
           h->device, h->row_lo, h->row_hi, h->size_global, h->nnz, h->n_ghost);
    printf("   state %d  its %d  reason %d  rnorm %.6e  t_assemble %.6f s  t_solve %.6f s\n", h->state, h->its, h->reason,
           h->rnorm, h->t_assemble, h->t_solve);
    fflush(stdout);
    return PFEM_OK;
}

}  // extern "C"
