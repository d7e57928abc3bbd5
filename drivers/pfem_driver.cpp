// pfem_driver.cpp -- C++ counterpart of the four *parallelimpl1 PROGRAMs of PFEMFort, above the C ABI.
//
// Same command line as the Fortran executables (bin/makefile:4-11), plus the physics selector:
//     pfem_driver <triapoisson|tetrapoisson|triaelasticity|tetraelasticity> nodes.dat elems.dat DirichBC.dat [ForceBC.dat]
//     pfem_driver <physics> mesh.pfemb        (one binary container of the same arrays: csrc/host_meshio.cu)
// and the same flow (tetrapoissonparallelimpl1.F): read the three text files (:216-355), number the DOFs (:357-367),
// partition with METIS and renumber when more than one rank runs (:423-677), build ElemDofArray (:698-713),
// initialise the solver (:759-779), pattern pass (:791-802), setZero (:817), value pass (:828-884, one batched call),
// ForceBC add (elasticity), factoriseAndSolve (:900), gather the solution (:922-933) and write temp.dat (:934-942).
// Ranks: one process per GPU; RANK / WORLD_SIZE / LOCAL_RANK come from the launcher (torchrun --no-python works);
// rank 0 passes the NCCL id to the others through the file named by PFEM_NCCL_ID_FILE (the MPI_Bcast of a real driver).
// Options in place of petsc_options.dat: PFEM_KSP_RTOL (default 1e-5, PETSc's), PFEM_KSP_MAX_IT.
// PFEM_WRITE_PFEMB=<file> saves the parsed input as a binary container.
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/pfem_b200.h"

extern "C" {
int pfem_host_partition_mesh(int nElem, int nNode, int npElem, const int *conn, int nparts, int dual, int ncommon,
                             int *elem_proc_id, int *node_proc_id, long long *objval);
int pfem_host_number_dofs(int nNode, int ndof, int nDBC, const int *dbc_node, const int *dbc_dof, const double *dbc_val,
                          int nparts, const int *node_proc_id, int *node_map_get_old, int *node_map_get_new,
                          int *NodeDofArrayNew, double *solnApplied, int *part_info);
void pfem_host_renumber_conn(long long n_entries, int *conn, const int *node_map_get_new);
void pfem_host_elem_dof_array(int nElem, int npElem, int ndof, int nNode, const int *conn_new, const int *NodeDofArrayNew,
                              int *elemDof);
int pfem_host_select_elements(int nElem, int nsize, const int *elemDof, int row_lo, int row_hi, int *list);
void pfem_host_gather_rows(int nElem, int ncol, const int *in, int nsel, const int *list, int *out);
long long pfem_host_read_table(const char *path, int ncols, double *out, long long nrows_cap);
int pfem_host_mesh_read_binary_header(const char *path, long long sizes[6]);
int pfem_host_mesh_read_binary(const char *path, double *coords, int *conn, int *dbc_node, int *dbc_dof, double *dbc_val,
                               int *fbc_node, int *fbc_dof, double *fbc_val);
int pfem_host_mesh_write_binary(const char *path, int ndim, int npElem, int nNode, int nElem, const double *coords,
                                const int *conn, int nDBC, const int *dbc_node, const int *dbc_dof, const double *dbc_val,
                                int nFBC, const int *fbc_node, const int *fbc_dof, const double *fbc_val);
}

#define CHECK(call)                                                                     \
    do {                                                                                \
        int st_ = (call);                                                               \
        if (st_ != PFEM_OK) {                                                           \
            fprintf(stderr, " libpfemb200 error %d: %s\n", st_, pfem_last_error());     \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

// one text table of the reference's input format, column-major (one pass over the file: pfem_host_read_table)
static bool read_table(const char *path, int ncols, std::vector<double> &tab, long long &nrows)
{
    nrows = pfem_host_read_table(path, ncols, nullptr, 0);
    if (nrows < 0) { fprintf(stderr, "%s\n", pfem_last_error()); return false; }
    tab.assign((size_t)ncols * (nrows > 0 ? nrows : 1), 0.0);
    return pfem_host_read_table(path, ncols, tab.data(), nrows) == nrows;
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s <triapoisson|tetrapoisson|triaelasticity|tetraelasticity> nodes elems DirichBC [ForceBC]\n"
                        "   or: %s <physics> mesh.pfemb      (binary container of the same arrays, tools/mesh_convert.py)\n", argv[0], argv[0]);
        return 1;
    }
    const std::string phys = argv[1];
    int kind, npElem, ndof, ndim;
    if (phys == "triapoisson") { kind = PFEM_POISSON_TRIA; npElem = 3; ndof = 1; ndim = 2; }
    else if (phys == "tetrapoisson") { kind = PFEM_POISSON_TETRA; npElem = 4; ndof = 1; ndim = 3; }
    else if (phys == "triaelasticity") { kind = PFEM_ELASTICITY_TRIA; npElem = 3; ndof = 2; ndim = 2; }
    else if (phys == "tetraelasticity") { kind = PFEM_ELASTICITY_TETRA; npElem = 4; ndof = 3; ndim = 3; }
    else { fprintf(stderr, "unknown physics %s\n", phys.c_str()); return 1; }
    const int rank = getenv("RANK") ? atoi(getenv("RANK")) : 0;
    const int nranks = getenv("WORLD_SIZE") ? atoi(getenv("WORLD_SIZE")) : 1;
    const int device = getenv("LOCAL_RANK") ? atoi(getenv("LOCAL_RANK")) : 0;
    const int nsize = npElem * ndof;

    // ---- read the input: the reference's three (four) text tables, or one .pfemb container holding the same arrays ----
    int nNode = 0, nElem = 0, nDBC = 0;
    std::vector<double> coords, dbc_val, fbc_val;
    std::vector<int> conn, dbc_node, dbc_dof, fbc_node, fbc_dof;
    const size_t l2 = strlen(argv[2]);
    if (l2 > 6 && !strcmp(argv[2] + l2 - 6, ".pfemb")) {
        long long sz[6];
        CHECK(pfem_host_mesh_read_binary_header(argv[2], sz));
        if (sz[0] != ndim || sz[1] != npElem) { fprintf(stderr, "%s holds a %lldD mesh with %lld nodes per element\n", argv[2], sz[0], sz[1]); return 1; }
        nNode = (int)sz[2]; nElem = (int)sz[3]; nDBC = (int)sz[4];
        coords.resize((size_t)ndim * nNode); conn.resize((size_t)npElem * nElem);
        dbc_node.resize(nDBC); dbc_dof.resize(nDBC); dbc_val.resize(nDBC);
        fbc_node.resize(sz[5]); fbc_dof.resize(sz[5]); fbc_val.resize(sz[5]);
        CHECK(pfem_host_mesh_read_binary(argv[2], coords.data(), conn.data(), dbc_node.data(), dbc_dof.data(), dbc_val.data(),
                                         fbc_node.data(), fbc_dof.data(), fbc_val.data()));
    } else {
        if (argc < 5) { fprintf(stderr, "text input needs nodes, elems and DirichBC files\n"); return 1; }
        std::vector<double> t;
        long long n = 0;
        if (!read_table(argv[2], 1 + ndim, t, n)) return 1;
        nNode = (int)n;
        coords.assign(t.begin() + n, t.begin() + n * (1 + ndim));                     // columns 1..ndim (column 0 is the id)
        if (!read_table(argv[3], 1 + npElem, t, n)) return 1;
        nElem = (int)n;
        conn.resize((size_t)npElem * nElem);
        for (size_t q = 0; q < conn.size(); q++) conn[q] = (int)t[(size_t)n + q];
        if (!read_table(argv[4], 3, t, n)) return 1;
        nDBC = (int)n;
        dbc_node.resize(nDBC); dbc_dof.resize(nDBC); dbc_val.resize(nDBC);
        for (int b = 0; b < nDBC; b++) { dbc_node[b] = (int)t[b]; dbc_dof[b] = (int)t[(size_t)n + b]; dbc_val[b] = t[2 * (size_t)n + b]; }
        if (argc >= 6 && read_table(argv[5], 3, t, n))
            for (long long b = 0; b < n; b++) { fbc_node.push_back((int)t[b]); fbc_dof.push_back((int)t[(size_t)n + b]); fbc_val.push_back(t[2 * (size_t)n + b]); }
    }
    // PFEM_WRITE_PFEMB=<file>: save what was just read as a binary container (text -> binary conversion by the driver itself)
    if (rank == 0 && getenv("PFEM_WRITE_PFEMB"))
        CHECK(pfem_host_mesh_write_binary(getenv("PFEM_WRITE_PFEMB"), ndim, npElem, nNode, nElem, coords.data(), conn.data(), nDBC,
                                          dbc_node.data(), dbc_dof.data(), dbc_val.data(), (int)fbc_node.size(), fbc_node.data(),
                                          fbc_dof.data(), fbc_val.data()));
    if (rank == 0) printf(" nElem_global = %d\n nNode_global = %d\n npElem = %d\n ndof = %d\n", nElem, nNode, npElem, ndof);

    // ---- partition (every rank computes the same METIS partition: deterministic, replaces the MPI_Bcast) ----
    std::vector<int> elem_proc_id(nElem, 0), node_proc_id(nNode, 0);
    CHECK(pfem_host_partition_mesh(nElem, nNode, npElem, conn.data(), nranks, npElem == 3 ? 1 : 0, 2, elem_proc_id.data(),
                                   node_proc_id.data(), nullptr));
    // ---- numbering ----
    std::vector<int> map_old(nNode), map_new(nNode), nda((size_t)ndof * nNode), part_info(5 * nranks);
    std::vector<double> solnApplied((size_t)nNode * ndof);
    const int size_global = pfem_host_number_dofs(nNode, ndof, nDBC, dbc_node.data(), dbc_dof.data(), dbc_val.data(), nranks,
                                                  node_proc_id.data(), map_old.data(), map_new.data(), nda.data(),
                                                  solnApplied.data(), part_info.data());
    if (size_global < 0) { fprintf(stderr, "%s\n", pfem_last_error()); return 1; }
    if (rank == 0) printf(" Total DOF = %d\n", size_global);
    pfem_host_renumber_conn((long long)conn.size(), conn.data(), map_new.data());
    std::vector<int> edof((size_t)nsize * nElem);
    pfem_host_elem_dof_array(nElem, npElem, ndof, nNode, conn.data(), nda.data(), edof.data());
    int row_lo = 0;
    for (int p = 0; p < rank; p++) row_lo += part_info[5 * p + 4];
    const int size_local = part_info[5 * rank + 4], row_hi = row_lo + size_local;

    // ---- solver ----
    unsigned char id[128] = {0};
    if (nranks > 1) {
        const char *idfile = getenv("PFEM_NCCL_ID_FILE");
        if (!idfile) { fprintf(stderr, "WORLD_SIZE > 1 needs PFEM_NCCL_ID_FILE\n"); return 1; }
        if (rank == 0) {
            CHECK(pfem_comm_unique_id(id));
            std::string tmp = std::string(idfile) + ".tmp";
            FILE *f = fopen(tmp.c_str(), "wb");
            fwrite(id, 1, 128, f);
            fclose(f);
            rename(tmp.c_str(), idfile);
        } else {
            FILE *f = nullptr;
            for (int tries = 0; tries < 6000 && !(f = fopen(idfile, "rb")); tries++) usleep(10000);
            if (!f || fread(id, 1, 128, f) != 128) { fprintf(stderr, "cannot read %s\n", idfile); return 1; }
            fclose(f);
        }
    }
    pfem_solver_t *solver = nullptr;
    CHECK(pfem_solver_create(&solver, device, rank, nranks, nranks > 1 ? id : nullptr));
    int n1 = 50, n2 = 25;                                          // tetrapoissonparallelimpl1.F:759-773
    if (size_local < 50) { n1 = size_local; n2 = n1; }
    std::vector<int> diag_nnz(size_local > 0 ? size_local : 1, n1), offdiag_nnz(size_local > 0 ? size_local : 1, n2);
    CHECK(pfem_solver_initialise(solver, size_local, size_global, diag_nnz.data(), offdiag_nnz.data()));
    const double rtol = getenv("PFEM_KSP_RTOL") ? atof(getenv("PFEM_KSP_RTOL")) : 1e-5;
    const int max_it = getenv("PFEM_KSP_MAX_IT") ? atoi(getenv("PFEM_KSP_MAX_IT")) : 10000;
    // options: like the Fortran PROGRAMs, from "petsc_options.dat" in the working directory (PetscInitialize, :168); the
    // environment (PFEM_KSP_RTOL, PFEM_KSP_MAX_IT, PFEM_PC_TYPE=none|jacobi|bjacobi) overrides.  With neither, the reference's
    // coded defaults apply: CG + PCBJACOBI/ILU(0), rtol 1e-5 (solverpetsc.F:187,206).
    CHECK(pfem_solver_set_options_from_file(solver, "petsc_options.dat"));
    int pc = -1;
    if (const char *pcs = getenv("PFEM_PC_TYPE"))
        pc = !strcmp(pcs, "none") ? PFEM_PC_NONE : !strcmp(pcs, "jacobi") ? PFEM_PC_JACOBI : PFEM_PC_BJACOBI_ILU0;
    CHECK(pfem_solver_set_options(solver, getenv("PFEM_KSP_RTOL") ? rtol : -1.0, -1.0, -1.0, getenv("PFEM_KSP_MAX_IT") ? max_it : -1, pc));
    // the elements this rank hands to its GPU: owned + overlap (every element with a dof in its row block)
    std::vector<int> lconn = conn, ledof = edof;
    int nLocal = nElem;
    if (nranks > 1) {
        nLocal = pfem_host_select_elements(nElem, nsize, edof.data(), row_lo, row_hi, nullptr);
        std::vector<int> list(nLocal > 0 ? nLocal : 1);
        pfem_host_select_elements(nElem, nsize, edof.data(), row_lo, row_hi, list.data());
        lconn.assign((size_t)npElem * nLocal, 0);
        ledof.assign((size_t)nsize * nLocal, 0);
        pfem_host_gather_rows(nElem, npElem, conn.data(), nLocal, list.data(), lconn.data());
        pfem_host_gather_rows(nElem, nsize, edof.data(), nLocal, list.data(), ledof.data());
    }
    printf(" Preparing matrix pattern \n");
    CHECK(pfem_solver_set_mesh(solver, kind, nLocal, lconn.data(), nNode, coords.data(), nranks > 1 ? map_old.data() : nullptr));
    CHECK(pfem_solver_set_pattern(solver, nLocal, nsize, ledof.data()));
    CHECK(pfem_solver_set_zero(solver));
    // material constants of the drivers: single-precision literals (tetrapoissonparallelimpl1.F:822-824,
    // tetraelasticityparallelimpl1.F:895-899)
    double elemData[8] = {1.0, 1.0, 1.0, 0, 0, 0, 0, 0}, timeData[8] = {0.0, 1.0, 0.0, 0, 0, 0, 0, 0};
    if (ndof > 1) {
        elemData[0] = (double)240.565f; elemData[1] = (double)0.3f; elemData[2] = 1.0;
        elemData[3] = kind == PFEM_ELASTICITY_TETRA ? (double)0.1f : 0.0; elemData[4] = 0.0; elemData[5] = 0.0;
    }
    printf(" Generating element matrices and vectors \n");
    CHECK(pfem_solver_set_applied(solver, solnApplied.data(), nNode * ndof));
    int nneg = 0;
    CHECK(pfem_solver_assemble(solver, elemData, timeData, &nneg));
    for (size_t b = 0; b < fbc_node.size(); b++) {                 // tetraelasticityparallelimpl1.F:971-982
        const int n1n = map_new[fbc_node[b] - 1];
        const int row = (n1n - 1) * ndof + fbc_dof[b] - 1;
        // every admissible row is added once overall (the reference: range test + PETSc stash); here by the rank that owns it
        if (row >= 1 && row < size_global && row >= row_lo && row < row_hi) CHECK(pfem_solver_add_value(solver, row, fbc_val[b]));
    }
    printf(" Solving the matrix system \n");
    CHECK(pfem_solver_factorise_and_solve(solver));
    int its = 0, reason = 0;
    double rnorm = 0, ta = 0, ts = 0;
    CHECK(pfem_solver_get_info(solver, &its, &reason, &rnorm, &ta, &ts));
    if (rank == 0) {
        printf(" That took %.6f seconds (assembly)\n", ta);
        if (reason < 0) printf("Divergence.\n"); else printf(" Convergence in %d iterations.\n", its);
        printf(" That took %.6f seconds (solve)\n", ts);
    }
    std::vector<double> x(size_global > 0 ? size_global : 1);
    CHECK(pfem_solver_get_solution(solver, x.data()));
    if (rank == 0) {                                               // temp.dat as the PROGRAMs write it (:934-942)
        std::vector<int> assy(size_global);
        int count = 0;
        for (int n = 0; n < nNode; n++)
            for (int d = 0; d < ndof; d++)
                if (nda[(size_t)d * nNode + n] != 0) assy[count++] = n * ndof + d + 1;
        FILE *f = fopen("temp.dat", "w");
        for (int ii = 0; ii < size_global; ii++) {
            const int slot = assy[ii] - 1, node_new = slot / ndof, d = slot % ndof;
            const int ind = (map_old[node_new] - 1) * ndof + d + 1;
            if (ndof == 1) fprintf(f, "%d %d %.17g\n", ii + 1, ind, x[ii]);   // write(1,*) ii, ind, fact   (tetrapoissonparallelimpl1.F:940)
            else           fprintf(f, "%.17g\n", x[ii]);                      // write(1,*) fact            (tetraelasticityparallelimpl1.F:1046)
        }
        fclose(f);
        printf(" Program is successful \n");
    }
    CHECK(pfem_solver_free(solver));
    return 0;
}
