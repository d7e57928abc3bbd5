// host_driver.cu -- host-side (CPU) driver logic that stays on the host in the reference too:
// DOF numbering, node renumbering by partition, ElemDofArray, METIS partitioning, and the selection of the
// elements a rank has to hand to its GPU.  These are the pieces of the *parallelimpl1 PROGRAMs that surround
// the hot path (tetrapoissonparallelimpl1.F:357-367, 402-734); they are exported so that the C++ driver
// (drivers/pfem_driver.cpp) and the Python harness exercise the C ABI the way the Fortran drivers would.
// No CUDA in this file; it is built with nvcc only to share the build recipe.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "internal.cuh"

// METIS 5.x from the CUDA toolkit's libmetis_static.a: 64-bit idx_t (probed), prototypes declared by hand
typedef int64_t metis_idx_t;
extern "C" int METIS_SetDefaultOptions(metis_idx_t *options);
extern "C" int METIS_PartMeshNodal(metis_idx_t *ne, metis_idx_t *nn, metis_idx_t *eptr, metis_idx_t *eind,
                                   metis_idx_t *vwgt, metis_idx_t *vsize, metis_idx_t *nparts, float *tpwgts,
                                   metis_idx_t *options, metis_idx_t *objval, metis_idx_t *epart, metis_idx_t *npart);
extern "C" int METIS_PartMeshDual(metis_idx_t *ne, metis_idx_t *nn, metis_idx_t *eptr, metis_idx_t *eind,
                                  metis_idx_t *vwgt, metis_idx_t *vsize, metis_idx_t *ncommon, metis_idx_t *nparts,
                                  float *tpwgts, metis_idx_t *options, metis_idx_t *objval, metis_idx_t *epart,
                                  metis_idx_t *npart);

#define PFEM_EXPORT extern "C" __attribute__((visibility("default")))

// METIS_PartMeshNodal (tetra drivers, tetrapoissonparallelimpl1.F:457-467) or METIS_PartMeshDual with ncommon
// (tria drivers, triapoissonparallelimpl1.F:486-491).  conn: SoA, 1-based.  Outputs 0-based part ids.
PFEM_EXPORT int pfem_host_partition_mesh(int nElem, int nNode, int npElem, const int *conn, int nparts, int dual,
                                         int ncommon, int *elem_proc_id, int *node_proc_id, long long *objval_out)
{
    if (nparts <= 1) {
        std::fill(elem_proc_id, elem_proc_id + nElem, 0);
        std::fill(node_proc_id, node_proc_id + nNode, 0);
        if (objval_out) *objval_out = 0;
        return PFEM_OK;
    }
    std::vector<metis_idx_t> eptr((size_t)nElem + 1), eind((size_t)nElem * npElem), epart(nElem), npart(nNode);
    for (int e = 0; e < nElem; e++) {
        eptr[e] = (metis_idx_t)e * npElem;
        for (int i = 0; i < npElem; i++) eind[(size_t)e * npElem + i] = conn[(size_t)i * nElem + e] - 1;
    }
    eptr[nElem] = (metis_idx_t)nElem * npElem;
    metis_idx_t ne = nElem, nn = nNode, np = nparts, nc = ncommon, objval = 0;
    metis_idx_t options[40];
    METIS_SetDefaultOptions(options);
    int rc;
    if (dual)
        rc = METIS_PartMeshDual(&ne, &nn, eptr.data(), eind.data(), nullptr, nullptr, &nc, &np, nullptr, options, &objval,
                                epart.data(), npart.data());
    else
        rc = METIS_PartMeshNodal(&ne, &nn, eptr.data(), eind.data(), nullptr, nullptr, &np, nullptr, options, &objval,
                                 epart.data(), npart.data());
    if (rc != 1) { pfem::set_error("METIS returned %d", rc); return PFEM_ERR_ARG; }   // METIS_OK == 1
    for (int e = 0; e < nElem; e++) elem_proc_id[e] = (int)epart[e];
    for (int n = 0; n < nNode; n++) node_proc_id[n] = (int)npart[n];
    if (objval_out) *objval_out = (long long)objval;
    return PFEM_OK;
}

// Node renumbering + DOF numbering of the drivers (tetrapoissonparallelimpl1.F:357-367 free-dof count,
// :402-421 np==1 identity, :500-677 partition-contiguous renumbering).  All node ids 1-based.
//   NodeDofArrayNew : column-major nNode x ndof, 1-based dof id or 0 for a Dirichlet dof
//   solnApplied     : (new node - 1) * ndof + dof, stale old-position entries kept as in the reference
//   part_info       : [nparts][5] = node_start, node_end, row_start, row_end (1-based, inclusive), size_local
// Returns size_global (>= 0) or -PFEM_ERR_NUMBERING.
PFEM_EXPORT int pfem_host_number_dofs(int nNode, int ndof, int nDBC, const int *dbc_node, const int *dbc_dof,
                                      const double *dbc_val, int nparts, const int *node_proc_id, int *node_map_get_old,
                                      int *node_map_get_new, int *NodeDofArrayNew, double *solnApplied, int *part_info)
{
    const size_t nn = (size_t)nNode;
    std::vector<unsigned char> type_old(nn * ndof, 0);
    std::fill(solnApplied, solnApplied + nn * ndof, 0.0);
    for (int b = 0; b < nDBC; b++) {
        const size_t n = (size_t)dbc_node[b] - 1, d = (size_t)dbc_dof[b] - 1;
        type_old[n * ndof + d] = 1;
        solnApplied[n * ndof + d] = dbc_val[b];
    }
    const int size_global = (int)std::count(type_old.begin(), type_old.end(), (unsigned char)0);
    // new -> old map: identity, or the concatenation of every part's ascending list of owned old ids
    if (nparts <= 1) {
        std::iota(node_map_get_old, node_map_get_old + nNode, 1);
    } else {
        std::vector<int> count(nparts + 1, 0);
        for (int n = 0; n < nNode; n++) {
            if (node_proc_id[n] < 0 || node_proc_id[n] >= nparts) { pfem::set_error("node_proc_id out of range"); return -PFEM_ERR_NUMBERING; }
            count[node_proc_id[n] + 1]++;
        }
        for (int p = 0; p < nparts; p++) count[p + 1] += count[p];
        std::vector<int> cursor(count.begin(), count.end() - 1);
        for (int n = 0; n < nNode; n++) node_map_get_old[cursor[node_proc_id[n]]++] = n + 1;   // stable counting sort
    }
    for (int n = 0; n < nNode; n++) node_map_get_new[node_map_get_old[n] - 1] = n + 1;
    // dofs in NEW node order, node-major; each part owns a contiguous node range, hence a contiguous row block
    const int np = std::max(nparts, 1);
    std::vector<int> pend(np, nNode);
    if (nparts > 1) {
        std::vector<int> cnt(nparts, 0);
        for (int n = 0; n < nNode; n++) cnt[node_proc_id[n]]++;
        int acc = 0;
        for (int p = 0; p < nparts; p++) { acc += cnt[p]; pend[p] = acc; }
    }
    int next = 0, n = 0;
    for (int p = 0; p < np; p++) {
        int *info = part_info + 5 * p;
        info[0] = n + 1; info[1] = pend[p];
        info[2] = 1000000000; info[3] = -1000000000; info[4] = 0;     // row_start = 1e9, row_end = -1e9 (:622-623)
        for (; n < pend[p]; n++) {
            const size_t o = (size_t)node_map_get_old[n] - 1;
            for (int d = 0; d < ndof; d++) {
                if (type_old[o * ndof + d]) { NodeDofArrayNew[(size_t)d * nn + n] = 0; continue; }
                const int id = ++next;
                NodeDofArrayNew[(size_t)d * nn + n] = id;
                info[2] = std::min(info[2], id);
                info[3] = std::max(info[3], id);
                info[4]++;
            }
        }
    }
    if (next != size_global) { pfem::set_error("Something wrong with NodeDofArrayNew"); return -PFEM_ERR_NUMBERING; }
    if (nparts > 1)   // re-key the applied values to NEW node ids (:668-677)
        for (int b = 0; b < nDBC; b++) {
            const size_t n = (size_t)node_map_get_new[dbc_node[b] - 1] - 1;
            solnApplied[n * ndof + (dbc_dof[b] - 1)] = dbc_val[b];
        }
    return size_global;
}

// elemNodeConn(e,i) = node_map_get_new(elemNodeConn(e,i))   (tetrapoissonparallelimpl1.F:659-664), in place
PFEM_EXPORT void pfem_host_renumber_conn(long long n_entries, int *conn, const int *node_map_get_new)
{
    for (long long t = 0; t < n_entries; t++) conn[t] = node_map_get_new[conn[t] - 1];
}

// ElemDofArray(e, ndof*(i-1)+j) = NodeDofArrayNew(conn(e,i), j) - 1   (tetrapoissonparallelimpl1.F:698-713)
PFEM_EXPORT void pfem_host_elem_dof_array(int nElem, int npElem, int ndof, int nNode, const int *conn_new,
                                          const int *NodeDofArrayNew, int *elemDof)
{
    for (int i = 0; i < npElem; i++)
        for (int d = 0; d < ndof; d++) {
            const int *c = conn_new + (size_t)i * nElem;
            const int *nd = NodeDofArrayNew + (size_t)d * nNode;
            int *out = elemDof + (size_t)(ndof * i + d) * nElem;
            for (int e = 0; e < nElem; e++) out[e] = nd[c[e] - 1] - 1;
        }
}

// Elements a rank must hand to its GPU: every element with at least one dof in the rank's row block
// [row_lo, row_hi) (0-based) -- the owned and overlap elements -- in ascending global id.
// Two-call protocol: list == NULL returns the count.
PFEM_EXPORT int pfem_host_select_elements(int nElem, int nsize, const int *elemDof, int row_lo, int row_hi, int *list)
{
    int count = 0;
    for (int e = 0; e < nElem; e++) {
        bool touch = false;
        for (int k = 0; k < nsize && !touch; k++) {
            const int d = elemDof[(size_t)k * nElem + e];
            touch = d >= row_lo && d < row_hi;
        }
        if (touch) { if (list) list[count] = e; count++; }
    }
    return count;
}

// gather rows of a column-major (nElem x ncol) int array: out(:, j) = in(list, j)
PFEM_EXPORT void pfem_host_gather_rows(int nElem, int ncol, const int *in, int nsel, const int *list, int *out)
{
    for (int j = 0; j < ncol; j++)
        for (int t = 0; t < nsel; t++) out[(size_t)j * nsel + t] = in[(size_t)j * nElem + list[t]];
}
