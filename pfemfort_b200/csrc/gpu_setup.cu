// gpu_setup.cu -- the driver-side set-up loops on the GPU.
//
//  * pfem_gpu_number_dofs / pfem_gpu_renumber_conn / pfem_gpu_elem_dof_array: node renumbering by partition,
//    NodeDofArrayNew, row ranges, re-keyed applied values, ElemDofArray, assyForSoln and the owned + overlap element list of a rank
//    (tetrapoissonparallelimpl1.F:357-367, 402-421, 500-677, 698-734) as sorts, scans and gathers.  Same arguments and
//    bit-identical outputs as the host functions of host_driver.cu (tests/test_gpu_setup.py compares them entry by entry).
//  * pfem_gpu_gen_tetra: the structured 6-tets-per-cell box mesh of genTetra.cpp:194-334 with its Dirichlet list
//    (:497-525), generated on the GPU: per-axis coordinate tables come from the host (n+1 accumulated doubles per axis,
//    the recipe's sequential `xx += dx`), everything of size O(nodes) / O(elements) is formed by kernels, including the
//    8-decimal text round trip of the boundary values (exact: FMA residual, round-half-even, correctly rounded divide).
#include <cub/cub.cuh>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "internal.cuh"

namespace pfem {

#define GS_CHECK_DEVICE(device, name)                                                              \
    do {                                                                                           \
        int nd_ = 0;                                                                               \
        if (cudaGetDeviceCount(&nd_) != cudaSuccess || nd_ <= 0) {                                 \
            cudaGetLastError();                                                                    \
            set_error(name ": no CUDA device (there is no CPU fallback)");                        \
            return -PFEM_ERR_CUDA;                                                                 \
        }                                                                                          \
        if ((device) < 0 || (device) >= nd_) { set_error(name ": bad device"); return -PFEM_ERR_ARG; } \
        if (cudaSetDevice(device) != cudaSuccess) { set_error(name ": cudaSetDevice failed"); return -PFEM_ERR_CUDA; } \
    } while (0)

#define GS_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            pfem::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return -PFEM_ERR_CUDA;                                                           \
        }                                                                                    \
    } while (0)
#define GS_TRY(call) do { if ((call) != PFEM_OK) return -PFEM_ERR_CUDA; } while (0)

static constexpr int GS_G = 148 * 8;

// "last writer wins" of a sequential loop over the Dirichlet rows, made deterministic: the winning row index per slot
__global__ void gs_dbc_owner_kernel(int nDBC, const int *__restrict__ node, const int *__restrict__ dof, int ndof, int nNode,
                                    const int *__restrict__ map_new, int *__restrict__ owner, int *__restrict__ bad)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nDBC; b += gridDim.x * blockDim.x) {
        int n = node[b] - 1;
        const int d = dof[b] - 1;
        if (n < 0 || n >= nNode || d < 0 || d >= ndof) { atomicAdd(bad, 1); continue; }
        if (map_new) n = map_new[n] - 1;
        atomicMax(&owner[(size_t)n * ndof + d], b);
    }
}

__global__ void gs_dbc_apply_kernel(long long nd, const int *__restrict__ owner, const double *__restrict__ val,
                                    unsigned char *__restrict__ type, double *__restrict__ applied)
{
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < nd; s += (long long)gridDim.x * blockDim.x) {
        const int b = owner[s];
        if (b >= 0) { if (type) type[s] = 1; applied[s] = val[b]; }
    }
}

__global__ void gs_iota_kernel(int n, int *__restrict__ v, int base)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = i + base;
}

__global__ void gs_check_part_kernel(int n, const int *__restrict__ part, int nparts, int *__restrict__ bad)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (part[i] < 0 || part[i] >= nparts) atomicAdd(bad, 1);
}

__global__ void gs_inverse_kernel(int n, const int *__restrict__ old, int *__restrict__ nw)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) nw[old[i] - 1] = i + 1;
}

// free flag of every dof in NEW node order, node-major
__global__ void gs_free_flags_kernel(int nNode, int ndof, const int *__restrict__ old, const unsigned char *__restrict__ type_old,
                                     int *__restrict__ flag)
{
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < (long long)nNode * ndof; s += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(s / ndof), d = (int)(s - (long long)n * ndof);
        flag[s] = type_old[(size_t)(old[n] - 1) * ndof + d] ? 0 : 1;
    }
}

__global__ void gs_node_dof_kernel(int nNode, int ndof, const int *__restrict__ flag, const int *__restrict__ scan, int *__restrict__ nda)
{
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < (long long)nNode * ndof; s += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(s / ndof), d = (int)(s - (long long)n * ndof);
        nda[(size_t)d * nNode + n] = flag[s] ? scan[s] + 1 : 0;
    }
}

// part_info[p] = node_start, node_end, row_start, row_end (1-based, inclusive), size_local   (:527-533, 622-636)
__global__ void gs_part_info_kernel(int nparts, int nNode, int ndof, const int *__restrict__ keys_sorted, const int *__restrict__ scan,
                                    int total_free, int *__restrict__ info)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nparts) return;
    auto lower = [&](int key) {          // first sorted position with part id >= key
        if (!keys_sorted) return key <= 0 ? 0 : nNode;
        int lo = 0, hi = nNode;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (keys_sorted[mid] < key) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    const int n0 = lower(p), n1 = lower(p + 1);
    const long long s0 = (long long)n0 * ndof, s1 = (long long)n1 * ndof;
    const int f0 = s0 < (long long)nNode * ndof ? scan[s0] : total_free;
    const int f1 = s1 < (long long)nNode * ndof ? scan[s1] : total_free;
    int *o = info + 5 * p;
    o[0] = n0 + 1; o[1] = n1; o[4] = f1 - f0;
    o[2] = f1 > f0 ? f0 + 1 : 1000000000;
    o[3] = f1 > f0 ? f1 : -1000000000;
}

__global__ void gs_renumber_kernel(long long n, int *__restrict__ conn, const int *__restrict__ map_new, int nNode, int *__restrict__ bad)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int c = conn[t];
        if (c < 1 || c > nNode) { atomicAdd(bad, 1); continue; }
        conn[t] = map_new[c - 1];
    }
}

__global__ void gs_elem_dof_kernel(int nElem, int npe, int ndof, int nNode, const int *__restrict__ conn, const int *__restrict__ nda,
                                   int *__restrict__ edof)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)nElem * npe * ndof; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t / nElem), e = (int)(t - (long long)k * nElem);
        const int i = k / ndof, d = k - i * ndof;
        edof[t] = nda[(size_t)d * nNode + conn[(size_t)i * nElem + e] - 1] - 1;
    }
}

__global__ void gs_touch_kernel(int nElem, int nsize, const int *__restrict__ edof, int row_lo, int row_hi, int *__restrict__ flag)
{
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nElem; e += gridDim.x * blockDim.x) {
        int touch = 0;
        for (int k = 0; k < nsize; k++) {
            const int d = edof[(size_t)k * nElem + e];
            touch |= (d >= row_lo && d < row_hi);
        }
        flag[e] = touch;
    }
}

__global__ void gs_compact_kernel(int n, const int *__restrict__ flag, const int *__restrict__ scan, int *__restrict__ list)
{
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x)
        if (flag[e]) list[scan[e]] = e;
}

__global__ void gs_assy_kernel(int nNode, int ndof, const int *__restrict__ nda, int *__restrict__ assy)
{
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < (long long)nNode * ndof; s += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(s / ndof), d = (int)(s - (long long)n * ndof);
        const int id = nda[(size_t)d * nNode + n];
        if (id > 0) assy[id - 1] = (int)s + 1;           // assyForSoln(dof) = (newnode-1)*ndof + j   (:722-734)
    }
}

static int gs_exclusive_scan(const int *in, int *out, long long n, cudaStream_t s)
{
    size_t bytes = 0;
    if (cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, s) != cudaSuccess) return PFEM_ERR_CUDA;
    DevBuf<char> tmp;
    if (tmp.alloc(bytes) != PFEM_OK) return PFEM_ERR_CUDA;
    if (cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, n, s) != cudaSuccess) return PFEM_ERR_CUDA;
    return cudaStreamSynchronize(s) == cudaSuccess ? PFEM_OK : PFEM_ERR_CUDA;
}

// ---- structured tetra mesh (genTetra.cpp) ---------------------------------------------------------------------------------

// float(f"{v:.8f}") exactly: k = round-half-even(v * 1e8) decided on the exact FMA residual, then the correctly rounded k / 1e8
__device__ __forceinline__ double text_round8(double v)
{
    const double a = fabs(v);
    double k = floor(a * 1e8);
    double r = fma(a, 1e8, -k);                           // exact: the residual needs < 53 bits for |v| < 2^19
    if (r < 0.0) { k -= 1.0; r += 1.0; }
    if (r >= 1.0) { k += 1.0; r -= 1.0; }
    if (r > 0.5 || (r == 0.5 && fmod(k, 2.0) == 1.0)) k += 1.0;
    const double q = k / 1e8;
    return v < 0.0 ? -q : q;
}

__global__ void gt_nodes_kernel(int nNx, int nNy, int nNz, const double *__restrict__ rx, const double *__restrict__ ry,
                                const double *__restrict__ rz, double *__restrict__ coords)
{
    const long long nN = (long long)nNx * nNy * nNz;
    for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < nN; n += (long long)gridDim.x * blockDim.x) {
        const int ii = (int)(n % nNx), jj = (int)((n / nNx) % nNy), kk = (int)(n / ((long long)nNx * nNy));
        coords[n] = rx[ii]; coords[nN + n] = ry[jj]; coords[2 * nN + n] = rz[kk];      // node = kk*nNx*nNy + jj*nNx + ii (:194-216)
    }
}

__global__ void gt_elems_kernel(int nEx, int nEy, int nEz, int *__restrict__ conn)
{
    const int nNx = nEx + 1, nNy = nEy + 1;
    const long long ncell = (long long)nEx * nEy * nEz, nE = 6 * ncell, nn = (long long)nNx * nNy;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncell; c += (long long)gridDim.x * blockDim.x) {
        const int ii = (int)(c % nEx), jj = (int)((c / nEx) % nEy), kk = (int)(c / ((long long)nEx * nEy));
        const int p0 = (int)(nn * kk + (long long)nNx * jj + ii), p1 = p0 + 1, p2 = p0 + nNx, p3 = p2 + 1;
        const int p4 = p0 + (int)nn, p5 = p4 + 1, p6 = p4 + nNx, p7 = p6 + 1;
        const int t[6][4] = {{p0, p1, p3, p5}, {p0, p3, p2, p5}, {p2, p3, p7, p5}, {p4, p6, p7, p2}, {p4, p7, p5, p2}, {p0, p4, p5, p2}};   // :317-322
#pragma unroll
        for (int q = 0; q < 6; q++)
#pragma unroll
            for (int a = 0; a < 4; a++) conn[(size_t)a * nE + 6 * c + q] = t[q][a] + 1;
    }
}

// boundary flag: mode 0 = all six faces (Poisson), mode 1 = the y = y0 face (clamped beam)
__global__ void gt_bflag_kernel(int nNx, int nNy, int nNz, int mode, int *__restrict__ flag)
{
    const long long nN = (long long)nNx * nNy * nNz;
    for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < nN; n += (long long)gridDim.x * blockDim.x) {
        const int ii = (int)(n % nNx), jj = (int)((n / nNx) % nNy), kk = (int)(n / ((long long)nNx * nNy));
        flag[n] = mode == 0 ? (ii == 0 || ii == nNx - 1 || jj == 0 || jj == nNy - 1 || kk == 0 || kk == nNz - 1) : (jj == 0);
    }
}

__global__ void gt_dbc_kernel(int nNx, int nNy, int nNz, int mode, int ndof, const int *__restrict__ flag, const int *__restrict__ scan,
                              const double *__restrict__ fx, const double *__restrict__ fy, const double *__restrict__ fz,
                              int *__restrict__ dnode, int *__restrict__ ddof, double *__restrict__ dval)
{
    const long long nN = (long long)nNx * nNy * nNz;
    for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < nN; n += (long long)gridDim.x * blockDim.x) {
        if (!flag[n]) continue;
        const int ii = (int)(n % nNx), jj = (int)((n / nNx) % nNy), kk = (int)(n / ((long long)nNx * nNy));
        const long long o = (long long)scan[n] * (mode == 0 ? 1 : ndof);
        if (mode == 0) {
            // x^2+y^2+z^2 on the float-rounded coordinates (vtkPoints stores floats), printed with 8 decimals (:518-524)
            const double v = __dadd_rn(__dadd_rn(__dmul_rn(fx[ii], fx[ii]), __dmul_rn(fy[jj], fy[jj])), __dmul_rn(fz[kk], fz[kk]));
            dnode[o] = (int)n + 1; ddof[o] = 1; dval[o] = text_round8(v);
        } else {
            for (int d = 0; d < ndof; d++) { dnode[o + d] = (int)n + 1; ddof[o + d] = d + 1; dval[o + d] = 0.0; }
        }
    }
}

}  // namespace pfem

using namespace pfem;
#define PFEM_EXPORT extern "C" __attribute__((visibility("default")))

// Same contract as pfem_host_number_dofs (host_driver.cu): returns size_global (>= 0) or -status.
PFEM_EXPORT int pfem_gpu_number_dofs(int device, int nNode, int ndof, int nDBC, const int *dbc_node, const int *dbc_dof,
                                     const double *dbc_val, int nparts, const int *node_proc_id, int *node_map_get_old,
                                     int *node_map_get_new, int *NodeDofArrayNew, double *solnApplied, int *part_info)
{
    GS_CHECK_DEVICE(device, "pfem_gpu_number_dofs");
    if (nNode <= 0 || ndof <= 0 || nDBC < 0 || !node_map_get_old || !node_map_get_new || !NodeDofArrayNew || !solnApplied || !part_info ||
        (nparts > 1 && !node_proc_id)) { set_error("pfem_gpu_number_dofs: bad argument"); return -PFEM_ERR_ARG; }
    cudaStream_t s = nullptr;
    const long long nd = (long long)nNode * ndof;
    const int np = nparts > 1 ? nparts : 1;
    DevBuf<int> dnode, ddof, owner, bad, old, nw, keys, keys_s, vals, flag, scan, nda, info;
    DevBuf<double> dval, applied;
    DevBuf<unsigned char> type_old;
    GS_TRY(dnode.alloc(nDBC + 1)); GS_TRY(ddof.alloc(nDBC + 1)); GS_TRY(dval.alloc(nDBC + 1)); GS_TRY(owner.alloc(nd)); GS_TRY(bad.alloc(1));
    GS_TRY(old.alloc(nNode)); GS_TRY(nw.alloc(nNode)); GS_TRY(flag.alloc(nd + 1)); GS_TRY(scan.alloc(nd + 1)); GS_TRY(nda.alloc(nd));
    GS_TRY(info.alloc(5 * np)); GS_TRY(applied.alloc(nd)); GS_TRY(type_old.alloc(nd));
    if (nDBC) {
        GS_CUDA(cudaMemcpyAsync(dnode.p, dbc_node, (size_t)nDBC * sizeof(int), cudaMemcpyHostToDevice, s));
        GS_CUDA(cudaMemcpyAsync(ddof.p, dbc_dof, (size_t)nDBC * sizeof(int), cudaMemcpyHostToDevice, s));
        GS_CUDA(cudaMemcpyAsync(dval.p, dbc_val, (size_t)nDBC * sizeof(double), cudaMemcpyHostToDevice, s));
    }
    GS_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
    GS_CUDA(cudaMemsetAsync(owner.p, 0xFF, (size_t)nd * sizeof(int), s));
    GS_CUDA(cudaMemsetAsync(type_old.p, 0, (size_t)nd, s));
    GS_CUDA(cudaMemsetAsync(applied.p, 0, (size_t)nd * sizeof(double), s));
    // NodeTypeOld / solnApplied at OLD positions (:302-355)
    gs_dbc_owner_kernel<<<GS_G, 256, 0, s>>>(nDBC, dnode.p, ddof.p, ndof, nNode, nullptr, owner.p, bad.p);
    gs_dbc_apply_kernel<<<GS_G, 256, 0, s>>>(nd, owner.p, dval.p, type_old.p, applied.p);
    // new -> old map: identity, or every part's ascending list of owned old ids = a STABLE sort by part id (:543-564)
    const int *keys_sorted = nullptr;
    if (nparts > 1) {
        GS_TRY(keys.alloc(nNode)); GS_TRY(keys_s.alloc(nNode)); GS_TRY(vals.alloc(nNode));
        GS_CUDA(cudaMemcpyAsync(keys.p, node_proc_id, (size_t)nNode * sizeof(int), cudaMemcpyHostToDevice, s));
        gs_check_part_kernel<<<GS_G, 256, 0, s>>>(nNode, keys.p, nparts, bad.p);
        gs_iota_kernel<<<GS_G, 256, 0, s>>>(nNode, vals.p, 1);
        int bits = 1;
        while ((1 << bits) < nparts) bits++;
        size_t bytes = 0;
        GS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.p, keys_s.p, vals.p, old.p, nNode, 0, bits, s));
        DevBuf<char> tmp;
        GS_TRY(tmp.alloc(bytes));
        GS_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys.p, keys_s.p, vals.p, old.p, nNode, 0, bits, s));
        GS_CUDA(cudaStreamSynchronize(s));
        keys_sorted = keys_s.p;
    } else {
        gs_iota_kernel<<<GS_G, 256, 0, s>>>(nNode, old.p, 1);
    }
    int nbad = 0;
    GS_CUDA(cudaMemcpyAsync(&nbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    GS_CUDA(cudaStreamSynchronize(s));
    if (nbad) { set_error("pfem_gpu_number_dofs: %d Dirichlet rows / node_proc_id entries out of range", nbad); return -PFEM_ERR_NUMBERING; }
    gs_inverse_kernel<<<GS_G, 256, 0, s>>>(nNode, old.p, nw.p);
    // free dofs numbered in NEW node order, node-major (:601-616): exclusive scan of the free flags
    gs_free_flags_kernel<<<GS_G, 256, 0, s>>>(nNode, ndof, old.p, type_old.p, flag.p);
    GS_CUDA(cudaMemsetAsync(flag.p + nd, 0, sizeof(int), s));
    if (gs_exclusive_scan(flag.p, scan.p, nd + 1, s) != PFEM_OK) { set_error("pfem_gpu_number_dofs: scan failed"); return -PFEM_ERR_CUDA; }
    int size_global = 0;
    GS_CUDA(cudaMemcpyAsync(&size_global, scan.p + nd, sizeof(int), cudaMemcpyDeviceToHost, s));
    GS_CUDA(cudaStreamSynchronize(s));
    gs_node_dof_kernel<<<GS_G, 256, 0, s>>>(nNode, ndof, flag.p, scan.p, nda.p);
    gs_part_info_kernel<<<1, 64, 0, s>>>(np, nNode, ndof, keys_sorted, scan.p, size_global, info.p);
    if (nparts > 1) {      // re-key the applied values to NEW node ids; the stale old-position entries stay (:668-677)
        GS_CUDA(cudaMemsetAsync(owner.p, 0xFF, (size_t)nd * sizeof(int), s));
        gs_dbc_owner_kernel<<<GS_G, 256, 0, s>>>(nDBC, dnode.p, ddof.p, ndof, nNode, nw.p, owner.p, bad.p);
        gs_dbc_apply_kernel<<<GS_G, 256, 0, s>>>(nd, owner.p, dval.p, nullptr, applied.p);
    }
    GS_CUDA(cudaMemcpyAsync(node_map_get_old, old.p, (size_t)nNode * sizeof(int), cudaMemcpyDeviceToHost, s));
    GS_CUDA(cudaMemcpyAsync(node_map_get_new, nw.p, (size_t)nNode * sizeof(int), cudaMemcpyDeviceToHost, s));
    GS_CUDA(cudaMemcpyAsync(NodeDofArrayNew, nda.p, (size_t)nd * sizeof(int), cudaMemcpyDeviceToHost, s));
    GS_CUDA(cudaMemcpyAsync(solnApplied, applied.p, (size_t)nd * sizeof(double), cudaMemcpyDeviceToHost, s));
    GS_CUDA(cudaMemcpyAsync(part_info, info.p, (size_t)5 * np * sizeof(int), cudaMemcpyDeviceToHost, s));
    GS_CUDA(cudaStreamSynchronize(s));
    GS_CUDA(cudaGetLastError());
    return size_global;
}

// elemNodeConn(e,i) = node_map_get_new(elemNodeConn(e,i)) in place (:659-664)
PFEM_EXPORT int pfem_gpu_renumber_conn(int device, long long n_entries, int *conn, int nNode, const int *node_map_get_new)
{
    GS_CHECK_DEVICE(device, "pfem_gpu_renumber_conn");
    if (n_entries <= 0 || !conn || !node_map_get_new) { set_error("pfem_gpu_renumber_conn: bad argument"); return -PFEM_ERR_ARG; }
    DevBuf<int> c, m, bad;
    GS_TRY(c.alloc((size_t)n_entries)); GS_TRY(m.alloc(nNode)); GS_TRY(bad.alloc(1));
    GS_CUDA(cudaMemcpy(c.p, conn, (size_t)n_entries * sizeof(int), cudaMemcpyHostToDevice));
    GS_CUDA(cudaMemcpy(m.p, node_map_get_new, (size_t)nNode * sizeof(int), cudaMemcpyHostToDevice));
    GS_CUDA(cudaMemset(bad.p, 0, sizeof(int)));
    gs_renumber_kernel<<<GS_G, 256>>>(n_entries, c.p, m.p, nNode, bad.p);
    int nbad = 0;
    GS_CUDA(cudaMemcpy(&nbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (nbad) { set_error("pfem_gpu_renumber_conn: %d entries outside 1..nNode", nbad); return -PFEM_ERR_ARG; }
    GS_CUDA(cudaMemcpy(conn, c.p, (size_t)n_entries * sizeof(int), cudaMemcpyDeviceToHost));
    return PFEM_OK;
}

// ElemDofArray (:698-713), assyForSoln (:722-734; may be NULL) and, when list != NULL, the rank's owned + overlap elements
// (every element with a dof in [row_lo, row_hi), ascending id); returns the number of listed elements (or -status)
PFEM_EXPORT int pfem_gpu_elem_dof_array(int device, int nElem, int npElem, int ndof, int nNode, const int *conn_new,
                                        const int *NodeDofArrayNew, int size_global, int *elemDof, int *assyForSoln, int row_lo,
                                        int row_hi, int *list)
{
    GS_CHECK_DEVICE(device, "pfem_gpu_elem_dof_array");
    if (nElem <= 0 || !conn_new || !NodeDofArrayNew || !elemDof) { set_error("pfem_gpu_elem_dof_array: bad argument"); return -PFEM_ERR_ARG; }
    const int nsize = npElem * ndof;
    DevBuf<int> c, nda, ed, assy, flag, scan;
    GS_TRY(c.alloc((size_t)nElem * npElem)); GS_TRY(nda.alloc((size_t)nNode * ndof)); GS_TRY(ed.alloc((size_t)nElem * nsize));
    GS_CUDA(cudaMemcpy(c.p, conn_new, (size_t)nElem * npElem * sizeof(int), cudaMemcpyHostToDevice));
    GS_CUDA(cudaMemcpy(nda.p, NodeDofArrayNew, (size_t)nNode * ndof * sizeof(int), cudaMemcpyHostToDevice));
    gs_elem_dof_kernel<<<GS_G, 256>>>(nElem, npElem, ndof, nNode, c.p, nda.p, ed.p);
    GS_CUDA(cudaMemcpy(elemDof, ed.p, (size_t)nElem * nsize * sizeof(int), cudaMemcpyDeviceToHost));
    if (assyForSoln) {
        GS_TRY(assy.alloc((size_t)size_global + 1));
        gs_assy_kernel<<<GS_G, 256>>>(nNode, ndof, nda.p, assy.p);
        GS_CUDA(cudaMemcpy(assyForSoln, assy.p, (size_t)size_global * sizeof(int), cudaMemcpyDeviceToHost));
    }
    int count = 0;
    if (list) {
        GS_TRY(flag.alloc((size_t)nElem + 1)); GS_TRY(scan.alloc((size_t)nElem + 1));
        gs_touch_kernel<<<GS_G, 256>>>(nElem, nsize, ed.p, row_lo, row_hi, flag.p);
        GS_CUDA(cudaMemset(flag.p + nElem, 0, sizeof(int)));
        if (gs_exclusive_scan(flag.p, scan.p, (long long)nElem + 1, nullptr) != PFEM_OK) { set_error("scan failed"); return -PFEM_ERR_CUDA; }
        GS_CUDA(cudaMemcpy(&count, scan.p + nElem, sizeof(int), cudaMemcpyDeviceToHost));
        DevBuf<int> out;
        GS_TRY(out.alloc((size_t)count + 1));
        gs_compact_kernel<<<GS_G, 256>>>(nElem, flag.p, scan.p, out.p);
        GS_CUDA(cudaMemcpy(list, out.p, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost));
    }
    GS_CUDA(cudaGetLastError());
    return count;
}

// genTetra.cpp on the GPU.  ax/ay/az: the accumulated axis coordinates (nE+1 each, `xx += dx` in double) BEFORE the text
// round trip; the function rounds them to 8 decimals for the node file, to float for the Dirichlet values.
// dbc_mode 0: all six faces, value x^2+y^2+z^2 (Poisson);  1: y = y0 face clamped, ndof zeros per node.
// Two-call protocol: with coords == NULL returns the number of Dirichlet rows.  conn SoA [4][6*cells] 1-based, coords SoA.
PFEM_EXPORT long long pfem_gpu_gen_tetra(int device, int nEx, int nEy, int nEz, const double *ax, const double *ay, const double *az,
                                         int dbc_mode, int ndof, double *coords, int *conn, int *dbc_node, int *dbc_dof, double *dbc_val)
{
    GS_CHECK_DEVICE(device, "pfem_gpu_gen_tetra");
    if (nEx <= 0 || nEy <= 0 || nEz <= 0 || !ax || !ay || !az || ndof < 1 || dbc_mode < 0 || dbc_mode > 1) { set_error("pfem_gpu_gen_tetra: bad argument"); return -PFEM_ERR_ARG; }
    const int nNx = nEx + 1, nNy = nEy + 1, nNz = nEz + 1;
    const long long nN = (long long)nNx * nNy * nNz, nE = 6LL * nEx * nEy * nEz;
    if (nN >= (1LL << 31) || nE * 4 >= (1LL << 33)) { set_error("pfem_gpu_gen_tetra: mesh too large for 32-bit ids"); return -PFEM_ERR_SIZE; }
    const long long nB = dbc_mode == 0 ? nN - (long long)(nNx - 2) * (nNy - 2) * (nNz - 2) : (long long)nNx * nNz;
    const long long nRows = dbc_mode == 0 ? nB : nB * ndof;
    if (!coords) return nRows;
    if (!conn || !dbc_node || !dbc_dof || !dbc_val) { set_error("pfem_gpu_gen_tetra: NULL output"); return -PFEM_ERR_ARG; }
    // per-axis tables: text-rounded coordinates and float-rounded coordinates (host: n+1 values per axis)
    std::vector<double> tab;
    auto axis = [&](const double *a, int n) {
        for (int i = 0; i < n; i++) {
            char buf[64];
            snprintf(buf, sizeof buf, "%.8f", a[i]);       // genTetra.cpp:187-189 fixed, precision(8)
            tab.push_back(strtod(buf, nullptr));
        }
        for (int i = 0; i < n; i++) tab.push_back((double)(float)a[i]);
    };
    axis(ax, nNx); axis(ay, nNy); axis(az, nNz);
    DevBuf<double> dtab, dco, dv;
    DevBuf<int> dconn, flag, scan, dn, dd;
    GS_TRY(dtab.alloc(tab.size())); GS_TRY(dco.alloc((size_t)3 * nN)); GS_TRY(dconn.alloc((size_t)4 * nE));
    GS_TRY(flag.alloc((size_t)nN + 1)); GS_TRY(scan.alloc((size_t)nN + 1));
    GS_TRY(dn.alloc((size_t)nRows + 1)); GS_TRY(dd.alloc((size_t)nRows + 1)); GS_TRY(dv.alloc((size_t)nRows + 1));
    GS_CUDA(cudaMemcpy(dtab.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    const double *rx = dtab.p, *fx = rx + nNx, *ry = fx + nNx, *fy = ry + nNy, *rz = fy + nNy, *fz = rz + nNz;
    gt_nodes_kernel<<<GS_G, 256>>>(nNx, nNy, nNz, rx, ry, rz, dco.p);
    gt_elems_kernel<<<GS_G, 256>>>(nEx, nEy, nEz, dconn.p);
    gt_bflag_kernel<<<GS_G, 256>>>(nNx, nNy, nNz, dbc_mode, flag.p);
    GS_CUDA(cudaMemset(flag.p + nN, 0, sizeof(int)));
    if (gs_exclusive_scan(flag.p, scan.p, nN + 1, nullptr) != PFEM_OK) { set_error("scan failed"); return -PFEM_ERR_CUDA; }
    gt_dbc_kernel<<<GS_G, 256>>>(nNx, nNy, nNz, dbc_mode, ndof, flag.p, scan.p, fx, fy, fz, dn.p, dd.p, dv.p);
    GS_CUDA(cudaGetLastError());
    GS_CUDA(cudaMemcpy(coords, dco.p, (size_t)3 * nN * sizeof(double), cudaMemcpyDeviceToHost));
    GS_CUDA(cudaMemcpy(conn, dconn.p, (size_t)4 * nE * sizeof(int), cudaMemcpyDeviceToHost));
    GS_CUDA(cudaMemcpy(dbc_node, dn.p, (size_t)nRows * sizeof(int), cudaMemcpyDeviceToHost));
    GS_CUDA(cudaMemcpy(dbc_dof, dd.p, (size_t)nRows * sizeof(int), cudaMemcpyDeviceToHost));
    GS_CUDA(cudaMemcpy(dbc_val, dv.p, (size_t)nRows * sizeof(double), cudaMemcpyDeviceToHost));
    return nRows;
}
