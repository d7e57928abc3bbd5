// pattern.cu -- mesh upload and the pattern pass on the GPU.
//
// Replaces the MatSetValues(INSERT_VALUES, zeros) loop + MatAssembly of the reference
// (tetrapoissonparallelimpl1.F:791-802, solverpetsc.F:222-246): the CSR pattern of the owned rows is
// built from the element dof lists only (never from values: explicit zeros are part of the pattern),
// with sorted unique global columns per row, which is what PETSc's AIJ stores after final assembly.
// Also builds the row -> element incidence lists (ascending element id) that the value pass gathers from.
#include <cub/cub.cuh>

#include <cstdlib>

#include "internal.cuh"

namespace pfem {

// ---- mesh upload ---------------------------------------------------------------------------------

// erec[e] = { conn[0..npe) 0-based, dof[0..nsize), pad }, SoA -> AoS (one 32/64-byte record per element)
__global__ void pack_conn_kernel(int nElem, int npe, int rec_ints, const int *__restrict__ conn_soa, int *__restrict__ erec)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)nElem * npe;
         t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / nElem), e = (int)(t - (long long)i * nElem);   // coalesced read of the SoA column
        erec[(size_t)e * rec_ints + i] = conn_soa[t] - 1;
    }
}

__global__ void pack_dof_kernel(int nElem, int npe, int nsize, int rec_ints, const int *__restrict__ dof_soa,
                                int *__restrict__ erec)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)nElem * nsize;
         t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t / nElem), e = (int)(t - (long long)k * nElem);
        erec[(size_t)e * rec_ints + npe + k] = dof_soa[t];
    }
}

// xyz[n_new] = coords(node_map_get_old(n_new), :)   (tetrapoissonparallelimpl1.F:832-838)
__global__ void pack_xyz_kernel(int nNode, int ndim, int stride, const double *__restrict__ coords_soa,
                                const int *__restrict__ map_old, double *__restrict__ xyz)
{
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < nNode; n += gridDim.x * blockDim.x) {
        const int o = map_old ? map_old[n] - 1 : n;
        for (int c = 0; c < stride; c++) xyz[(size_t)n * stride + c] = c < ndim ? coords_soa[(size_t)c * nNode + o] : 0.0;
    }
}

__global__ void check_range_kernel(long long n, const int *__restrict__ v, int lo, int hi, int *__restrict__ bad)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        if (v[t] < lo || v[t] >= hi) atomicAdd(bad, 1);
}

static int kind_dims(int kind, int &npe, int &ndof, int &ndim)
{
    switch (kind) {
    case PFEM_POISSON_TRIA: npe = 3; ndof = 1; ndim = 2; return PFEM_OK;
    case PFEM_POISSON_TETRA: npe = 4; ndof = 1; ndim = 3; return PFEM_OK;
    case PFEM_ELASTICITY_TRIA: npe = 3; ndof = 2; ndim = 2; return PFEM_OK;
    case PFEM_ELASTICITY_TETRA: npe = 4; ndof = 3; ndim = 3; return PFEM_OK;
    }
    set_error("unknown element kind %d", kind);
    return PFEM_ERR_ARG;
}

int upload_mesh(pfem_solver *h, int kind, int nElem, const int *conn, int nNode, const double *coords,
                const int *node_map_get_old)
{
    int npe, ndof, ndim;
    PFEM_TRY(kind_dims(kind, npe, ndof, ndim));
    if (nElem <= 0 || nNode <= 0 || !conn || !coords) { set_error("pfem_solver_set_mesh: bad argument"); return PFEM_ERR_ARG; }
    const int nsize = npe * ndof;
    if ((long long)nElem * nsize >= (1LL << 31)) { set_error("nElem*nsize exceeds 2^31"); return PFEM_ERR_SIZE; }
    h->kind = kind; h->npe = npe; h->ndof = ndof; h->ndim = ndim; h->nsize = nsize;
    h->nElem = nElem; h->nNode = nNode;
    h->rec_ints = ((npe + nsize + 3) / 4) * 4;
    h->have_mesh = false; h->have_dofs = false;
    cudaStream_t s = h->stream;
    PFEM_TRY(h->erec.alloc((size_t)nElem * h->rec_ints));
    PFEM_CUDA(cudaMemsetAsync(h->erec.p, 0xFF, (size_t)nElem * h->rec_ints * sizeof(int), s));
    {
        struct { int *p; } tmp;
        PFEM_TRY(scratch_get<int>(h, 1, (size_t)nElem * npe, &tmp.p));
        PFEM_CUDA(cudaMemcpyAsync(tmp.p, conn, (size_t)nElem * npe * sizeof(int), cudaMemcpyHostToDevice, s));
        Tmp<int> bad;
        PFEM_TRY(bad.alloc(h, 24, 1));
        PFEM_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
        check_range_kernel<<<h->sm_count * 8, 256, 0, s>>>((long long)nElem * npe, tmp.p, 1, nNode + 1, bad.p);
        pack_conn_kernel<<<h->sm_count * 8, 256, 0, s>>>(nElem, npe, h->rec_ints, tmp.p, h->erec.p);
        h->launches += 2;
        int nbad = 0;
        PFEM_CUDA(cudaMemcpyAsync(&nbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaStreamSynchronize(s));
        if (nbad) { set_error("pfem_solver_set_mesh: %d connectivity entries outside 1..nNode", nbad); return PFEM_ERR_ARG; }
    }
    {
        const int stride = ndim == 3 ? 4 : 2;
        struct { double *p; } tmp;
        struct { int *p; } map;
        map.p = nullptr;
        PFEM_TRY(scratch_get<double>(h, 2, (size_t)nNode * ndim, &tmp.p));
        PFEM_CUDA(cudaMemcpyAsync(tmp.p, coords, (size_t)nNode * ndim * sizeof(double), cudaMemcpyHostToDevice, s));
        if (node_map_get_old) {
            PFEM_TRY(scratch_get<int>(h, 3, (size_t)nNode, &map.p));
            PFEM_CUDA(cudaMemcpyAsync(map.p, node_map_get_old, (size_t)nNode * sizeof(int), cudaMemcpyHostToDevice, s));
            Tmp<int> bad;
            PFEM_TRY(bad.alloc(h, 24, 1));
            PFEM_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
            check_range_kernel<<<h->sm_count * 4, 256, 0, s>>>(nNode, map.p, 1, nNode + 1, bad.p);
            h->launches++;
            int nbad = 0;
            PFEM_CUDA(cudaMemcpyAsync(&nbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
            PFEM_CUDA(cudaStreamSynchronize(s));
            if (nbad) { set_error("pfem_solver_set_mesh: node_map_get_old outside 1..nNode"); return PFEM_ERR_ARG; }
        }
        PFEM_TRY(h->xyz.alloc((size_t)nNode * stride));
        pack_xyz_kernel<<<h->sm_count * 4, 256, 0, s>>>(nNode, ndim, stride, tmp.p, node_map_get_old ? map.p : nullptr, h->xyz.p);
        h->launches++;
        PFEM_CUDA(cudaGetLastError());
        PFEM_CUDA(cudaStreamSynchronize(s));
    }
    PFEM_TRY(h->applied.alloc((size_t)nNode * ndof));
    PFEM_CUDA(cudaMemsetAsync(h->applied.p, 0, (size_t)nNode * ndof * sizeof(double), s));
    h->have_mesh = true;
    return PFEM_OK;
}

// ---- incidence lists -------------------------------------------------------------------------------

// key = owned local row of the dof at code = e*nsize+k, or the sentinel nloc (sorted to the end)
__global__ void inc_keys_kernel(long long total, int nsize, int npe, int rec_ints, const int *__restrict__ erec,
                                int row_lo, int row_hi, int *__restrict__ keys, int *__restrict__ vals)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(t / nsize), k = (int)(t - (long long)e * nsize);
        const int d = erec[(size_t)e * rec_ints + npe + k];
        keys[t] = (d >= row_lo && d < row_hi) ? d - row_lo : row_hi - row_lo;
        vals[t] = (int)t;
    }
}

// ptr[r] = lower_bound(keys_sorted, r) for r = 0..nloc
__global__ void lower_bound_kernel(int nloc, long long total, const int *__restrict__ keys, int *__restrict__ ptr)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= nloc; r += gridDim.x * blockDim.x) {
        long long lo = 0, hi = total;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (keys[mid] < r) lo = mid + 1; else hi = mid;
        }
        ptr[r] = (int)lo;
    }
}

// upper bound on the row length: sum over incident elements of their free-dof count
__global__ void cand_count_kernel(int nloc, int nsize, int npe, int rec_ints, const int *__restrict__ erec,
                                  const int *__restrict__ rinc_ptr, const int *__restrict__ rinc,
                                  long long *__restrict__ cand)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nloc; r += gridDim.x * blockDim.x) {
        long long c = 0;
        for (int m = rinc_ptr[r]; m < rinc_ptr[r + 1]; m++) {
            const int e = rinc[m] / nsize;
            const int *dof = erec + (size_t)e * rec_ints + npe;
            for (int j = 0; j < nsize; j++) c += dof[j] >= 0;
        }
        cand[r] = c;
    }
}

// per-row sorted-unique insertion of the candidate columns into scratch[cand_off[r] ...]
// The sorted list of a row is built in SHARED memory (one private strip of CAP ints per thread, interleaved by thread so that
// the 32 lanes of a warp hit 32 different banks) and written out once; only rows with more than CAP distinct columns fall
// back to building in the global scratch strip.  r01 built every list in global memory: 16 ms on C5, the largest kernel of
// the pattern pass.
static constexpr int RU_THREADS = 64, RU_CAP = 96;

__global__ void __launch_bounds__(RU_THREADS)
row_unique_kernel(int nloc, int nsize, int npe, int rec_ints, const int *__restrict__ erec,
                  const int *__restrict__ rinc_ptr, const int *__restrict__ rinc,
                  const long long *__restrict__ cand_off, int *__restrict__ scratch,
                  int *__restrict__ rowlen)
{
    __shared__ int strip[RU_CAP * RU_THREADS];
    int *loc = strip + threadIdx.x;                       // element k of this thread's list: loc[k * RU_THREADS]
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nloc; r += gridDim.x * blockDim.x) {
        int *s = scratch + cand_off[r];
        int len = 0;
        bool in_smem = true;
        for (int m = rinc_ptr[r]; m < rinc_ptr[r + 1]; m++) {
            const int e = rinc[m] / nsize;
            const int *dof = erec + (size_t)e * rec_ints + npe;
            for (int j = 0; j < nsize; j++) {
                const int c = dof[j];
                if (c < 0) continue;
                int lo = 0, hi = len;
                if (in_smem) {
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (loc[mid * RU_THREADS] < c) lo = mid + 1; else hi = mid;
                    }
                    if (lo < len && loc[lo * RU_THREADS] == c) continue;
                    if (len < RU_CAP) {
                        for (int t = len; t > lo; t--) loc[t * RU_THREADS] = loc[(t - 1) * RU_THREADS];
                        loc[lo * RU_THREADS] = c;
                        len++;
                        continue;
                    }
                    // strip full: move the list to the global scratch strip (sized by the candidate count) and go on there
                    for (int t = 0; t < len; t++) s[t] = loc[t * RU_THREADS];
                    in_smem = false;
                }
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (s[mid] < c) lo = mid + 1; else hi = mid;
                }
                if (lo < len && s[lo] == c) continue;
                for (int t = len; t > lo; t--) s[t] = s[t - 1];
                s[lo] = c;
                len++;
            }
        }
        if (in_smem)
            for (int t = 0; t < len; t++) s[t] = loc[t * RU_THREADS];
        rowlen[r] = len;
    }
}

__global__ void sum_int_kernel(int n, const int *__restrict__ v, unsigned long long *__restrict__ out)
{
    unsigned long long s = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += (unsigned long long)v[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

__global__ void compact_cols_kernel(int nloc, const long long *__restrict__ cand_off, const int *__restrict__ scratch,
                                    const int *__restrict__ rowptr, int *__restrict__ col)
{
    // one warp per row
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nloc; r += warps) {
        const int a = rowptr[r], n = rowptr[r + 1] - a;
        const int *s = scratch + cand_off[r];
        for (int k = lane; k < n; k += 32) col[a + k] = s[k];
    }
}

// ---- value-pass streams ------------------------------------------------------------------------------------------

__global__ void conn4_kernel(int nElem, int npe, int rec_ints, const int *__restrict__ erec, int *__restrict__ conn4)
{
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nElem; e += gridDim.x * blockDim.x) {
        int4 v;
        const int *r = erec + (size_t)e * rec_ints;
        v.x = r[0]; v.y = r[1]; v.z = r[2]; v.w = npe == 4 ? r[3] : r[2];
        reinterpret_cast<int4 *>(conn4)[e] = v;
    }
}

// one warp per 32-row slice: padded incidence count of the slice (in entries)
__global__ void inc_width_kernel(int nloc, int nslices, const int *__restrict__ rinc_ptr, const int *__restrict__ rowptr,
                                 long long *__restrict__ slice_sz, int *__restrict__ wide_flag)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < nslices; s += warps) {
        const int r = s * 32 + lane;
        int w = r < nloc ? rinc_ptr[r + 1] - rinc_ptr[r] : 0;
        if (r < nloc && rowptr[r + 1] - rowptr[r] > 254) atomicOr(wide_flag, 1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
        if (lane == 0) slice_sz[s] = (long long)w * 32;
    }
}

// entry = { code = e*nsize+k, slot bytes: position of dof j of element e inside the row's sorted column list, 255 = Dirichlet }
__global__ void fill_asm_inc_kernel(int nloc, int nrows_padded, int nsize, int npe, int rec_ints, int words,
                                    const int *__restrict__ erec, const int *__restrict__ rinc_ptr,
                                    const int *__restrict__ rinc, const int *__restrict__ rowptr, const int *__restrict__ col,
                                    const long long *__restrict__ inc_off, int *__restrict__ ainc)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows_padded; r += gridDim.x * blockDim.x) {
        const int s = r >> 5, lane = r & 31;
        const long long base = inc_off[s] + lane;
        const int width = (int)((inc_off[s + 1] - inc_off[s]) >> 5);
        int m = 0;
        if (r < nloc) {
            const int c0 = rowptr[r], len = rowptr[r + 1] - c0;
            for (int q = rinc_ptr[r]; q < rinc_ptr[r + 1]; q++, m++) {
                const int code = rinc[q];
                const int e = code / nsize;
                const int *dof = erec + (size_t)e * rec_ints + npe;
                unsigned int w[3] = {0u, 0u, 0u};             // unused bytes 0x00, Dirichlet dofs 0xFF
                for (int j = 0; j < nsize; j++) {
                    const int c = dof[j];
                    if (c < 0) { w[j >> 2] |= 0xFFu << (8u * (j & 3)); continue; }
                    int lo = 0, hi = len;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (col[c0 + mid] < c) lo = mid + 1; else hi = mid;
                    }
                    const unsigned int sh = 8u * (j & 3);
                    w[j >> 2] = (w[j >> 2] & ~(0xFFu << sh)) | ((unsigned int)(lo & 255) << sh);
                }
                int *out = ainc + (size_t)(base + (long long)m * 32) * words;
                out[0] = code;
                for (int q2 = 1; q2 < words; q2++) out[q2] = (int)w[q2 - 1];
            }
        }
        for (; m < width; m++) {
            int *out = ainc + (size_t)(base + (long long)m * 32) * words;
            out[0] = -1;
            for (int q2 = 1; q2 < words; q2++) out[q2] = -1;
        }
    }
}

int build_asm_streams(pfem_solver *h)
{
    StageTimer tm("value pass: row-kernel streams");
    cudaStream_t s = h->stream;
    const int G = h->sm_count * 8, nloc = h->size_local;
    const int nslices = (nloc + 31) / 32;
    h->asm_sell = false;
    h->ainc_words = h->nsize <= 4 ? 2 : 4;
    Tmp<long long> sz;
    Tmp<int> wide;
    PFEM_TRY(sz.alloc(h, 25, (size_t)nslices + 1));
    PFEM_TRY(wide.alloc(h, 26, 1));
    PFEM_CUDA(cudaMemsetAsync(sz.p, 0, ((size_t)nslices + 1) * sizeof(long long), s));
    PFEM_CUDA(cudaMemsetAsync(wide.p, 0, sizeof(int), s));
    inc_width_kernel<<<G, 256, 0, s>>>(nloc, nslices, h->rinc_ptr.p, h->rowptr.p, sz.p, wide.p);
    h->launches++;
    PFEM_TRY(h->ainc_off.alloc((size_t)nslices + 1));
    size_t bytes = 0;
    PFEM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, sz.p, h->ainc_off.p, nslices + 1, s));
    Tmp<char> tmp;
    PFEM_TRY(tmp.alloc(h, 27, bytes));
    PFEM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, sz.p, h->ainc_off.p, nslices + 1, s));
    h->launches++;
    long long nent = 0;
    int iswide = 0;
    PFEM_CUDA(cudaMemcpyAsync(&nent, h->ainc_off.p + nslices, sizeof(long long), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaMemcpyAsync(&iswide, wide.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    const char *force = getenv("PFEM_FORCE_GENERIC_ASM");                 // test hook: exercise the generic kernel
    if (iswide || nent * h->ainc_words >= (1LL << 31) || (force && force[0] == '1')) return PFEM_OK;   // generic kernel handles it
    PFEM_TRY(h->ainc.alloc((size_t)nent * h->ainc_words));
    PFEM_TRY(h->conn4.alloc((size_t)h->nElem * 4));
    conn4_kernel<<<G, 256, 0, s>>>(h->nElem, h->npe, h->rec_ints, h->erec.p, h->conn4.p);
    fill_asm_inc_kernel<<<G, 128, 0, s>>>(nloc, nslices * 32, h->nsize, h->npe, h->rec_ints, h->ainc_words, h->erec.p,
                                          h->rinc_ptr.p, h->rinc.p, h->rowptr.p, h->col.p, h->ainc_off.p, h->ainc.p);
    h->launches += 2;
    PFEM_CUDA(cudaGetLastError());
    PFEM_CUDA(cudaStreamSynchronize(s));
    h->asm_sell = true;
    return PFEM_OK;
}

// ElemDofArray(e, ndof*(i-1)+j) = NodeDofArrayNew(conn(e,i), j) - 1   (tetrapoissonparallelimpl1.F:698-713), on the GPU
__global__ void nodal_dof_kernel(int nElem, int npe, int ndof, int nNode, int rec_ints, const int *__restrict__ nda,
                                 int *__restrict__ erec)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)nElem * npe;
         t += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(t / npe), i = (int)(t - (long long)e * npe);
        int *rec = erec + (size_t)e * rec_ints;
        const int node = rec[i];
        for (int d = 0; d < ndof; d++) rec[npe + ndof * i + d] = nda[(size_t)d * nNode + node] - 1;
    }
}

int build_pattern(pfem_solver *h, int nElem, int nsize, const int *elemDof, const int *nodeDof)
{
    if (!h->initialised) { set_error("pfem_solver_set_pattern: call pfem_solver_initialise first"); return PFEM_ERR_STATE; }
    if (!h->have_mesh) { set_error("pfem_solver_set_pattern: call pfem_solver_set_mesh first"); return PFEM_ERR_STATE; }
    if (nElem != h->nElem || nsize != h->nsize || (!elemDof && !nodeDof)) {
        set_error("pfem_solver_set_pattern: nElem/nsize (%d,%d) do not match the mesh (%d,%d)", nElem, nsize, h->nElem, h->nsize);
        return PFEM_ERR_ARG;
    }
    cudaStream_t s = h->stream;
    const int G = h->sm_count * 8;
    const int nloc = h->size_local;
    const long long total = (long long)nElem * nsize;
    // 1. element dof records
    if (!elemDof) {
        StageTimer tm("pattern: nodal dofs -> element records");
        Tmp<int> bad;
        int *tmp = nullptr;
        const size_t nn = (size_t)h->nNode * h->ndof;
        PFEM_TRY(scratch_get<int>(h, 0, nn, &tmp));
        PFEM_TRY(bad.alloc(h, 24, 1));
        PFEM_CUDA(cudaMemcpyAsync(tmp, nodeDof, nn * sizeof(int), cudaMemcpyHostToDevice, s));
        PFEM_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
        check_range_kernel<<<G, 256, 0, s>>>((long long)nn, tmp, 0, h->size_global + 1, bad.p);
        nodal_dof_kernel<<<G, 256, 0, s>>>(nElem, h->npe, h->ndof, h->nNode, h->rec_ints, tmp, h->erec.p);
        h->launches += 2;
        int nbad = 0;
        PFEM_CUDA(cudaMemcpyAsync(&nbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaStreamSynchronize(s));
        if (nbad) { set_error("pfem_solver_set_pattern_nodal: %d dof ids outside 0..size_global", nbad); return PFEM_ERR_NUMBERING; }
    } else {
        StageTimer tm("pattern: upload+pack dofs");
        Tmp<int> bad;
        int *tmp = nullptr;
        PFEM_TRY(scratch_get<int>(h, 0, (size_t)total, &tmp));
        PFEM_TRY(bad.alloc(h, 24, 1));
        PFEM_CUDA(cudaMemcpyAsync(tmp, elemDof, (size_t)total * sizeof(int), cudaMemcpyHostToDevice, s));
        PFEM_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
        check_range_kernel<<<G, 256, 0, s>>>(total, tmp, -1, h->size_global, bad.p);
        pack_dof_kernel<<<G, 256, 0, s>>>(nElem, h->npe, nsize, h->rec_ints, tmp, h->erec.p);
        h->launches += 2;
        int nbad = 0;
        PFEM_CUDA(cudaMemcpyAsync(&nbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaStreamSynchronize(s));
        if (nbad) { set_error("pfem_solver_set_pattern: %d dof ids outside -1..size_global-1", nbad); return PFEM_ERR_NUMBERING; }
    }
    h->have_dofs = true;
    // 2. row -> incidence lists: stable radix sort of (local row, e*nsize+k); ties keep ascending code order
    PFEM_TRY(h->rinc_ptr.alloc((size_t)nloc + 1));
    {
        StageTimer tm("pattern: incidence sort");
        struct { int *p; } k_in, k_out, v_in, v_out;
        PFEM_TRY(scratch_get<int>(h, 0, (size_t)total, &k_in.p));
        PFEM_TRY(scratch_get<int>(h, 1, (size_t)total, &k_out.p));
        PFEM_TRY(scratch_get<int>(h, 2, (size_t)total, &v_in.p));
        PFEM_TRY(scratch_get<int>(h, 3, (size_t)total, &v_out.p));
        inc_keys_kernel<<<G, 256, 0, s>>>(total, nsize, h->npe, h->rec_ints, h->erec.p, h->row_lo, h->row_hi, k_in.p, v_in.p);
        h->launches++;
        int bits = 1;
        while ((1LL << bits) <= nloc) bits++;
        size_t tmp_bytes = 0;
        PFEM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, total, 0, bits, s));
        char *tmp = nullptr;
        PFEM_TRY(scratch_get<char>(h, 4, tmp_bytes, &tmp));
        PFEM_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, total, 0, bits, s));
        lower_bound_kernel<<<G, 256, 0, s>>>(nloc, total, k_out.p, h->rinc_ptr.p);
        h->launches += 2;
        int ninc = 0;
        PFEM_CUDA(cudaMemcpyAsync(&ninc, h->rinc_ptr.p + nloc, sizeof(int), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaStreamSynchronize(s));
        h->ninc = ninc;
        PFEM_TRY(h->rinc.alloc((size_t)ninc));
        PFEM_CUDA(cudaMemcpyAsync(h->rinc.p, v_out.p, (size_t)ninc * sizeof(int), cudaMemcpyDeviceToDevice, s));
        PFEM_CUDA(cudaStreamSynchronize(s));
    }
    // 3. pattern: per-row sorted unique columns
    {
        StageTimer tm("pattern: row unique + csr");
        Tmp<long long> cand, cand_off;                 // temporaries in the handle's persistent scratch: no cudaMalloc/cudaFree per pass
        Tmp<int> rowlen;
        PFEM_TRY(cand.alloc(h, 19, (size_t)nloc + 1));
        PFEM_TRY(cand_off.alloc(h, 20, (size_t)nloc + 1));
        PFEM_TRY(rowlen.alloc(h, 21, (size_t)nloc + 1));
        PFEM_CUDA(cudaMemsetAsync(cand.p, 0, ((size_t)nloc + 1) * sizeof(long long), s));
        PFEM_CUDA(cudaMemsetAsync(rowlen.p, 0, ((size_t)nloc + 1) * sizeof(int), s));
        cand_count_kernel<<<G, 256, 0, s>>>(nloc, nsize, h->npe, h->rec_ints, h->erec.p, h->rinc_ptr.p, h->rinc.p, cand.p);
        h->launches++;
        size_t tmp_bytes = 0, tmp_bytes2 = 0;
        PFEM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cand.p, cand_off.p, nloc + 1, s));
        PFEM_TRY(h->rowptr.alloc((size_t)nloc + 1));
        PFEM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes2, rowlen.p, h->rowptr.p, nloc + 1, s));
        Tmp<char> tmp;
        PFEM_TRY(tmp.alloc(h, 22, tmp_bytes > tmp_bytes2 ? tmp_bytes : tmp_bytes2));
        PFEM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cand.p, cand_off.p, nloc + 1, s));
        h->launches++;
        long long ncand = 0;
        PFEM_CUDA(cudaMemcpyAsync(&ncand, cand_off.p + nloc, sizeof(long long), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaStreamSynchronize(s));
        struct { int *p; } scratch;
        PFEM_TRY(scratch_get<int>(h, 5, (size_t)ncand, &scratch.p));
        row_unique_kernel<<<(nloc + RU_THREADS - 1) / RU_THREADS, RU_THREADS, 0, s>>>(nloc, nsize, h->npe, h->rec_ints, h->erec.p, h->rinc_ptr.p, h->rinc.p,
                                            cand_off.p, scratch.p, rowlen.p);
        h->launches++;
        // nnz must fit the 32-bit rowptr: reduce in 64 bits first
        Tmp<long long> nnz64;
        PFEM_TRY(nnz64.alloc(h, 23, 1));
        PFEM_CUDA(cudaMemsetAsync(nnz64.p, 0, sizeof(long long), s));
        sum_int_kernel<<<G, 256, 0, s>>>(nloc, rowlen.p, (unsigned long long *)nnz64.p);
        h->launches++;
        long long nnz = 0;
        PFEM_CUDA(cudaMemcpyAsync(&nnz, nnz64.p, sizeof(long long), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaStreamSynchronize(s));
        if (nnz >= (1LL << 31)) { set_error("pattern has %lld nonzeros on this rank: exceeds 2^31", nnz); return PFEM_ERR_SIZE; }
        PFEM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes2, rowlen.p, h->rowptr.p, nloc + 1, s));
        h->launches++;
        h->nnz = nnz;
        PFEM_TRY(h->col.alloc((size_t)nnz));
        PFEM_TRY(h->val.alloc((size_t)nnz));
        PFEM_TRY(h->rhs.alloc((size_t)nloc));
        compact_cols_kernel<<<G, 256, 0, s>>>(nloc, cand_off.p, scratch.p, h->rowptr.p, h->col.p);
        h->launches++;
        PFEM_CUDA(cudaMemsetAsync(h->val.p, 0, (size_t)(nnz > 0 ? nnz : 1) * sizeof(double), s));
        PFEM_CUDA(cudaMemsetAsync(h->rhs.p, 0, (size_t)(nloc > 0 ? nloc : 1) * sizeof(double), s));
        PFEM_CUDA(cudaGetLastError());
        PFEM_CUDA(cudaStreamSynchronize(s));
    }
    h->values_zero = true;
    h->rhs_zero = true;
    h->pattern_seq++;                  // structures derived from the pattern (ILU level schedule) are stale
    PFEM_TRY(h->neg_count.alloc(1));
    // the value-pass streams (row kernels: build_asm_streams + plan_assembly; tile kernel: build_ctiles) are built on
    // first use by assemble_values, for the kernel that actually runs
    h->asm_sell = false;
    h->rows_ready = false;
    h->asm_rows_per_cta = 0;
    h->tiles_ready = false;            // tiles of the previous pattern (if any) are stale
    h->asm_tiled = false;
    h->ct_ready = false;
    h->ct_tried = false;
    return PFEM_OK;
}

}  // namespace pfem
