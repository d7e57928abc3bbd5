// cg.cu -- the KSP replacement: Jacobi-preconditioned CG with PETSc's KSPSolve_CG semantics.
//
// Replaces PetscSolver%solve -> KSPSolve (solverpetsc.F:431-490) and, inside PETSc, KSPSolve_CG,
// PCApply_Jacobi, MatMult_MPIAIJ and VecDot/VecNorm/VecAXPY/VecAYPX.
//
//  * Matrix layout for the solve: the owned rows are split like PETSc's MPIAIJ into a diagonal block
//    (columns owned by this rank) and an off-diagonal block (ghost columns).  The diagonal block is stored
//    as SELL-32 (slices of 32 rows, column-major inside a slice): one thread per row, one warp per slice,
//    every load of values/indices is a fully coalesced 256/128-byte warp access and there is no
//    intra-row reduction.  Entries of a row are visited in ascending column order, like MatMult_SeqAIJ.
//  * All CG scalars live on the device (CgState).  Each reduction is finalised by the last CTA to finish
//    (fixed-order sum of the per-CTA partials => run-to-run deterministic); kernels early-out once
//    `reason` is set, so the host only polls every few iterations and never stalls the pipeline.
//  * Per iteration: direction (p = z + b p), SpMV fused with p.w, update (x, r, z, z.z, z.r fused).
#include <cooperative_groups.h>
#include <cub/cub.cuh>

#include <cstdlib>
#include <cstring>

#include "internal.cuh"

namespace cgx = cooperative_groups;

namespace pfem {

static constexpr int CG_THREADS = 256;

// Build with -DPFEM_PCG_TRACE to time-stamp (clock64, CTA 0 / thread 0) the stages of the persistent kernel's last
// iteration; read back with pfem_debug_pcg_trace (tools/pcg_trace.py).  Compiled out by default.
#ifdef PFEM_PCG_TRACE
__device__ long long g_pcg_trace[64];
#define PCG_T(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_pcg_trace[(k)] = clock64(); } while (0)
#else
#define PCG_T(k) do { } while (0)
#endif

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// CTA-wide sum, result valid in thread 0.  Fixed tree => deterministic.
__device__ __forceinline__ double block_sum(double v, double *sh)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double t = 0.0;
    if (wid == 0) {
        t = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
        t = warp_sum(t);
    }
    return t;
}

// Last-CTA-done: returns true (in every thread of the last CTA) once all CTAs have published partials.
__device__ __forceinline__ bool last_block(unsigned int *ticket)
{
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}

// fixed-order sum of n partials by one CTA (thread 0 gets the result)
__device__ __forceinline__ double reduce_partials(const double *partials, int n, double *sh)
{
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += __ldcg(partials + i);
    return block_sum(s, sh);
}


// ---- peer-memory exchange primitives (NVLink, CUDA IPC mapped) ------------------------------------------------------

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.global.release.sys.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.global.acquire.sys.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double *p)
{
    double v;
    asm volatile("ld.global.relaxed.sys.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double *p, double v)
{
    asm volatile("st.global.relaxed.sys.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}

// Tag-validated halo entries: the sender writes {value, tag} with ONE 16-byte store over NVLink, the consumer spins on the
// entry it needs until the tag is this iteration's.  No fence, no flag, no ticket: a rank never waits for data it does not
// read, and nobody waits for the sender's remote stores to be acknowledged.
__device__ __forceinline__ void st_ghost_tagged(double *p, double v, unsigned long long tag)
{
    asm volatile("st.global.relaxed.sys.v2.b64 [%0], {%1, %2};" :: "l"(p), "l"(__double_as_longlong(v)), "l"(tag) : "memory");
}
__device__ __forceinline__ void ld_ghost_raw(const double *p, long long &bits, unsigned long long &tg)
{
    asm volatile("ld.global.relaxed.sys.v2.b64 {%0, %1}, [%2];" : "=l"(bits), "=l"(tg) : "l"(p) : "memory");
}
__device__ __forceinline__ double ld_ghost_tagged(const double *p, unsigned long long tag, bool &ok)
{
    long long bits;
    unsigned long long tg;
    ld_ghost_raw(p, bits, tg);
    if (tg != tag) {
        const long long t0 = clock64();
        do {
            if (clock64() - t0 > 20000000000LL) { ok = false; break; }     // ~10 s: a dead peer must not hang the GPU
            ld_ghost_raw(p, bits, tg);
        } while (tg != tag);
    }
    return __longlong_as_double(bits);
}
// off-diagonal part of one row: sum_q bval[q] * ghost[bcol[q]] in ascending q, the tagged entries loaded four at a time
// (independent loads in flight; only an entry whose tag is still old is polled again)
__device__ __forceinline__ double offdiag_row_tagged(const double *__restrict__ bval, const int *__restrict__ bcol, const double *ghost_t,
                                                     int lo, int hi, unsigned long long tag, bool &ok)
{
    double osum = 0.0;
    int q = lo;
    for (; q + 4 <= hi; q += 4) {
        const double *p0 = ghost_t + 2 * (size_t)bcol[q], *p1 = ghost_t + 2 * (size_t)bcol[q + 1];
        const double *p2 = ghost_t + 2 * (size_t)bcol[q + 2], *p3 = ghost_t + 2 * (size_t)bcol[q + 3];
        long long b0, b1, b2, b3;
        unsigned long long t0, t1, t2, t3;
        ld_ghost_raw(p0, b0, t0); ld_ghost_raw(p1, b1, t1); ld_ghost_raw(p2, b2, t2); ld_ghost_raw(p3, b3, t3);
        const double g0 = t0 == tag ? __longlong_as_double(b0) : ld_ghost_tagged(p0, tag, ok);
        const double g1 = t1 == tag ? __longlong_as_double(b1) : ld_ghost_tagged(p1, tag, ok);
        const double g2 = t2 == tag ? __longlong_as_double(b2) : ld_ghost_tagged(p2, tag, ok);
        const double g3 = t3 == tag ? __longlong_as_double(b3) : ld_ghost_tagged(p3, tag, ok);
        osum = fma(bval[q], g0, osum); osum = fma(bval[q + 1], g1, osum);
        osum = fma(bval[q + 2], g2, osum); osum = fma(bval[q + 3], g3, osum);
    }
    for (; q < hi; q++) osum = fma(bval[q], ld_ghost_tagged(ghost_t + 2 * (size_t)bcol[q], tag, ok), osum);
    return osum;
}

__device__ __forceinline__ unsigned long long p2p_tag(const CgState *st, int kind)
{
    return (st->seq << 32) | (unsigned long long)(4u * (unsigned int)st->iter + (unsigned int)kind);
}

// spin until *flag == tag (bounded: a dead peer must not hang the GPU); returns false on time-out
__device__ __forceinline__ bool p2p_wait(const unsigned long long *flag, unsigned long long tag)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) != tag) {
        if (clock64() - t0 > 20000000000LL) return false;     // ~10 s
        __nanosleep(20);
    }
    return true;
}

__device__ __forceinline__ void st_slot_sys(P2pSlot *p, double v, unsigned long long tag)
{
    asm volatile("st.global.relaxed.sys.v2.b64 [%0], {%1, %2};" :: "l"(p), "l"(__double_as_longlong(v)), "l"(tag) : "memory");
}
__device__ __forceinline__ void ld_slot_sys(const P2pSlot *p, double &v, unsigned long long &tag)
{
    long long bits;
    asm volatile("ld.global.relaxed.sys.v2.b64 {%0, %1}, [%2];" : "=l"(bits), "=l"(tag) : "l"(p) : "memory");
    v = __longlong_as_double(bits);
}

// All-reduce (sum) of up to 4 doubles across the ranks through the peers' mailboxes, executed by the FIRST WARP of one
// CTA.  Each contribution travels as one 16-byte {value, tag} store over NVLink (no fence, no separate flag); every
// rank sums the P contributions in rank order, so all ranks obtain bit-identical results.
template <int NV>
__device__ __forceinline__ bool p2p_allreduce(const P2pCtx *c, CgState *st, int phase, unsigned long long tag, double (&v)[NV],
                                              double *sh)
{
    const int lane = threadIdx.x;      // caller guarantees threadIdx.x < 32
    const int P = c->nranks, me = c->rank;
    if (lane < P) {
        P2pMail *dst = c->mail[lane];
#pragma unroll
        for (int i = 0; i < NV; i++) st_slot_sys(&dst->red[phase][me][i], v[i], tag);
    }
    bool ok = true;
    if (lane < P) {
        const P2pMail *mine = c->mail[me];
        const long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < NV; i++) {
            double val;
            unsigned long long tg;
            ld_slot_sys(&mine->red[phase][lane][i], val, tg);
            while (tg != tag) {
                if (clock64() - t0 > 20000000000LL) { ok = false; break; }     // ~10 s: a dead peer must not hang the GPU
                ld_slot_sys(&mine->red[phase][lane][i], val, tg);
            }
            sh[NV * lane + i] = val;
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double s0 = 0.0;
        for (int q = 0; q < P; q++) s0 += sh[NV * q + i];
        v[i] = s0;
    }
    __syncwarp();
    if (!ok && lane == 0 && st) st->reason = -101;             // peer exchange timed out
    return ok;
}

__device__ __forceinline__ bool p2p_allreduce2(const P2pCtx *c, CgState *st, int phase, unsigned long long tag, double &v0,
                                               double &v1, double *sh)
{
    double v[2] = {v0, v1};
    const bool ok = p2p_allreduce<2>(c, st, phase, tag, v, sh);
    v0 = v[0]; v1 = v[1];
    return ok;
}

// ---- scalar steps of KSPSolve_CG (PETSc 3.6 cg.c), executed by one thread -------------------------------------

__device__ int converged_default(CgState *st, int it, double rnorm)
{
    if (it == 0) { st->ttol = fmax(st->rtol * rnorm, st->abstol); st->rnorm0 = rnorm; }
    if (isnan(rnorm) || isinf(rnorm)) return PFEM_DIVERGED_NANORINF;
    if (rnorm <= st->ttol) return rnorm < st->abstol ? PFEM_CONVERGED_ATOL : PFEM_CONVERGED_RTOL;
    if (rnorm >= st->dtol * st->rnorm0) return PFEM_DIVERGED_DTOL;
    return 0;
}

__device__ void begin_iteration(CgState *st)
{
    st->its = st->iter + 1;
    if (st->beta == 0.0) { st->reason = PFEM_CONVERGED_ATOL; return; }
    if (st->iter > 0 && st->beta * st->betaold < 0.0) { st->reason = PFEM_DIVERGED_INDEFINITE_PC; return; }
    st->b = st->iter == 0 ? 0.0 : st->beta / st->betaold;
}

__device__ void step_after_setup(CgState *st, double zz, double zr)
{
    st->dp = sqrt(zz);
    st->its = 0; st->iter = 0; st->dpi = 0.0; st->betaold = 0.0;
    st->reason = converged_default(st, 0, st->dp);
    if (st->reason) return;
    st->beta = zr;
    begin_iteration(st);
}

__device__ void step_after_spmv(CgState *st, double pw)
{
    st->dpiold = st->dpi;
    st->dpi = pw;
    st->betaold = st->beta;
    if (pw == 0.0 || (st->iter > 0 && pw * st->dpiold <= 0.0)) { st->reason = PFEM_DIVERGED_INDEFINITE_MAT; return; }
    st->a = st->beta / pw;
}

__device__ void step_after_update(CgState *st, double zz, double zr)
{
    st->dp = sqrt(zz);
    st->reason = converged_default(st, st->iter + 1, st->dp);
    if (st->reason) return;
    st->beta = zr;
    st->iter++;
    if (st->iter >= st->max_it) { st->reason = PFEM_DIVERGED_ITS; return; }
    begin_iteration(st);
}

// single-thread kernels used when the sums had to cross ranks first (nranks > 1)
__global__ void scalar_after_setup_kernel(CgState *st) { if (st->reason == 0 || st->iter < 0) step_after_setup(st, st->red[0], st->red[1]); }
__global__ void scalar_after_spmv_kernel(CgState *st) { if (st->reason == 0) step_after_spmv(st, st->red[0]); }
__global__ void scalar_after_update_kernel(CgState *st) { if (st->reason == 0) step_after_update(st, st->red[0], st->red[1]); }

// ---- solver-structure construction (pattern time) -----------------------------------------------------------------

__global__ void classify_rows_kernel(int nloc, int row_lo, int row_hi, const int *__restrict__ rowptr,
                                     const int *__restrict__ col, int *__restrict__ ndiag, int *__restrict__ noff)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nloc; r += gridDim.x * blockDim.x) {
        int nd = 0, no = 0;
        for (int k = rowptr[r]; k < rowptr[r + 1]; k++) {
            const int c = col[k];
            if (c >= row_lo && c < row_hi) nd++; else no++;
        }
        ndiag[r] = nd;
        noff[r] = no;
    }
}

// one warp per slice: padded slice size = 32 * max(ndiag)
__global__ void slice_width_kernel(int nloc, int nslices, const int *__restrict__ ndiag, long long *__restrict__ slice_sz)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < nslices; s += warps) {
        const int r = s * 32 + lane;
        int w = r < nloc ? ndiag[r] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
        if (lane == 0) slice_sz[s] = (long long)w * 32;
    }
}

__global__ void gather_offdiag_cols_kernel(int nloc, int row_lo, int row_hi, const int *__restrict__ rowptr,
                                           const int *__restrict__ col, const int *__restrict__ off_ptr,
                                           int *__restrict__ out)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nloc; r += gridDim.x * blockDim.x) {
        int o = off_ptr[r];
        for (int k = rowptr[r]; k < rowptr[r + 1]; k++) {
            const int c = col[k];
            if (c < row_lo || c >= row_hi) out[o++] = c;
        }
    }
}

__global__ void fill_structures_kernel(int nloc, int nrows_padded, int row_lo, int row_hi, const int *__restrict__ rowptr,
                                       const int *__restrict__ col, const long long *__restrict__ slice_off,
                                       const int *__restrict__ off_ptr, const int *__restrict__ ghost, int n_ghost,
                                       int *__restrict__ sell_col, int *__restrict__ bcol, int *__restrict__ csr2sell)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows_padded; r += gridDim.x * blockDim.x) {
        const int s = r >> 5, lane = r & 31;
        const long long base = slice_off[s] + lane;
        const int width = (int)((slice_off[s + 1] - slice_off[s]) >> 5);
        int kd = 0;
        if (r < nloc) {
            int ko = off_ptr[r];
            for (int k = rowptr[r]; k < rowptr[r + 1]; k++) {
                const int c = col[k];
                if (c >= row_lo && c < row_hi) {
                    const long long d = base + (long long)kd * 32;
                    sell_col[d] = c - row_lo;
                    csr2sell[k] = (int)d;
                    kd++;
                } else {
                    int lo = 0, hi = n_ghost;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (ghost[mid] < c) lo = mid + 1; else hi = mid;
                    }
                    bcol[ko] = lo;
                    csr2sell[k] = -ko - 1;
                    ko++;
                }
            }
        }
        // padding: value 0 times a valid local entry (the row itself, or row 0 past the end)
        const int self = r < nloc ? r : 0;
        for (; kd < width; kd++) sell_col[base + (long long)kd * 32] = self;
    }
}

__global__ void boundary_rows_kernel(int nloc, const int *__restrict__ noff, const int *__restrict__ brow_rank,
                                     const int *__restrict__ off_ptr, int *__restrict__ brow_ids, int *__restrict__ brow_ptr,
                                     int n_brows, int nnz_off)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nloc; r += gridDim.x * blockDim.x) {
        if (noff[r] > 0) {
            const int q = brow_rank[r];
            brow_ids[q] = r;
            brow_ptr[q] = off_ptr[r];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) brow_ptr[n_brows] = nnz_off;
}

__global__ void flag_kernel(int n, const int *__restrict__ v, int *__restrict__ f)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) f[i] = v[i] > 0 ? 1 : 0;
}

template <typename T>
static int exclusive_scan(pfem_solver *h, const T *in, T *out, int n)
{
    size_t bytes = 0;
    PFEM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, h->stream));
    Tmp<char> tmp;
    PFEM_TRY(tmp.alloc(h, 18, bytes));
    PFEM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, n, h->stream));
    h->launches++;
    return PFEM_OK;
}

int build_solver_structures(pfem_solver *h)
{
    cudaStream_t s = h->stream;
    const int nloc = h->size_local, G = h->sm_count * 8;
    const int nslices = (nloc + 31) / 32, nrows_padded = nslices * 32;
    Tmp<int> ndiag, noff, flags, brow_rank;          // temporaries live in the handle's persistent scratch (internal.cuh)
    DevBuf<int> &off_ptr = h->off_ptr;
    Tmp<long long> slice_sz;
    PFEM_TRY(ndiag.alloc(h, 8, (size_t)nloc + 1));
    PFEM_TRY(noff.alloc(h, 9, (size_t)nloc + 1));
    PFEM_TRY(off_ptr.alloc((size_t)nloc + 1));
    PFEM_TRY(flags.alloc(h, 10, (size_t)nloc + 1));
    PFEM_TRY(brow_rank.alloc(h, 11, (size_t)nloc + 1));
    PFEM_TRY(slice_sz.alloc(h, 12, (size_t)nslices + 1));
    PFEM_CUDA(cudaMemsetAsync(ndiag.p, 0, ((size_t)nloc + 1) * sizeof(int), s));
    PFEM_CUDA(cudaMemsetAsync(noff.p, 0, ((size_t)nloc + 1) * sizeof(int), s));
    PFEM_CUDA(cudaMemsetAsync(flags.p, 0, ((size_t)nloc + 1) * sizeof(int), s));
    PFEM_CUDA(cudaMemsetAsync(slice_sz.p, 0, ((size_t)nslices + 1) * sizeof(long long), s));
    classify_rows_kernel<<<G, 256, 0, s>>>(nloc, h->row_lo, h->row_hi, h->rowptr.p, h->col.p, ndiag.p, noff.p);
    slice_width_kernel<<<G, 256, 0, s>>>(nloc, nslices, ndiag.p, slice_sz.p);
    flag_kernel<<<G, 256, 0, s>>>(nloc, noff.p, flags.p);
    h->launches += 3;
    h->A.nrows = nloc; h->A.nslices = nslices;
    PFEM_TRY(h->A.slice_off.alloc((size_t)nslices + 1));
    PFEM_TRY(exclusive_scan<long long>(h, slice_sz.p, h->A.slice_off.p, nslices + 1));
    PFEM_TRY(exclusive_scan<int>(h, noff.p, off_ptr.p, nloc + 1));
    PFEM_TRY(exclusive_scan<int>(h, flags.p, brow_rank.p, nloc + 1));
    long long nstored = 0;
    int nnz_off = 0, n_brows = 0;
    PFEM_CUDA(cudaMemcpyAsync(&nstored, h->A.slice_off.p + nslices, sizeof(long long), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaMemcpyAsync(&nnz_off, off_ptr.p + nloc, sizeof(int), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaMemcpyAsync(&n_brows, brow_rank.p + nloc, sizeof(int), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    if (nstored >= (1LL << 31)) { set_error("SELL storage (%lld entries) exceeds 2^31", nstored); return PFEM_ERR_SIZE; }
    h->A.nstored = nstored; h->nnz_off = nnz_off; h->n_brows = n_brows;
    PFEM_TRY(h->A.col.alloc((size_t)nstored));
    PFEM_TRY(h->A.val.alloc((size_t)nstored));
    PFEM_CUDA(cudaMemsetAsync(h->A.val.p, 0, (size_t)(nstored > 0 ? nstored : 1) * sizeof(double), s));
    PFEM_TRY(h->csr2sell.alloc((size_t)h->nnz));
    PFEM_TRY(h->bcol.alloc((size_t)nnz_off));
    PFEM_TRY(h->bval.alloc((size_t)nnz_off));
    PFEM_TRY(h->brow_ids.alloc((size_t)n_brows + 1));
    PFEM_TRY(h->brow_ptr.alloc((size_t)n_brows + 1));
    // ghost list = sorted unique off-diagonal global columns (PETSc's garray)
    Tmp<int> ghost;
    h->ghost_cols.clear();
    h->n_ghost = 0;
    if (nnz_off > 0) {
        Tmp<int> oc, oc_sorted, nsel;
        PFEM_TRY(oc.alloc(h, 13, (size_t)nnz_off));
        PFEM_TRY(oc_sorted.alloc(h, 14, (size_t)nnz_off));
        PFEM_TRY(ghost.alloc(h, 15, (size_t)nnz_off));
        PFEM_TRY(nsel.alloc(h, 16, 1));
        gather_offdiag_cols_kernel<<<G, 256, 0, s>>>(nloc, h->row_lo, h->row_hi, h->rowptr.p, h->col.p, off_ptr.p, oc.p);
        h->launches++;
        size_t b1 = 0, b2 = 0;
        PFEM_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, b1, oc.p, oc_sorted.p, nnz_off, 0, 32, s));
        PFEM_CUDA(cub::DeviceSelect::Unique(nullptr, b2, oc_sorted.p, ghost.p, nsel.p, nnz_off, s));
        Tmp<char> tmp;
        PFEM_TRY(tmp.alloc(h, 17, b1 > b2 ? b1 : b2));
        PFEM_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, b1, oc.p, oc_sorted.p, nnz_off, 0, 32, s));
        PFEM_CUDA(cub::DeviceSelect::Unique(tmp.p, b2, oc_sorted.p, ghost.p, nsel.p, nnz_off, s));
        h->launches += 2;
        int ng = 0;
        PFEM_CUDA(cudaMemcpyAsync(&ng, nsel.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaStreamSynchronize(s));
        h->n_ghost = ng;
        h->ghost_cols.resize(ng);
        PFEM_CUDA(cudaMemcpy(h->ghost_cols.data(), ghost.p, (size_t)ng * sizeof(int), cudaMemcpyDeviceToHost));
    } else {
        PFEM_TRY(ghost.alloc(h, 15, 1));
    }
    fill_structures_kernel<<<G, 256, 0, s>>>(nloc, nrows_padded, h->row_lo, h->row_hi, h->rowptr.p, h->col.p,
                                             h->A.slice_off.p, off_ptr.p, ghost.p, h->n_ghost, h->A.col.p, h->bcol.p,
                                             h->csr2sell.p);
    boundary_rows_kernel<<<G, 256, 0, s>>>(nloc, noff.p, brow_rank.p, off_ptr.p, h->brow_ids.p, h->brow_ptr.p, n_brows, nnz_off);
    h->launches += 2;
    PFEM_CUDA(cudaGetLastError());
    // vectors (padded to whole slices so the SpMV tail needs no guards on reads)
    const size_t nv = (size_t)nrows_padded + 32;
    PFEM_TRY(h->x.alloc(nv)); PFEM_TRY(h->r.alloc(nv)); PFEM_TRY(h->z.alloc(nv));
    PFEM_TRY(h->p.alloc(nv)); PFEM_TRY(h->w.alloc(nv)); PFEM_TRY(h->dinv.alloc(nv)); PFEM_TRY(h->sv.alloc(nv));
    PFEM_CUDA(cudaMemsetAsync(h->x.p, 0, nv * sizeof(double), s));
    PFEM_CUDA(cudaMemsetAsync(h->p.p, 0, nv * sizeof(double), s));
    PFEM_CUDA(cudaMemsetAsync(h->w.p, 0, nv * sizeof(double), s));
    if (h->nranks > 1) {
        // peers still map the previous ghost buffer: everybody unmaps before anybody frees
        comm_p2p_teardown(h, false);
        std::vector<int> sync;
        PFEM_TRY(comm_allgather_int(h, 0, sync));
    }
    // [0, n_ghost]: plain ghost values (NCCL / flag-synchronised halo); from ghost_tag_off on: 16-byte {value, tag} entries
    // of the tag-validated halo (persistent kernel).  One allocation = one IPC handle.
    h->ghost_tag_off = ((size_t)h->n_ghost + 2) & ~(size_t)1;
    PFEM_TRY(h->ghost_buf.alloc(h->ghost_tag_off + 2 * ((size_t)h->n_ghost + 1)));
    PFEM_CUDA(cudaMemsetAsync(h->ghost_buf.p, 0, h->ghost_buf.n * sizeof(double), h->stream));
    PFEM_TRY(h->partials.alloc((size_t)4 * h->sm_count * 16));
    if (!h->cg.p) {
        PFEM_TRY(h->cg.alloc(1));
        PFEM_CUDA(cudaMemsetAsync(h->cg.p, 0, sizeof(CgState), s));
    }
    if (!h->cg_host) PFEM_CUDA(cudaMallocHost((void **)&h->cg_host, sizeof(CgState)));
    // halo plan: ghosts are sorted by global id, hence grouped by owner rank in ascending order
    const int P = h->nranks;
    h->send_counts.assign(P, 0); h->recv_counts.assign(P, 0);
    h->send_displs.assign(P + 1, 0); h->recv_displs.assign(P + 1, 0);
    if (P > 1) {
        std::vector<int> need_counts(P, 0);
        int q = 0;
        for (int g : h->ghost_cols) {
            while (g >= h->row_starts[q + 1]) q++;
            need_counts[q]++;
        }
        std::vector<int> asked, asked_counts;
        PFEM_TRY(comm_alltoallv_int(h, h->ghost_cols, need_counts, asked, asked_counts));
        h->recv_counts = need_counts;        // what I receive during a halo exchange = my ghosts
        h->send_counts = asked_counts;       // what I send = rows the others asked for
        for (int i = 0; i < P; i++) {
            h->send_displs[i + 1] = h->send_displs[i] + h->send_counts[i];
            h->recv_displs[i + 1] = h->recv_displs[i] + h->recv_counts[i];
        }
        std::vector<int> sidx(asked.size());
        for (size_t i = 0; i < asked.size(); i++) {
            sidx[i] = asked[i] - h->row_lo;
            if (sidx[i] < 0 || sidx[i] >= nloc) { set_error("halo plan: rank asked for a row this rank does not own"); return PFEM_ERR_NUMBERING; }
        }
        PFEM_TRY(h->send_idx.alloc(sidx.size() + 1));
        PFEM_TRY(h->send_buf.alloc(sidx.size() + 1));
        if (!sidx.empty()) PFEM_CUDA(cudaMemcpy(h->send_idx.p, sidx.data(), sidx.size() * sizeof(int), cudaMemcpyHostToDevice));
        PFEM_TRY(comm_p2p_setup(h));
    }
    PFEM_CUDA(cudaStreamSynchronize(s));
    return PFEM_OK;
}

// ---- solve-time kernels ----------------------------------------------------------------------------------------------

// CSR values -> solver storage (SELL diagonal block / off-diagonal CSR)
__global__ void values_to_solver_kernel(long long nnz, const double *__restrict__ val, const int *__restrict__ csr2sell,
                                        double *__restrict__ sell_val, double *__restrict__ bval)
{
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x) {
        const int d = csr2sell[k];
        const double v = val[k];
        if (d >= 0) sell_val[d] = v; else bval[-d - 1] = v;
    }
}

// PCSetUp_Jacobi (reciprocal diagonal, 0 -> 1) fused with the CG start: x = 0, r = b, z = M^-1 r, (z.z, z.r)
__global__ void __launch_bounds__(CG_THREADS)
cg_setup_kernel(int nloc, int row_lo, int pc_type, const int *__restrict__ rowptr, const int *__restrict__ col,
                const double *__restrict__ val, const double *__restrict__ b, double *__restrict__ x,
                double *__restrict__ r, double *__restrict__ z, double *__restrict__ dinv, double *__restrict__ partials,
                int pstride, CgState *st, int finalize /*0: publish red, 1: finalise, 2: peer all-reduce + finalise*/,
                const P2pCtx *ctx, int stage /*0: everything; ILU(0): 1 = x, r only, 2 = (z.z, z.r) of the solved z only*/)
{
    __shared__ double sh[64];
    double zz = 0.0, zr = 0.0;
    if (stage == 1) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += gridDim.x * blockDim.x) { x[i] = 0.0; r[i] = b[i]; }
        return;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nloc && stage == 2; i += gridDim.x * blockDim.x) {
        const double zi = z[i], ri = r[i];
        zz += zi * zi; zr += zi * ri;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nloc && stage == 0; i += gridDim.x * blockDim.x) {
        double d = 0.0;
        int lo = rowptr[i], hi = rowptr[i + 1];
        const int end = hi, c = row_lo + i;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (col[mid] < c) lo = mid + 1; else hi = mid;
        }
        if (lo < end && col[lo] == c) d = val[lo];
        const double di = pc_type == PFEM_PC_JACOBI ? (d == 0.0 ? 1.0 : 1.0 / d) : 1.0;
        const double ri = b[i], zi = ri * di;
        dinv[i] = di; x[i] = 0.0; r[i] = ri; z[i] = zi;
        zz += zi * zi; zr += zi * ri;
    }
    zz = block_sum(zz, sh);
    zr = block_sum(zr, sh);
    if (threadIdx.x == 0) { partials[blockIdx.x] = zz; partials[pstride + blockIdx.x] = zr; }
    if (last_block(&st->ticket[0])) {
        double a = reduce_partials(partials, gridDim.x, sh);
        double c = reduce_partials(partials + pstride, gridDim.x, sh);
        if (finalize == 2) {
            __syncthreads();
            if (threadIdx.x < 32) {
                a = __shfl_sync(0xffffffffu, a, 0); c = __shfl_sync(0xffffffffu, c, 0);
                p2p_allreduce2(ctx, st, 1, (st->seq << 32), a, c, sh);
            }
        }
        if (threadIdx.x == 0) {
            st->ticket[0] = 0;
            if (finalize) step_after_setup(st, a, c);
            else { st->red[0] = a; st->red[1] = c; }
        }
    }
}

// p = z + b p   (VecAYPX; p = z on the first iteration)
__global__ void __launch_bounds__(CG_THREADS)
cg_direction_kernel(int n, const double *__restrict__ z, double *__restrict__ p, const CgState *__restrict__ st)
{
    if (st->reason != 0) return;
    const double b = st->b;
    const bool first = st->iter == 0;
    const int n2 = n >> 1;
    const double2 *z2 = reinterpret_cast<const double2 *>(z);
    double2 *p2 = reinterpret_cast<double2 *>(p);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        const double2 zv = z2[i];
        double2 pv;
        if (first) pv = zv;
        else { pv = p2[i]; pv.x = zv.x + b * pv.x; pv.y = zv.y + b * pv.y; }
        p2[i] = pv;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) p[n - 1] = first ? z[n - 1] : z[n - 1] + b * p[n - 1];
}

// w = A_diag p over SELL-32 slices (one warp per slice), fused with the local part of p.w
__global__ void __launch_bounds__(CG_THREADS)
spmv_sell_kernel(int nslices, int nloc, const long long *__restrict__ slice_off, const int *__restrict__ col,
                 const double *__restrict__ val, const double *__restrict__ p, double *__restrict__ w,
                 double *__restrict__ partials, CgState *st, int mode /*0: no dot, 1: dot + finalize, 2: dot, publish red*/)
{
    __shared__ double sh[32];
    if (st && st->reason != 0) return;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    double pw = 0.0;
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < nslices; s += warps) {
        const long long o0 = slice_off[s], o1 = slice_off[s + 1];
        const int width = (int)((o1 - o0) >> 5);
        const int *cp = col + o0 + lane;
        const double *vp = val + o0 + lane;
        double sum = 0.0;
        int k = 0;
        for (; k + 4 <= width; k += 4) {
            const int c0 = __ldcs(cp + (k + 0) * 32), c1 = __ldcs(cp + (k + 1) * 32);
            const int c2 = __ldcs(cp + (k + 2) * 32), c3 = __ldcs(cp + (k + 3) * 32);
            const double v0 = __ldcs(vp + (k + 0) * 32), v1 = __ldcs(vp + (k + 1) * 32);
            const double v2 = __ldcs(vp + (k + 2) * 32), v3 = __ldcs(vp + (k + 3) * 32);
            const double x0 = __ldg(p + c0), x1 = __ldg(p + c1), x2 = __ldg(p + c2), x3 = __ldg(p + c3);
            sum = fma(v0, x0, sum); sum = fma(v1, x1, sum); sum = fma(v2, x2, sum); sum = fma(v3, x3, sum);
        }
        for (; k < width; k++) sum = fma(__ldcs(vp + k * 32), __ldg(p + __ldcs(cp + k * 32)), sum);
        const int r = s * 32 + lane;
        if (r < nloc) {
            w[r] = sum;
            if (mode) pw = fma(__ldg(p + r), sum, pw);
        }
    }
    if (!mode) return;
    pw = block_sum(pw, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = pw;
    if (last_block(&st->ticket[1])) {
        const double t = reduce_partials(partials, gridDim.x, sh);
        if (threadIdx.x == 0) {
            st->ticket[1] = 0;
            if (mode == 1) step_after_spmv(st, t); else st->red[2] = t;
        }
    }
}

// w[brow] += B ghost (off-diagonal block, CSR over the boundary rows), plus its share of p.w; then publishes
// red[0] = local p.w (diag + offdiag parts) for the all-reduce
__global__ void __launch_bounds__(CG_THREADS)
spmv_offdiag_kernel(int n_brows, const int *__restrict__ brow_ids, const int *__restrict__ brow_ptr,
                    const int *__restrict__ bcol, const double *__restrict__ bval, const double *__restrict__ ghost,
                    const double *__restrict__ p, double *__restrict__ w, double *__restrict__ partials, CgState *st)
{
    __shared__ double sh[32];
    if (st->reason != 0) return;
    double pw = 0.0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_brows; q += gridDim.x * blockDim.x) {
        double sum = 0.0;
        for (int k = brow_ptr[q]; k < brow_ptr[q + 1]; k++) sum = fma(bval[k], ghost[bcol[k]], sum);
        const int r = brow_ids[q];
        w[r] += sum;
        pw = fma(p[r], sum, pw);
    }
    pw = block_sum(pw, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = pw;
    if (last_block(&st->ticket[2])) {
        const double t = reduce_partials(partials, gridDim.x, sh);
        if (threadIdx.x == 0) { st->ticket[2] = 0; st->red[0] = st->red[2] + t; }
    }
}


// Peer-memory halo: write this rank's boundary values of p straight into the owners' ghost buffers over NVLink, then
// (last CTA) raise this rank's flag in every neighbour's exchange area.  Replaces pack + ncclSend/ncclRecv.
__global__ void __launch_bounds__(CG_THREADS)
halo_push_kernel(int n, const int *__restrict__ idx, const double *__restrict__ p, double *const *__restrict__ dst,
                 CgState *st, const P2pCtx *ctx)
{
    if (st->reason != 0) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) st_relaxed_sys_f64(dst[i], p[idx[i]]);
    __threadfence_system();
    if (last_block(&st->ticket2[0])) {
        __threadfence_system();
        const unsigned long long tag = p2p_tag(st, 1);
        if (threadIdx.x < ctx->nranks && ctx->sends_to[threadIdx.x]) st_release_sys(&ctx->mail[threadIdx.x]->halo_flag[ctx->rank], tag);
        if (threadIdx.x == 0) st->ticket2[0] = 0;
    }
}

// w[brow] += B ghost once every neighbour's values have landed; the last CTA then all-reduces p.w through the peers'
// mailboxes and executes the scalar step.  Replaces spmv_offdiag_kernel + ncclAllReduce + scalar kernel.
__global__ void __launch_bounds__(CG_THREADS)
spmv_offdiag_p2p_kernel(int n_brows, const int *__restrict__ brow_ids, const int *__restrict__ brow_ptr,
                        const int *__restrict__ bcol, const double *__restrict__ bval, const double *ghost,
                        const double *__restrict__ p, double *__restrict__ w, double *__restrict__ partials, CgState *st,
                        const P2pCtx *ctx)
{
    __shared__ double sh[64];
    __shared__ int halo_ok;
    if (st->reason != 0) return;
    if (threadIdx.x == 0) halo_ok = 1;
    __syncthreads();
    if (threadIdx.x < ctx->nranks && ctx->recvs_from[threadIdx.x]) {
        if (!p2p_wait(&ctx->mail[ctx->rank]->halo_flag[threadIdx.x], p2p_tag(st, 1))) halo_ok = 0;
    }
    __syncthreads();
    double pw = 0.0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_brows; q += gridDim.x * blockDim.x) {
        double sum = 0.0;
        for (int k = brow_ptr[q]; k < brow_ptr[q + 1]; k++) sum = fma(bval[k], __ldcg(ghost + bcol[k]), sum);
        const int r = brow_ids[q];
        w[r] += sum;
        pw = fma(p[r], sum, pw);
    }
    pw = block_sum(pw, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = pw;
    const int bad = halo_ok ? 0 : 1;
    if (last_block(&st->ticket[2])) {
        double t = reduce_partials(partials, gridDim.x, sh);
        double dummy = 0.0;
        __syncthreads();
        if (threadIdx.x < 32) {
            t = __shfl_sync(0xffffffffu, t, 0) + st->red[2];
            p2p_allreduce2(ctx, st, 0, p2p_tag(st, 2), t, dummy, sh);
        }
        if (threadIdx.x == 0) {
            st->ticket[2] = 0;
            if (bad) st->reason = -101;
            if (st->reason == 0) step_after_spmv(st, t);
        }
    } else if (bad && threadIdx.x == 0) {
        st->reason = -101;
    }
}

__global__ void pack_halo_kernel(int n, const int *__restrict__ idx, const double *__restrict__ p, double *__restrict__ buf,
                                 const CgState *__restrict__ st)
{
    if (st->reason != 0) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) buf[i] = p[idx[i]];
}

// x += a p ; r -= a w ; z = M^-1 r ; partial (z.z, z.r)      (VecAXPY x2, PCApply_Jacobi, VecNorm, VecDot)
__global__ void __launch_bounds__(CG_THREADS)
cg_update_kernel(int n, const double *__restrict__ p, const double *__restrict__ w, const double *__restrict__ dinv,
                 double *__restrict__ x, double *__restrict__ r, double *__restrict__ z, double *__restrict__ partials,
                 int pstride, CgState *st, int finalize, const P2pCtx *ctx, int stage /*as cg_setup_kernel*/)
{
    __shared__ double sh[64];
    if (st->reason != 0) return;
    const double a = st->a;
    double zz = 0.0, zr = 0.0;
    if (stage == 1) {          // x += a p ; r -= a w   (the ILU solves produce z)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            x[i] = fma(a, p[i], x[i]);
            r[i] = fma(-a, w[i], r[i]);
        }
        return;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n && stage == 2; i += gridDim.x * blockDim.x) {
        const double zv = z[i], rv = r[i];
        zz = fma(zv, zv, zz); zr = fma(zv, rv, zr);
    }
    const int n2 = stage == 2 ? 0 : n >> 1;
    const double2 *p2 = reinterpret_cast<const double2 *>(p), *w2 = reinterpret_cast<const double2 *>(w);
    const double2 *d2 = reinterpret_cast<const double2 *>(dinv);
    double2 *x2 = reinterpret_cast<double2 *>(x), *r2 = reinterpret_cast<double2 *>(r), *z2 = reinterpret_cast<double2 *>(z);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        const double2 pv = p2[i], wv = w2[i], dv = d2[i];
        double2 xv = x2[i], rv = r2[i], zv;
        xv.x = fma(a, pv.x, xv.x); xv.y = fma(a, pv.y, xv.y);
        rv.x = fma(-a, wv.x, rv.x); rv.y = fma(-a, wv.y, rv.y);
        zv.x = rv.x * dv.x; zv.y = rv.y * dv.y;
        x2[i] = xv; r2[i] = rv; z2[i] = zv;
        zz = fma(zv.x, zv.x, zz); zz = fma(zv.y, zv.y, zz);
        zr = fma(zv.x, rv.x, zr); zr = fma(zv.y, rv.y, zr);
    }
    if (stage == 0 && (n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int i = n - 1;
        const double xv = fma(a, p[i], x[i]), rv = fma(-a, w[i], r[i]), zv = rv * dinv[i];
        x[i] = xv; r[i] = rv; z[i] = zv;
        zz = fma(zv, zv, zz); zr = fma(zv, rv, zr);
    }
    zz = block_sum(zz, sh);
    zr = block_sum(zr, sh);
    if (threadIdx.x == 0) { partials[blockIdx.x] = zz; partials[pstride + blockIdx.x] = zr; }
    if (last_block(&st->ticket[3])) {
        double s0 = reduce_partials(partials, gridDim.x, sh);
        double s1 = reduce_partials(partials + pstride, gridDim.x, sh);
        if (finalize == 2) {
            __syncthreads();
            if (threadIdx.x < 32) {
                s0 = __shfl_sync(0xffffffffu, s0, 0); s1 = __shfl_sync(0xffffffffu, s1, 0);
                p2p_allreduce2(ctx, st, 1, p2p_tag(st, 3), s0, s1, sh);
            }
        }
        if (threadIdx.x == 0) {
            st->ticket[3] = 0;
            if (finalize) step_after_update(st, s0, s1);
            else { st->red[0] = s0; st->red[1] = s1; }
        }
    }
}


// =====================================================================================================================
// PCBJACOBI + ILU(0): the reference's DEFAULT preconditioner (solverpetsc.F:206 PCSetType(PCBJACOBI); PETSc's default
// sub-PC is ILU(0) in natural ordering, one block per rank = this rank's diagonal block).  Restated from PETSc 3.6
// MatLUFactorNumeric_SeqAIJ / MatSolve_SeqAIJ (not in the tree); the CPU checker restates the same two routines.
//
// All three kernels are "synchronisation-free" (no level sets, no grid barriers): one warp owns one row, rows are handed
// out in dependency order through a ticket taken when a CTA starts running -- so every row a warp waits for belongs to a
// CTA that is already running or finished -- and a warp spins on exactly the rows it reads.
//   factor : IKJ over the row's lower entries in ascending column order; the k-th update needs row k's U part (flag per
//            row, release/acquire); the lanes update the row's remaining entries (binary search in the sorted row).
//   solves : every value is published as ONE 16-byte {value, tag} store and consumers spin on the tagged value itself
//            (no fence, no flag): a dependent hop costs one L2 store + one L2 load.
// The arithmetic is the sequential algorithm's, operation for operation (explicit __dmul_rn/__dsub_rn: no FMA
// contraction, ascending column order) => factor and preconditioned residuals are bit-identical to the oracle whatever
// the schedule, and run-to-run deterministic.
// =====================================================================================================================

struct IluTagged { double v; unsigned long long tag; };

__device__ __forceinline__ void st_tagged_gpu(IluTagged *p, double v, unsigned long long tag)
{
    asm volatile("st.global.relaxed.gpu.v2.b64 [%0], {%1, %2};" :: "l"(p), "l"(__double_as_longlong(v)), "l"(tag) : "memory");
}
__device__ __forceinline__ double ld_tagged_gpu(const IluTagged *p, unsigned long long tag, int *fail)
{
    long long bits;
    unsigned long long tg;
    asm volatile("ld.global.relaxed.gpu.v2.b64 {%0, %1}, [%2];" : "=l"(bits), "=l"(tg) : "l"(p) : "memory");
    if (tg != tag) {
        const long long t0 = clock64();
        do {
            if (clock64() - t0 > 20000000000LL) { *fail = 1; break; }      // ~10 s: never hang the GPU
            asm volatile("ld.global.relaxed.gpu.v2.b64 {%0, %1}, [%2];" : "=l"(bits), "=l"(tg) : "l"(p) : "memory");
        } while (tg != tag);
    }
    return __longlong_as_double(bits);
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.global.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned int *p, unsigned int v)
{
    asm volatile("st.global.release.gpu.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// per row: first slot of the diagonal block, slot of the diagonal, end of the diagonal block (columns are sorted)
__global__ void ilu_rows_kernel(int nloc, int row_lo, int row_hi, const int *__restrict__ rowptr, const int *__restrict__ col,
                                int *__restrict__ dlo, int *__restrict__ ddiag, int *__restrict__ dhi, int *__restrict__ bad)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += gridDim.x * blockDim.x) {
        const int a = rowptr[i], b = rowptr[i + 1];
        auto lower = [&](int key) {
            int lo = a, hi = b;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (col[mid] < key) lo = mid + 1; else hi = mid;
            }
            return lo;
        };
        const int l = lower(row_lo), d = lower(row_lo + i), h = lower(row_hi);
        dlo[i] = l; dhi[i] = h;
        ddiag[i] = (d < b && col[d] == row_lo + i) ? d : -1;
        if (ddiag[i] < 0) atomicAdd(bad, 1);          // PETSc: "Matrix is missing diagonal entry"
    }
}

static constexpr int ILU_WARPS = 8;
// rows per warp and ticket.  1 is the measured best: 8 rows per warp widens the window of slots in flight to ~25 levels and the
// solves get 2.4x slower (tet100: 4.2 -> 10.2 ms per iteration); the solves are bound by #levels x per-level latency (~4 us)
static constexpr int ILU_RPW = 1, ILU_CHUNK = ILU_WARPS * ILU_RPW;
static int grid_for(pfem_solver *h, long long work_items, int per_thread);

// numeric ILU(0) of the diagonal block, in place on fval (a copy of the CSR values); invd = inverted pivots
__global__ void __launch_bounds__(ILU_WARPS * 32)
ilu_factor_kernel(int nloc, int row_lo, const int *__restrict__ col, const int *__restrict__ dlo, const int *__restrict__ ddiag,
                  const int *__restrict__ dhi, double *fval, double *invd, unsigned int *ready, unsigned int epoch,
                  unsigned long long *ticket, unsigned long long ticket_base, CgState *st, const int *__restrict__ order)
{
    __shared__ unsigned long long s_blk;
    if (threadIdx.x == 0) s_blk = atomicAdd(ticket, 1ULL) - ticket_base;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // slots of one chunk are interleaved over the warps: a dependency has a lower slot => another warp, or this warp earlier
    for (int k = 0; k < ILU_RPW; k++) {
    const int slot = (int)s_blk * ILU_CHUNK + k * ILU_WARPS + (threadIdx.x >> 5);
    if (slot >= nloc) return;
    const int i = order[slot];             // rows in level order: a row's dependencies sit earlier in the ticket order
    const int q0 = dlo[i], qd = ddiag[i], q1 = dhi[i];
    int fail = 0;
    for (int q = q0; q < qd; q++) {
        const int k = col[q] - row_lo;
        if (lane == 0) {
            const long long t0 = clock64();
            while (ld_acquire_gpu_u32(ready + k) != epoch)
                if (clock64() - t0 > 20000000000LL) { fail = 1; break; }
        }
        __syncwarp();
        const double m = __dmul_rn(__ldcg(fval + q), __ldcg(invd + k));
        __syncwarp();
        if (lane == 0) __stcg(fval + q, m);
        // U(k): columns > k of row k inside the block; each lands on a distinct entry of row i (or on none)
        const int u0 = ddiag[k] + 1, u1 = dhi[k];
        for (int t = u0 + lane; t < u1; t += 32) {
            const int j = col[t];
            int lo = q + 1, hi = q1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (col[mid] < j) lo = mid + 1; else hi = mid;
            }
            if (lo < q1 && col[lo] == j) __stcg(fval + lo, __dsub_rn(__ldcg(fval + lo), __dmul_rn(m, __ldcg(fval + t))));
        }
        __syncwarp();
    }
    if (lane == 0) {
        const double piv = __ldcg(fval + qd);
        if (piv == 0.0 || fail) st->reason = fail ? -101 : PFEM_DIVERGED_PCSETUP_FAILED;     // zero pivot (no shift, PETSc default)
        __stcg(invd + i, 1.0 / piv);
        __threadfence();
        st_release_gpu_u32(ready + i, epoch);
    }
    __syncwarp();
    }
}

// forward substitution  y_i = r_i - sum_{j in L(i)} l_ij y_j   (ascending j)
__global__ void __launch_bounds__(ILU_WARPS * 32)
ilu_lower_kernel(int nloc, int row_lo, const int *__restrict__ col, const int *__restrict__ dlo, const int *__restrict__ ddiag,
                 const double *__restrict__ fval, const double *__restrict__ r, IluTagged *y, unsigned long long tag,
                 unsigned long long *ticket, unsigned long long ticket_base, CgState *st, const int *__restrict__ order)
{
    __shared__ unsigned long long s_blk;
    if (threadIdx.x == 0) s_blk = atomicAdd(ticket, 1ULL) - ticket_base;
    __syncthreads();
    if (st->reason != 0) return;
    const int lane = threadIdx.x & 31;
    for (int k = 0; k < ILU_RPW; k++) {
    const int slot = (int)s_blk * ILU_CHUNK + k * ILU_WARPS + (threadIdx.x >> 5);
    if (slot >= nloc) return;
    const int i = order[slot];
    const int q0 = dlo[i], qd = ddiag[i];
    double sum = r[i];
    int fail = 0;
    for (int qb = q0; qb < qd; qb += 32) {
        const int q = qb + lane;
        double prod = 0.0;
        if (q < qd) prod = __dmul_rn(fval[q], ld_tagged_gpu(y + (col[q] - row_lo), tag, &fail));
        const int n = min(32, qd - qb);
        for (int t = 0; t < n; t++) sum = __dsub_rn(sum, __shfl_sync(0xffffffffu, prod, t));
    }
    if (__any_sync(0xffffffffu, fail) && lane == 0) st->reason = -101;
    if (lane == 0) st_tagged_gpu(y + i, sum, tag);
    }
}

// backward substitution  z_i = (y_i - sum_{j in U(i), j > i} u_ij z_j) * (1/u_ii)   (ascending j), rows in descending order
__global__ void __launch_bounds__(ILU_WARPS * 32)
ilu_upper_kernel(int nloc, int row_lo, const int *__restrict__ col, const int *__restrict__ ddiag, const int *__restrict__ dhi,
                 const double *__restrict__ fval, const double *__restrict__ invd, const IluTagged *__restrict__ y,
                 IluTagged *zt, double *__restrict__ z, unsigned long long tag, unsigned long long *ticket,
                 unsigned long long ticket_base, CgState *st, const int *__restrict__ order)
{
    __shared__ unsigned long long s_blk;
    if (threadIdx.x == 0) s_blk = atomicAdd(ticket, 1ULL) - ticket_base;
    __syncthreads();
    if (st->reason != 0) return;
    const int lane = threadIdx.x & 31;
    for (int k = 0; k < ILU_RPW; k++) {
    const int slot = (int)s_blk * ILU_CHUNK + k * ILU_WARPS + (threadIdx.x >> 5);
    if (slot >= nloc) return;
    const int i = order[slot];             // backward levels: rows whose upper entries are all solved come first
    const int q0 = ddiag[i] + 1, q1 = dhi[i];
    double sum = y[i].v;                           // written by the forward kernel (kernel boundary)
    int fail = 0;
    for (int qb = q0; qb < q1; qb += 32) {
        const int q = qb + lane;
        double prod = 0.0;
        if (q < q1) prod = __dmul_rn(fval[q], ld_tagged_gpu(zt + (col[q] - row_lo), tag, &fail));
        const int n = min(32, q1 - qb);
        for (int t = 0; t < n; t++) sum = __dsub_rn(sum, __shfl_sync(0xffffffffu, prod, t));
    }
    if (__any_sync(0xffffffffu, fail) && lane == 0) st->reason = -101;
    if (lane == 0) {
        const double zi = __dmul_rn(sum, invd[i]);
        st_tagged_gpu(zt + i, zi, tag);
        z[i] = zi;
    }
    }
}

__global__ void ilu_iota_kernel(int n, int *__restrict__ v)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = i;
}

// ---- level schedule (symbolic, once per pattern) ----------------------------------------------------------------------
// In natural order consecutive rows depend on each other (x-neighbours), so warps taken in row order spend their time
// polling: 60 ms per triangular solve on 1 M rows.  Rows are therefore handed out in LEVEL order (level = longest
// dependency chain below the row; a stable sort by level): the rows in flight are then mutually independent and a
// dependency is almost always already there when a warp asks for it.  Levels by chaotic relaxation (monotone, converges in
// at most #levels sweeps, far fewer in practice), rows sorted by a stable radix sort.  The numerical result does not
// depend on the order (each row is summed by one warp in ascending column order).
__global__ void ilu_level_sweep_kernel(int nloc, int row_lo, const int *__restrict__ col, const int *__restrict__ lo_arr,
                                       const int *__restrict__ hi_arr, int upper, int *level, int *__restrict__ changed)
{
    bool any = false;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += gridDim.x * blockDim.x) {
        // lower: entries [dlo, ddiag)   upper: entries (ddiag, dhi)
        const int a = upper ? lo_arr[i] + 1 : lo_arr[i], b = hi_arr[i];
        int lv = 0;
        for (int q = a; q < b; q++) lv = max(lv, 1 + __ldcg(level + (col[q] - row_lo)));
        if (lv > level[i]) { level[i] = lv; any = true; }
    }
    if (any) *changed = 1;
}

static int ilu_schedule(pfem_solver *h)
{
    if (h->ilu_sched_seq == h->pattern_seq && h->ilu_order_l.p) return PFEM_OK;
    cudaStream_t s = h->stream;
    const int nloc = h->size_local;
    const size_t n1 = (size_t)(nloc > 0 ? nloc : 1);
    PFEM_TRY(h->ilu_order_l.alloc(n1)); PFEM_TRY(h->ilu_order_u.alloc(n1));
    if (nloc == 0) { h->ilu_sched_seq = h->pattern_seq; return PFEM_OK; }
    DevBuf<int> level, level_s, rows, changed;
    PFEM_TRY(level.alloc(n1)); PFEM_TRY(level_s.alloc(n1)); PFEM_TRY(rows.alloc(n1)); PFEM_TRY(changed.alloc(1));
    for (int upper = 0; upper < 2; upper++) {
        PFEM_CUDA(cudaMemsetAsync(level.p, 0, n1 * sizeof(int), s));
        int sweeps = 0;
        while (true) {
            PFEM_CUDA(cudaMemsetAsync(changed.p, 0, sizeof(int), s));
            for (int k = 0; k < 8; k++)
                ilu_level_sweep_kernel<<<grid_for(h, nloc, 1), CG_THREADS, 0, s>>>(nloc, h->row_lo, h->col.p, upper ? h->ilu_ddiag.p : h->ilu_dlo.p,
                                                                                 upper ? h->ilu_dhi.p : h->ilu_ddiag.p, upper, level.p, changed.p);
            h->launches += 8;
            sweeps += 8;
            int ch = 0;
            PFEM_CUDA(cudaMemcpyAsync(&ch, changed.p, sizeof(int), cudaMemcpyDeviceToHost, s));
            PFEM_CUDA(cudaStreamSynchronize(s));
            if (!ch) break;
            if (sweeps > nloc + 16) { set_error("ILU(0): level schedule did not converge"); return PFEM_ERR_STATE; }
        }
        // stable sort of the rows by level
        ilu_iota_kernel<<<grid_for(h, nloc, 1), CG_THREADS, 0, s>>>(nloc, rows.p);
        size_t bytes = 0;
        int *out = upper ? h->ilu_order_u.p : h->ilu_order_l.p;
        PFEM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, level.p, level_s.p, rows.p, out, nloc, 0, 32, s));
        DevBuf<char> tmp;
        PFEM_TRY(tmp.alloc(bytes));
        PFEM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, level.p, level_s.p, rows.p, out, nloc, 0, 32, s));
        int maxlev = 0;
        PFEM_CUDA(cudaMemcpyAsync(&maxlev, level_s.p + nloc - 1, sizeof(int), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaStreamSynchronize(s));
        (upper ? h->ilu_levels_u : h->ilu_levels_l) = maxlev + 1;
        h->launches += 2;
    }
    h->ilu_sched_seq = h->pattern_seq;
    return PFEM_OK;
}

// z = M^-1 r with M = ILU(0) of the diagonal block: two launches
static int ilu_apply(pfem_solver *h)
{
    cudaStream_t s = h->stream;
    const int nloc = h->size_local;
    const unsigned int nblk = (unsigned int)((nloc + ILU_CHUNK - 1) / ILU_CHUNK);
    if (nblk == 0) return PFEM_OK;
    IluTagged *y = reinterpret_cast<IluTagged *>(h->ilu_y.p), *zt = reinterpret_cast<IluTagged *>(h->ilu_z.p);
    unsigned long long *ticket = reinterpret_cast<unsigned long long *>(h->ilu_ticket.p);
    const unsigned long long tag = ++h->ilu_tag;
    ilu_lower_kernel<<<nblk, ILU_WARPS * 32, 0, s>>>(nloc, h->row_lo, h->col.p, h->ilu_dlo.p, h->ilu_ddiag.p, h->ilu_fval.p, h->r.p, y, tag,
                                                     ticket, h->ilu_tickets, h->cg.p, h->ilu_order_l.p);
    h->ilu_tickets += nblk;
    ilu_upper_kernel<<<nblk, ILU_WARPS * 32, 0, s>>>(nloc, h->row_lo, h->col.p, h->ilu_ddiag.p, h->ilu_dhi.p, h->ilu_fval.p, h->ilu_invd.p, y, zt,
                                                     h->z.p, tag, ticket, h->ilu_tickets, h->cg.p, h->ilu_order_u.p);
    h->ilu_tickets += nblk;
    h->launches += 2;
    return PFEM_OK;
}

// PCSetUp: diagonal-block row ranges + numeric factorisation
static int ilu_setup(pfem_solver *h)
{
    cudaStream_t s = h->stream;
    const int nloc = h->size_local;
    const size_t n1 = (size_t)(nloc > 0 ? nloc : 1);
    PFEM_TRY(h->ilu_dlo.alloc(n1)); PFEM_TRY(h->ilu_ddiag.alloc(n1)); PFEM_TRY(h->ilu_dhi.alloc(n1));
    PFEM_TRY(h->ilu_fval.alloc((size_t)(h->nnz > 0 ? h->nnz : 1)));
    PFEM_TRY(h->ilu_invd.alloc(n1));
    if (!h->ilu_ticket.p) {
        PFEM_TRY(h->ilu_ticket.alloc(2));
        PFEM_CUDA(cudaMemsetAsync(h->ilu_ticket.p, 0, 2 * sizeof(double), s));
        h->ilu_tickets = 0; h->ilu_tag = 0; h->ilu_epoch = 0;
    }
    if (h->ilu_y.n < 2 * n1 || h->ilu_ready.n < n1) {      // tags / epochs never repeat on a handle: cleared only when (re)allocated
        PFEM_TRY(h->ilu_y.alloc(2 * n1)); PFEM_TRY(h->ilu_z.alloc(2 * n1)); PFEM_TRY(h->ilu_ready.alloc(n1));
        PFEM_CUDA(cudaMemsetAsync(h->ilu_y.p, 0, h->ilu_y.n * sizeof(double), s));
        PFEM_CUDA(cudaMemsetAsync(h->ilu_z.p, 0, h->ilu_z.n * sizeof(double), s));
        PFEM_CUDA(cudaMemsetAsync(h->ilu_ready.p, 0, h->ilu_ready.n * sizeof(unsigned int), s));
    }
    if (nloc == 0) return PFEM_OK;
    int *bad = reinterpret_cast<int *>(h->ilu_ticket.p + 1);
    PFEM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
    ilu_rows_kernel<<<grid_for(h, nloc, 1), CG_THREADS, 0, s>>>(nloc, h->row_lo, h->row_hi, h->rowptr.p, h->col.p, h->ilu_dlo.p, h->ilu_ddiag.p,
                                                              h->ilu_dhi.p, bad);
    int nbad = 0;
    PFEM_CUDA(cudaMemcpyAsync(&nbad, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaMemcpyAsync(h->ilu_fval.p, h->val.p, (size_t)h->nnz * sizeof(double), cudaMemcpyDeviceToDevice, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    if (nbad) { set_error("ILU(0): %d rows of the diagonal block have no diagonal entry", nbad); return PFEM_ERR_STATE; }
    PFEM_TRY(ilu_schedule(h));           // symbolic part: once per pattern
    const unsigned int nblk = (unsigned int)((nloc + ILU_CHUNK - 1) / ILU_CHUNK);
    ilu_factor_kernel<<<nblk, ILU_WARPS * 32, 0, s>>>(nloc, h->row_lo, h->col.p, h->ilu_dlo.p, h->ilu_ddiag.p, h->ilu_dhi.p, h->ilu_fval.p,
                                                      h->ilu_invd.p, h->ilu_ready.p, ++h->ilu_epoch,
                                                      reinterpret_cast<unsigned long long *>(h->ilu_ticket.p), h->ilu_tickets, h->cg.p,
                                                      h->ilu_order_l.p);
    h->ilu_tickets += nblk;
    h->launches += 2;
    return PFEM_OK;
}

// =====================================================================================================================
// Persistent CG: the whole KSPSolve in ONE cooperative kernel per GPU.
//
// Kernel boundaries cost ~9 us each on this part (launch gap + last-CTA reduction tail), i.e. ~30 us of a 390 us
// iteration on one GPU and most of a 50 us iteration on eight.  Here every CTA stays resident for the entire solve;
// the three phases of an iteration (direction | SpMV + p.w | update + z.z, z.r) are separated by grid barriers, CTA 0
// finishes each reduction in a fixed order (and, for nranks > 1, all-reduces it through the peers' NVLink mailboxes)
// and republishes it through a flag in local memory; the CG scalars live in registers, identically in every CTA.
// For nranks > 1 the halo is pushed into the neighbours' ghost buffers right after the direction phase; SpMV warps
// whose slice has off-diagonal entries wait for the neighbours' flags, all others never wait.
// =====================================================================================================================

struct PcgArgs {
    int nloc, nslices, row_lo, pc_type, has_off, multi, pstride, n_send;
    const long long *slice_off; const int *scol; const double *sval;
    const int *off_ptr; const int *bcol; const double *bval; const double *ghost;
    const int *rowptr; const int *col; const double *val;
    const double *b;
    double *x, *r, *z, *p, *w, *dinv;
    double *partials;
    CgState *st;
    const P2pCtx *ctx;
    const int *send_idx; double *const *send_dst;
    const double *ghost_t; double *const *send_dst_t; int halo_tag;   // tag-validated halo: 16-byte {value, tag} ghost entries
    double *bcast;                         // 4 PcgSlot {value, epoch}: [0..2] reduction results, [3] ok flag / release
    unsigned long long *arrive;            // monotonically increasing CTA arrival counter of the barriers
    unsigned int *push_ticket;             // monotonically increasing CTA arrival counter of the halo pushes
    double *parts;                         // pcg_sync_ctr: replicated per-CTA partial records [2][PCG_REPL][pstride][4]
};
static constexpr int PCG_REPL = 16;

__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.global.release.gpu.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.global.acquire.gpu.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Grid-wide reduce-and-broadcast barrier.  Every CTA deposits its NV block sums and arrives on a monotonically increasing
// counter; the LAST CTA to arrive sums all partials in a fixed order, all-reduces them across the ranks if needed
// (one 16-byte NVLink store per value and peer), publishes the result and releases everybody through one flag.
// One synchronisation per phase instead of "grid barrier, then reduce, then broadcast".  NV = 0: plain barrier.
struct PcgSlot { double value; unsigned long long tag; };     // published with ONE 16-byte store: the tag validates the value

__device__ __forceinline__ void st_slot_gpu(PcgSlot *p, double v, unsigned long long tag)
{
    asm volatile("st.global.relaxed.gpu.v2.b64 [%0], {%1, %2};" :: "l"(p), "l"(__double_as_longlong(v)), "l"(tag) : "memory");
}
__device__ __forceinline__ void ld_slot_gpu(const PcgSlot *p, double &v, unsigned long long &tag)
{
    long long bits;
    asm volatile("ld.global.relaxed.gpu.v2.b64 {%0, %1}, [%2];" : "=l"(bits), "=l"(tag) : "l"(p) : "memory");
    v = __longlong_as_double(bits);
}
__device__ __forceinline__ unsigned long long atom_add_release_gpu(unsigned long long *p, unsigned long long v)
{
    unsigned long long old;
    asm volatile("atom.add.release.gpu.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
    return old;
}

template <int NV>
__device__ __forceinline__ bool pcg_sync(const PcgArgs &a, double (&v)[NV > 0 ? NV : 1], int phase, unsigned long long ptag,
                                         unsigned long long &epoch, double *sh, double *s_bc)
{
    __shared__ int s_last;
    PcgSlot *slots = reinterpret_cast<PcgSlot *>(a.bcast);     // [0..2] values, [3] ok/plain-barrier slot
    epoch++;
    __syncthreads();                       // every thread's global writes of this phase precede thread 0's release
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) a.partials[i * a.pstride + blockIdx.x] = v[i];
        const unsigned long long t = atom_add_release_gpu(a.arrive, 1ULL);
        s_last = (t == (unsigned long long)gridDim.x * epoch - 1ULL) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();                   // acquire side of the arrivals
        double r[NV > 0 ? NV : 1];
        bool ok = true;
        if (NV > 0) {
#pragma unroll
            for (int i = 0; i < NV; i++) r[i] = reduce_partials(a.partials + i * a.pstride, gridDim.x, sh);
            __syncthreads();
            if (threadIdx.x < 32) {
#pragma unroll
                for (int i = 0; i < NV; i++) r[i] = __shfl_sync(0xffffffffu, r[i], 0);
                if (a.multi) ok = p2p_allreduce<(NV > 0 ? NV : 1)>(a.ctx, nullptr, phase, ptag, r, sh);
            }
        }
        if (threadIdx.x == 0) {
            __threadfence();               // everything the grid wrote in this phase is ordered before the publication
#pragma unroll
            for (int i = 0; i < NV; i++) { st_slot_gpu(slots + i, r[i], epoch); s_bc[i] = r[i]; }
            st_slot_gpu(slots + 3, ok ? 1.0 : 0.0, epoch);
            s_bc[3] = ok ? 1.0 : 0.0;
        }
    } else if (threadIdx.x == 0) {
        const long long c0 = clock64();
        bool ok = true;
        double val;
        unsigned long long tg;
        ld_slot_gpu(slots + 3, val, tg);
        while (tg != epoch) {              // relaxed polling: no cache maintenance inside the loop
            if (clock64() - c0 > 40000000000LL) { ok = false; break; }       // ~20 s
            __nanosleep(32);
            ld_slot_gpu(slots + 3, val, tg);
        }
        s_bc[3] = ok ? val : 0.0;
#pragma unroll
        for (int i = 0; i < NV; i++) {
            ld_slot_gpu(slots + i, val, tg);
            while (ok && tg != epoch) ld_slot_gpu(slots + i, val, tg);
            s_bc[i] = val;
        }
        __threadfence();                   // one gpu-scope fence: acquire + drop this SM's stale L1 lines
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = s_bc[i];
    const bool okall = s_bc[3] != 0.0;
    __syncthreads();                       // s_bc is reused by the next call
    return okall;
}

// Lean variant of the barrier (SYNC = 1), shaped by tools/sync_micro.cu on the B200 (profiles/r02_sync_micro.md): a CTA
// barrier of 1024 threads costs ~0.2 us, an L2 round trip ~0.2 us, a gpu-scope fence ~0.3 us, an arrival-counter grid barrier
// 1.45 us, the "last CTA reduces and publishes" scheme above 2.25 us before its two serial block reductions, and an
// all-to-all of tagged per-CTA slots 3.6 us (148 CTAs polling the same 19 lines).  So: ONE CTA barrier, then warp 0 alone
// finishes the block sums, deposits them, arrives on the counter, polls it, and then sums the per-CTA partials ITSELF
// (every CTA does, in the same fixed order: lane-strided, then the xor tree => identical in all CTAs, deterministic) --
// no second hop, no broadcast slot, no further CTA barrier until the caller's own one after the scalar step.
// For nranks > 1 CTA 0 forwards the local sums into the peers' mailboxes and warp 0 of EVERY CTA polls the local mailbox.
// v[]: per-thread partial sums on entry; grid (and cross-rank) sums in warp 0 on return (thread 0 is the only consumer).
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.global.relaxed.gpu.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_add_relaxed_gpu_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

template <int NV>
__device__ __forceinline__ bool pcg_sync_ctr(const PcgArgs &a, double (&v)[NV > 0 ? NV : 1], int phase, unsigned long long ptag,
                                             unsigned long long &epoch, double *sh2, int tb = 0)
{
    constexpr int NS = NV > 0 ? NV : 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    epoch++;
    double *sh = sh2 + (epoch & 1ULL) * 96;          // [3 values][32 warps], double-buffered by parity
    // the per-CTA partials are double-buffered by parity too: a fast CTA deposits for the NEXT barrier while a slow one is
    // still summing this one's; the same parity is rewritten two barriers later, which nobody reaches before all have read
    // Layout [parity][replica][CTA] of 32-byte records {v0, v1, v2, -}: every CTA deposits PCG_REPL copies (16 lanes, one store
    // each) and CTA c sums replica c % PCG_REPL, so a 128-byte line has ~9 readers instead of 148 (the trace showed the
    // all-CTAs-read-the-same-19-lines sum costing 1500-5600 cycles), and all NV values of a CTA come in one record.
    double *part = a.parts + (size_t)(epoch & 1ULL) * PCG_REPL * a.pstride * 4;
    if (NV > 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) { const double t = warp_sum(v[i]); if (lane == 0) sh[i * 32 + wid] = t; }
    }
    __syncthreads();                                 // the CTA's writes of this phase precede lane 0's release fence
    PCG_T(tb + 1);
    bool ok = true;
    if (wid == 0) {
        double t[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) {
            t[i] = 0.0;
            if (NV > 0) { t[i] = lane < (int)(blockDim.x >> 5) ? sh[i * 32 + lane] : 0.0; t[i] = warp_sum(t[i]); }
        }
        if (NV > 0 && lane < PCG_REPL) {
            double2 *rec = reinterpret_cast<double2 *>(part + ((size_t)lane * a.pstride + blockIdx.x) * 4);
            __stcg(rec, make_double2(t[0], NV > 1 ? t[1] : 0.0));
            if (NV > 2) __stcg(rec + 1, make_double2(t[2], 0.0));
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence();                         // release (cumulative over the warp's deposits and the CTA's phase writes)
            PCG_T(tb + 2);
            red_add_relaxed_gpu_u64(a.arrive, 1ULL);
            const unsigned long long target = (unsigned long long)gridDim.x * epoch;
            const long long c0 = clock64();
            while (ld_relaxed_gpu_u64(a.arrive) < target) {
                if (clock64() - c0 > 40000000000LL) { ok = false; break; }       // ~20 s: never hang the GPU
            }
            PCG_T(tb + 3);
            __threadfence();                         // acquire (+ drops this SM's stale L1 lines)
            PCG_T(tb + 4);
        }
        ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
        if (NV > 0) {
            double r[NS];
            {
                const double *rep = part + (size_t)(blockIdx.x % PCG_REPL) * a.pstride * 4;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
                for (int c = lane; c < (int)gridDim.x; c += 32) {
                    const double2 u = __ldcg(reinterpret_cast<const double2 *>(rep + (size_t)c * 4));
                    s0 += u.x;
                    if (NV > 1) s1 += u.y;
                    if (NV > 2) s2 += __ldcg(rep + (size_t)c * 4 + 2);
                }
                r[0] = warp_sum(s0);
                if (NV > 1) r[NV > 1 ? 1 : 0] = warp_sum(s1);
                if (NV > 2) r[NV > 2 ? 2 : 0] = warp_sum(s2);
            }
            if (a.multi) {
                // cross-rank sum: CTA 0 forwards, everybody polls the local mailbox, rank-order sum (bit-identical everywhere)
                const P2pCtx *c = a.ctx;
                const int P = c->nranks, me = c->rank;
                if (lane < P) {
                    if (blockIdx.x == 0) {
                        P2pMail *dst = c->mail[lane];
#pragma unroll
                        for (int i = 0; i < NV; i++) st_slot_sys(&dst->red[phase][me][i], r[i], ptag);
                    }
                    const P2pMail *mine = c->mail[me];
                    const long long t0 = clock64();
#pragma unroll
                    for (int i = 0; i < NV; i++) {
                        double val;
                        unsigned long long tg;
                        ld_slot_sys(&mine->red[phase][lane][i], val, tg);
                        while (ok && tg != ptag) {
                            if (clock64() - t0 > 20000000000LL) ok = false;          // ~10 s: dead peer
                            ld_slot_sys(&mine->red[phase][lane][i], val, tg);
                        }
                        r[i] = val;
                    }
                }
                ok = __all_sync(0xffffffffu, ok);
#pragma unroll
                for (int i = 0; i < NV; i++) {
                    double s0 = 0.0;
                    for (int q = 0; q < P; q++) s0 += __shfl_sync(0xffffffffu, r[i], q);
                    r[i] = s0;
                }
            }
#pragma unroll
            for (int i = 0; i < NV; i++) v[i] = r[i];
        }
    }
    PCG_T(tb + 5);
    if (NV == 0) __syncthreads();                    // plain barrier: nobody leaves before lane 0 has seen all arrivals
    return ok;                                       // (NV > 0: the caller's CTA barrier after its scalar step does that)
}

// phase-end reduction + grid barrier of the persistent kernels; v[] = per-thread partial sums on entry, grid (and cross-rank)
// sums in thread 0 on return.  SYNC 0: block sums, arrival counter, last CTA reduces and broadcasts (pcg_sync);
// SYNC 1: lean counter barrier, warp 0 reduces in every CTA (pcg_sync_ctr).
template <int NV, int SYNC>
__device__ __forceinline__ bool pcg_reduce(const PcgArgs &a, double (&v)[NV > 0 ? NV : 1], int phase, unsigned long long ptag,
                                           unsigned long long &epoch, unsigned long long seq, double *sh, double *s_bc, double *sh2,
                                           int tb = 0)
{
    if (SYNC == 1) return pcg_sync_ctr<NV>(a, v, phase, ptag, epoch, sh2, tb);
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = block_sum(v[i], sh);
    return pcg_sync<NV>(a, v, phase, ptag, epoch, sh, s_bc);
}

// FUSED: the direction phase (p = z + b p) and its grid barrier are folded into the SpMV: every gathered entry is formed
// on the fly as fma(b, p_old[c], z[c]) from the previous direction (p is double-buffered), the halo push forms its values
// the same way, and the row's own entry is written out as the new direction.  Two barriers per iteration instead of three
// at the price of a second gather stream (L2-resident once the rows are split over several GPUs).
template <int THREADS, int MINB, bool FUSED, int SYNC>
__global__ void __launch_bounds__(THREADS, MINB)
cg_persistent_kernel(const PcgArgs a, double *__restrict__ pbuf2)
{
    __shared__ double sh[64];
    __shared__ double s_bc[4];
    __shared__ double s_push;
    __shared__ double sh2[SYNC == 1 ? 192 : 1];
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, gwarp = gtid >> 5, gwarps = gthreads >> 5;
    const int nloc = a.nloc;
    unsigned long long epoch = 0;
    unsigned int pushes = 0;

    // the CG scalars: one copy per CTA in shared memory, advanced identically in every CTA by its thread 0
    __shared__ CgState ls;
    if (threadIdx.x == 0) {
        ls.rtol = a.st->rtol; ls.abstol = a.st->abstol; ls.dtol = a.st->dtol; ls.max_it = a.st->max_it; ls.seq = a.st->seq;
        ls.beta = ls.betaold = ls.dpi = ls.dpiold = ls.dp = ls.a = ls.b = ls.ttol = ls.rnorm0 = 0.0;
        ls.its = 0; ls.reason = 0; ls.iter = 0;
    }
    __syncthreads();

    // ---- set-up: PCSetUp_Jacobi, x = 0, r = b, z = M^-1 r, (z.z, z.r) ----
    {
        double zz = 0.0, zr = 0.0;
        for (int i = gtid; i < nloc; i += gthreads) {
            double d = 0.0;
            int lo = a.rowptr[i], hi = a.rowptr[i + 1];
            const int end = hi, c = a.row_lo + i;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (a.col[mid] < c) lo = mid + 1; else hi = mid;
            }
            if (lo < end && a.col[lo] == c) d = a.val[lo];
            const double di = a.pc_type == PFEM_PC_JACOBI ? (d == 0.0 ? 1.0 : 1.0 / d) : 1.0;
            const double ri = a.b[i], zi = ri * di;
            a.dinv[i] = di; a.x[i] = 0.0; a.r[i] = ri; a.z[i] = zi;
            zz += zi * zi; zr += zi * ri;
        }
        double v[2] = {zz, zr};
        const bool ok = pcg_reduce<2, SYNC>(a, v, 1, ls.seq << 32, epoch, ls.seq, sh, s_bc, sh2);
        zz = v[0]; zr = v[1];
        if (threadIdx.x == 0) {
            step_after_setup(&ls, zz, zr);
            if (!ok) ls.reason = -101;
        }
        __syncthreads();
    }

    while (ls.reason == 0) {
        PCG_T(0);
        // ---- direction: p = z + b p ----
        const bool first = ls.iter == 0;
        const double bdir = ls.b;
        double *p_new = a.p;
        const double *p_old = a.p;
        if (FUSED) { p_new = (ls.iter & 1) ? a.p : pbuf2; p_old = (ls.iter & 1) ? pbuf2 : a.p; }
        if (!FUSED) {
            const double b = ls.b;
            const int n2 = nloc >> 1;
            const double2 *z2 = reinterpret_cast<const double2 *>(a.z);
            double2 *p2 = reinterpret_cast<double2 *>(a.p);
            for (int i = gtid; i < n2; i += gthreads) {
                const double2 zv = z2[i];
                double2 pv;
                if (first) pv = zv;
                else { pv = p2[i]; pv.x = zv.x + b * pv.x; pv.y = zv.y + b * pv.y; }
                p2[i] = pv;
            }
            if ((nloc & 1) && gtid == 0) a.p[nloc - 1] = first ? a.z[nloc - 1] : a.z[nloc - 1] + b * a.p[nloc - 1];
            double none[1] = {0.0};
            PCG_T(1);
            pcg_reduce<0, SYNC>(a, none, 0, 0ULL, epoch, ls.seq, sh, s_bc, sh2, 1);
            PCG_T(7);
        }
        auto pval = [&](int c) -> double {     // entry c of the current direction
            if (!FUSED) return a.p[c];
            const double zc = a.z[c];
            return first ? zc : fma(bdir, p_old[c], zc);
        };
        // ---- halo push (nranks > 1): boundary values straight into the neighbours' ghost buffers ----
        const unsigned long long htag = (ls.seq << 32) | (unsigned long long)(4u * (unsigned int)ls.iter + 1u);
        if (a.multi && a.halo_tag == 1) {
            // The LAST warp of every CTA pushes (4 736 threads share the entries) and then collects the NVLink write
            // acknowledgements right away with a system fence of its own.  The 2-GPU stage trace (profiles/
            // r02_pcg_trace_2gpu_tet100.log) showed why: when the threads that pushed include lane 0 of warp 0, that lane's
            // release fence at the END of the phase waits one NVLink round trip for the acknowledgements of stores issued
            // 20 us earlier (6 500 instead of 2 000 cycles, on every rank, every iteration).
            if ((threadIdx.x >> 5) == (blockDim.x >> 5) - 1) {
                bool sent = false;
                for (int i = blockIdx.x * 32 + lane; i < a.n_send; i += gridDim.x * 32) {
                    st_ghost_tagged(a.send_dst_t[i], pval(a.send_idx[i]), htag);
                    sent = true;
                }
                if (sent) __threadfence_system();
            }
        } else if (a.multi && a.halo_tag) {
            for (int i = gtid; i < a.n_send; i += gthreads) st_ghost_tagged(a.send_dst_t[i], pval(a.send_idx[i]), htag);
            if (a.halo_tag == 2 && gtid < a.n_send) __threadfence_system();      // push the posted stores out now (senders only)
        } else if (a.multi) {
            for (int i = gtid; i < a.n_send; i += gthreads) st_relaxed_sys_f64(a.send_dst[i], pval(a.send_idx[i]));
            __threadfence_system();
            __syncthreads();
            pushes++;
            if (threadIdx.x == 0) {
                __threadfence_system();      // cumulative release: the CTA's ghost stores (ordered by the barrier) precede the ticket
                const unsigned int t = atomicAdd(a.push_ticket, 1u);
                s_push = (t == gridDim.x * pushes - 1u) ? 1.0 : 0.0;
            }
            __syncthreads();
            if (s_push != 0.0) {
                __threadfence_system();
                if (threadIdx.x < a.ctx->nranks && a.ctx->sends_to[threadIdx.x])
                    st_release_sys(&a.ctx->mail[threadIdx.x]->halo_flag[a.ctx->rank], htag);
            }
        }
        // ---- SpMV: w = A_diag p (+ B ghost), fused with p.w ----
        {
            double pw = 0.0;
            bool halo_ready = false, halo_ok = true;
            const int rot = a.has_off ? (a.nslices >> 1) : 0;
            for (int s0 = gwarp; s0 < a.nslices; s0 += gwarps) {
                int s = s0 + rot;
                if (s >= a.nslices) s -= a.nslices;
                const long long o0 = a.slice_off[s], o1 = a.slice_off[s + 1];
                const int width = (int)((o1 - o0) >> 5);
                const int *cp = a.scol + o0 + lane;
                const double *vp = a.sval + o0 + lane;
                double sum = 0.0;
                int k = 0;
                for (; k + 4 <= width; k += 4) {
                    const int c0 = __ldcs(cp + (k + 0) * 32), c1 = __ldcs(cp + (k + 1) * 32);
                    const int c2 = __ldcs(cp + (k + 2) * 32), c3 = __ldcs(cp + (k + 3) * 32);
                    const double v0 = __ldcs(vp + (k + 0) * 32), v1 = __ldcs(vp + (k + 1) * 32);
                    const double v2 = __ldcs(vp + (k + 2) * 32), v3 = __ldcs(vp + (k + 3) * 32);
                    const double x0 = pval(c0), x1 = pval(c1), x2 = pval(c2), x3 = pval(c3);
                    sum = fma(v0, x0, sum); sum = fma(v1, x1, sum); sum = fma(v2, x2, sum); sum = fma(v3, x3, sum);
                }
                for (; k < width; k++) sum = fma(__ldcs(vp + k * 32), pval(__ldcs(cp + k * 32)), sum);
                const int r = s * 32 + lane;
                if (a.has_off) {
                    int lo = 0, hi = 0;
                    if (r < nloc) { lo = a.off_ptr[r]; hi = a.off_ptr[r + 1]; }
                    if (a.halo_tag) {
                        if (hi > lo) sum = sum + offdiag_row_tagged(a.bval, a.bcol, a.ghost_t, lo, hi, htag, halo_ok);
                    } else if (__any_sync(0xffffffffu, hi > lo)) {
                        if (!halo_ready) {       // first boundary slice of this warp in this iteration: wait for the neighbours
                            bool okw = true;
                            if (lane < a.ctx->nranks && a.ctx->recvs_from[lane])
                                okw = p2p_wait(&a.ctx->mail[a.ctx->rank]->halo_flag[lane], htag);
                            halo_ok = __all_sync(0xffffffffu, okw);
                            halo_ready = true;
                        }
                        double osum = 0.0;
                        for (int q = lo; q < hi; q++) osum = fma(a.bval[q], __ldcg(a.ghost + a.bcol[q]), osum);
                        sum = sum + osum;
                    }
                }
                if (r < nloc) {
                    const double pr = pval(r);
                    if (FUSED) p_new[r] = pr;
                    a.w[r] = sum;
                    pw = fma(pr, sum, pw);
                }
            }
            double v[1] = {pw};
            if (!halo_ok) v[0] = __longlong_as_double(0x7ff8000000000000LL);
            PCG_T(10);
            const bool ok = pcg_reduce<1, SYNC>(a, v, 0, (ls.seq << 32) | (unsigned long long)(4u * (unsigned int)ls.iter + 2u), epoch,
                                                ls.seq, sh, s_bc, sh2, 10);
            pw = v[0];
            if (threadIdx.x == 0) {
                step_after_spmv(&ls, pw);
                if (!ok || pw != pw) ls.reason = -101;
            }
            PCG_T(16);
            __syncthreads();
            PCG_T(17);
        }
        if (ls.reason != 0) break;
        // ---- update: x += a p, r -= a w, z = M^-1 r, (z.z, z.r) ----
        {
            const double al = ls.a;
            double zz = 0.0, zr = 0.0;
            const int n2 = nloc >> 1;
            const double2 *p2 = reinterpret_cast<const double2 *>(p_new), *w2 = reinterpret_cast<const double2 *>(a.w);
            const double2 *d2 = reinterpret_cast<const double2 *>(a.dinv);
            double2 *x2 = reinterpret_cast<double2 *>(a.x), *r2 = reinterpret_cast<double2 *>(a.r), *z2 = reinterpret_cast<double2 *>(a.z);
            for (int i = gtid; i < n2; i += gthreads) {
                const double2 pv = p2[i], wv = w2[i], dv = d2[i];
                double2 xv = x2[i], rv = r2[i], zv;
                xv.x = fma(al, pv.x, xv.x); xv.y = fma(al, pv.y, xv.y);
                rv.x = fma(-al, wv.x, rv.x); rv.y = fma(-al, wv.y, rv.y);
                zv.x = rv.x * dv.x; zv.y = rv.y * dv.y;
                x2[i] = xv; r2[i] = rv; z2[i] = zv;
                zz = fma(zv.x, zv.x, zz); zz = fma(zv.y, zv.y, zz);
                zr = fma(zv.x, rv.x, zr); zr = fma(zv.y, rv.y, zr);
            }
            if ((nloc & 1) && gtid == 0) {
                const int i = nloc - 1;
                const double xv = fma(al, p_new[i], a.x[i]), rv = fma(-al, a.w[i], a.r[i]), zv = rv * a.dinv[i];
                a.x[i] = xv; a.r[i] = rv; a.z[i] = zv;
                zz = fma(zv, zv, zz); zr = fma(zv, rv, zr);
            }
            double v[2] = {zz, zr};
            PCG_T(20);
            const bool ok = pcg_reduce<2, SYNC>(a, v, 1, (ls.seq << 32) | (unsigned long long)(4u * (unsigned int)ls.iter + 3u), epoch,
                                                ls.seq, sh, s_bc, sh2, 20);
            zz = v[0]; zr = v[1];
            if (threadIdx.x == 0) {
                step_after_update(&ls, zz, zr);
                if (!ok) ls.reason = -101;
            }
            PCG_T(26);
            __syncthreads();
            PCG_T(27);
        }
    }
    if (gtid == 0) {
        a.st->its = ls.its; a.st->reason = ls.reason; a.st->iter = ls.iter; a.st->dp = ls.dp;
        a.st->beta = ls.beta; a.st->a = ls.a; a.st->b = ls.b;
    }
}


// ---- single-reduction CG (PETSc's KSPCGUseSingleReduction / -ksp_cg_single_reduction recurrences) -----------------------
// s = A z is formed instead of w = A p; w follows by recurrence (w = s + b w) and p.w by
// dpi = delta - beta^2 dpiold / betaold^2 with delta = z.s, so ONE fused reduction (z.z, z.r, z.s) per iteration remains and
// all vector updates collapse into one phase: two grid barriers and one cross-rank all-reduce per iteration.
// Same iterates in exact arithmetic; selected for nranks > 1 where synchronisation, not bandwidth, bounds the iteration.

__device__ void begin_iteration_sr(CgState *st)
{
    st->its = st->iter + 1;
    if (st->beta == 0.0) { st->reason = PFEM_CONVERGED_ATOL; return; }
    if (st->iter > 0 && st->beta * st->betaold < 0.0) { st->reason = PFEM_DIVERGED_INDEFINITE_PC; return; }
    st->dpiold = st->dpi;
    if (st->iter == 0) { st->b = 0.0; st->dpi = st->delta; }
    else {
        st->b = st->beta / st->betaold;
        st->dpi = st->delta - st->beta * st->beta * st->dpiold / (st->betaold * st->betaold);
    }
    st->betaold = st->beta;
    if (st->dpi == 0.0 || (st->iter > 0 && st->dpi * st->dpiold <= 0.0)) { st->reason = PFEM_DIVERGED_INDEFINITE_MAT; return; }
    st->a = st->beta / st->dpi;
}

__device__ void step_sr(CgState *st, bool first, double zz, double zr, double zs)
{
    st->dp = sqrt(zz);
    if (first) {
        st->its = 0; st->iter = 0; st->dpi = 0.0; st->betaold = 0.0;
        st->reason = converged_default(st, 0, st->dp);
        if (st->reason) return;
    } else {
        st->reason = converged_default(st, st->iter + 1, st->dp);
        if (st->reason) return;
        st->iter++;
        if (st->iter >= st->max_it) { st->reason = PFEM_DIVERGED_ITS; return; }
    }
    st->beta = zr;
    st->delta = zs;
    begin_iteration_sr(st);
}

template <int THREADS, int MINB, int SYNC>
__global__ void __launch_bounds__(THREADS, MINB)
cg_persistent_sr_kernel(const PcgArgs a, double *__restrict__ sv)
{
    __shared__ double sh[64];
    __shared__ double s_bc[4];
    __shared__ double s_push;
    __shared__ double sh2[SYNC == 1 ? 192 : 1];
    __shared__ CgState ls;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, gwarp = gtid >> 5, gwarps = gthreads >> 5;
    const int nloc = a.nloc;
    unsigned long long epoch = 0;
    unsigned int pushes = 0;
    if (threadIdx.x == 0) {
        ls.rtol = a.st->rtol; ls.abstol = a.st->abstol; ls.dtol = a.st->dtol; ls.max_it = a.st->max_it; ls.seq = a.st->seq;
        ls.beta = ls.betaold = ls.dpi = ls.dpiold = ls.dp = ls.a = ls.b = ls.ttol = ls.rnorm0 = ls.delta = 0.0;
        ls.its = 0; ls.reason = 0; ls.iter = 0;
    }
    __syncthreads();
    // ---- set-up vectors: PCSetUp_Jacobi, x = 0, r = b, z = M^-1 r, p = w = 0 ----
    for (int i = gtid; i < nloc; i += gthreads) {
        double d = 0.0;
        int lo = a.rowptr[i], hi = a.rowptr[i + 1];
        const int end = hi, c = a.row_lo + i;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (a.col[mid] < c) lo = mid + 1; else hi = mid;
        }
        if (lo < end && a.col[lo] == c) d = a.val[lo];
        const double di = a.pc_type == PFEM_PC_JACOBI ? (d == 0.0 ? 1.0 : 1.0 / d) : 1.0;
        const double ri = a.b[i];
        a.dinv[i] = di; a.x[i] = 0.0; a.r[i] = ri; a.z[i] = ri * di; a.p[i] = 0.0; a.w[i] = 0.0;
    }
    bool first = true;
    unsigned int round = 0;
    while (true) {
        {
            double none[1] = {0.0};
            pcg_reduce<0, SYNC>(a, none, 0, 0ULL, epoch, ls.seq, sh, s_bc, sh2);
        }
        // ---- halo push of z (nranks > 1) ----
        const unsigned long long htag = (ls.seq << 32) | (unsigned long long)(4u * round + 1u);
        if (a.multi && a.halo_tag == 1) {
            if ((threadIdx.x >> 5) == (blockDim.x >> 5) - 1) {       // see cg_persistent_kernel
                bool sent = false;
                for (int i = blockIdx.x * 32 + lane; i < a.n_send; i += gridDim.x * 32) {
                    st_ghost_tagged(a.send_dst_t[i], a.z[a.send_idx[i]], htag);
                    sent = true;
                }
                if (sent) __threadfence_system();
            }
        } else if (a.multi && a.halo_tag) {
            for (int i = gtid; i < a.n_send; i += gthreads) st_ghost_tagged(a.send_dst_t[i], a.z[a.send_idx[i]], htag);
            if (a.halo_tag == 2 && gtid < a.n_send) __threadfence_system();
        } else if (a.multi) {
            for (int i = gtid; i < a.n_send; i += gthreads) st_relaxed_sys_f64(a.send_dst[i], a.z[a.send_idx[i]]);
            __threadfence_system();
            __syncthreads();
            pushes++;
            if (threadIdx.x == 0) {
                __threadfence_system();      // cumulative release (see cg_persistent_kernel)
                const unsigned int t = atomicAdd(a.push_ticket, 1u);
                s_push = (t == gridDim.x * pushes - 1u) ? 1.0 : 0.0;
            }
            __syncthreads();
            if (s_push != 0.0) {
                __threadfence_system();
                if (threadIdx.x < a.ctx->nranks && a.ctx->sends_to[threadIdx.x])
                    st_release_sys(&a.ctx->mail[threadIdx.x]->halo_flag[a.ctx->rank], htag);
            }
        }
        // ---- s = A z, fused with (z.z, z.r, z.s) ----
        double zz = 0.0, zr = 0.0, zs = 0.0;
        bool halo_ready = false, halo_ok = true;
        const int rot = a.has_off ? (a.nslices >> 1) : 0;
        for (int s0 = gwarp; s0 < a.nslices; s0 += gwarps) {
            int s = s0 + rot;
            if (s >= a.nslices) s -= a.nslices;
            const long long o0 = a.slice_off[s], o1 = a.slice_off[s + 1];
            const int width = (int)((o1 - o0) >> 5);
            const int *cp = a.scol + o0 + lane;
            const double *vp = a.sval + o0 + lane;
            double sum = 0.0;
            int k = 0;
            for (; k + 4 <= width; k += 4) {
                const int c0 = __ldcs(cp + (k + 0) * 32), c1 = __ldcs(cp + (k + 1) * 32);
                const int c2 = __ldcs(cp + (k + 2) * 32), c3 = __ldcs(cp + (k + 3) * 32);
                const double v0 = __ldcs(vp + (k + 0) * 32), v1 = __ldcs(vp + (k + 1) * 32);
                const double v2 = __ldcs(vp + (k + 2) * 32), v3 = __ldcs(vp + (k + 3) * 32);
                const double x0 = a.z[c0], x1 = a.z[c1], x2 = a.z[c2], x3 = a.z[c3];
                sum = fma(v0, x0, sum); sum = fma(v1, x1, sum); sum = fma(v2, x2, sum); sum = fma(v3, x3, sum);
            }
            for (; k < width; k++) sum = fma(__ldcs(vp + k * 32), a.z[__ldcs(cp + k * 32)], sum);
            const int r = s * 32 + lane;
            if (a.has_off) {
                int lo = 0, hi = 0;
                if (r < nloc) { lo = a.off_ptr[r]; hi = a.off_ptr[r + 1]; }
                if (a.halo_tag) {
                    if (hi > lo) sum = sum + offdiag_row_tagged(a.bval, a.bcol, a.ghost_t, lo, hi, htag, halo_ok);
                } else if (__any_sync(0xffffffffu, hi > lo)) {
                    if (!halo_ready) {
                        bool okw = true;
                        if (lane < a.ctx->nranks && a.ctx->recvs_from[lane])
                            okw = p2p_wait(&a.ctx->mail[a.ctx->rank]->halo_flag[lane], htag);
                        halo_ok = __all_sync(0xffffffffu, okw);
                        halo_ready = true;
                    }
                    double osum = 0.0;
                    for (int q = lo; q < hi; q++) osum = fma(a.bval[q], __ldcg(a.ghost + a.bcol[q]), osum);
                    sum = sum + osum;
                }
            }
            if (r < nloc) {
                sv[r] = sum;
                const double zi = a.z[r], ri = a.r[r];
                zz = fma(zi, zi, zz); zr = fma(zi, ri, zr); zs = fma(zi, sum, zs);
            }
        }
        double v[3] = {zz, zr, zs};
        if (!halo_ok) v[0] = __longlong_as_double(0x7ff8000000000000LL);
        const bool ok = pcg_reduce<3, SYNC>(a, v, (int)(round & 1u), (ls.seq << 32) | (unsigned long long)(4u * round + 2u), epoch,
                                            ls.seq, sh, s_bc, sh2);   // mailbox parity: see p2p_allreduce
        if (threadIdx.x == 0) {
            step_sr(&ls, first, v[0], v[1], v[2]);
            if (!ok || v[0] != v[0]) ls.reason = -101;
        }
        __syncthreads();
        first = false;
        round++;
        if (ls.reason != 0) break;
        // ---- all vector updates of the iteration in one pass ----
        {
            const double al = ls.a, b = ls.b;
            const int n2 = nloc >> 1;
            const double2 *s2 = reinterpret_cast<const double2 *>(sv), *d2 = reinterpret_cast<const double2 *>(a.dinv);
            double2 *p2 = reinterpret_cast<double2 *>(a.p), *w2 = reinterpret_cast<double2 *>(a.w);
            double2 *x2 = reinterpret_cast<double2 *>(a.x), *r2 = reinterpret_cast<double2 *>(a.r), *z2 = reinterpret_cast<double2 *>(a.z);
            for (int i = gtid; i < n2; i += gthreads) {
                const double2 zv = z2[i], svv = s2[i], dv = d2[i];
                double2 pv = p2[i], wv = w2[i], xv = x2[i], rv = r2[i], zn;
                pv.x = fma(b, pv.x, zv.x); pv.y = fma(b, pv.y, zv.y);            // p = z + b p
                wv.x = fma(b, wv.x, svv.x); wv.y = fma(b, wv.y, svv.y);          // w = s + b w  (= A p)
                xv.x = fma(al, pv.x, xv.x); xv.y = fma(al, pv.y, xv.y);
                rv.x = fma(-al, wv.x, rv.x); rv.y = fma(-al, wv.y, rv.y);
                zn.x = rv.x * dv.x; zn.y = rv.y * dv.y;
                p2[i] = pv; w2[i] = wv; x2[i] = xv; r2[i] = rv; z2[i] = zn;
            }
            if ((nloc & 1) && gtid == 0) {
                const int i = nloc - 1;
                const double pv = fma(b, a.p[i], a.z[i]), wv = fma(b, a.w[i], sv[i]);
                const double xv = fma(al, pv, a.x[i]), rv = fma(-al, wv, a.r[i]);
                a.p[i] = pv; a.w[i] = wv; a.x[i] = xv; a.r[i] = rv; a.z[i] = rv * a.dinv[i];
            }
        }
    }
    if (gtid == 0) {
        a.st->its = ls.its; a.st->reason = ls.reason; a.st->iter = ls.iter; a.st->dp = ls.dp;
        a.st->beta = ls.beta; a.st->a = ls.a; a.st->b = ls.b;
    }
}

static int cg_solve_persistent(pfem_solver *h, bool &used)
{
    used = false;
    const char *env = getenv("PFEM_CG");
    if (env && strcmp(env, "kernels") == 0) return PFEM_OK;
    if (h->nranks > 1 && !h->p2p) return PFEM_OK;            // the NCCL path needs host-launched collectives
    if (h->profile) return PFEM_OK;                           // per-launch SpMV timing needs separate launches
    if (h->pc_type == PFEM_PC_BJACOBI_ILU0) return PFEM_OK;   // triangular solves run as their own (sync-free) kernels
    int coop = 0;
    PFEM_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
    if (!coop) return PFEM_OK;
    // CTA shape: many resident threads feed HBM best on big blocks (5 x 256 per SM); fewer, fatter CTAs make the grid
    // barriers cheaper when the per-GPU block is small (1 x 1024 per SM = 148 arrivals)
    const char *cfg = getenv("PFEM_PCG_CFG");
    int threads = h->size_local >= 3000000 ? 256 : 1024, minb = h->size_local >= 3000000 ? 5 : 1;
    if (cfg) sscanf(cfg, "%dx%d", &threads, &minb);
    // recurrence: PETSc's default two-reduction CG.  Its single-reduction variant (PFEM_CG_SR=1) saves one barrier and one
    // all-reduce per iteration but moves one more vector; measured slower on 1 and 2 B200s (0.228 vs 0.211 ms/iteration on
    // C5 at 2 GPUs), so it is opt-in.
    const char *srenv = getenv("PFEM_CG_SR");
    const bool sr = srenv ? (srenv[0] == '1') : false;
    const void *fn = nullptr;
    // optional: fold the direction phase into the SpMV (PFEM_PCG_FUSED=1).  Measured on 2 x B200 it is a wash (the
    // cross-GPU latency chain, not the barrier count, bounds the iteration) and on one GPU the second gather stream costs
    // 20 %, so it is off by default.
    const char *fenv = getenv("PFEM_PCG_FUSED");
    const bool fused = fenv ? (fenv[0] == '1') : false;
    // barrier flavour: "lean" (pcg_sync_ctr: one CTA barrier, warp 0 of every CTA reduces) or "last" (pcg_sync: last CTA
    // reduces and publishes).  PFEM_PCG_SYNC=lean|last overrides.
    const char *syenv = getenv("PFEM_PCG_SYNC");
    const bool lean = syenv ? (strcmp(syenv, "lean") == 0) : true;
#define PCG_PICK2(T, M, S) (sr ? (const void *)cg_persistent_sr_kernel<T, M, S> : fused ? (const void *)cg_persistent_kernel<T, M, true, S> : (const void *)cg_persistent_kernel<T, M, false, S>)
#define PCG_PICK(T, M) (lean ? PCG_PICK2(T, M, 1) : PCG_PICK2(T, M, 0))
    if (threads == 256 && minb == 5) fn = PCG_PICK(256, 5);
    else if (threads == 1024 && minb == 1) fn = PCG_PICK(1024, 1);
    else if (threads == 640 && minb == 2) fn = PCG_PICK(640, 2);
    else if (threads == 512 && minb == 2) fn = PCG_PICK(512, 2);
    else { set_error("PFEM_PCG_CFG: unsupported shape %dx%d", threads, minb); return PFEM_ERR_ARG; }
#undef PCG_PICK2
#undef PCG_PICK
    int per_sm = 0;
    PFEM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&per_sm, fn, threads, 0, 0));
    if (per_sm < minb) return PFEM_OK;
    const int grid = h->sm_count * minb;
    cudaStream_t s = h->stream;
    if (!h->pcg_bcast.p) PFEM_TRY(h->pcg_bcast.alloc(16));
    PFEM_CUDA(cudaMemsetAsync(h->pcg_bcast.p, 0, 16 * sizeof(double), s));
    PcgArgs a;
    memset(&a, 0, sizeof a);
    a.nloc = h->size_local; a.nslices = h->A.nslices; a.row_lo = h->row_lo; a.pc_type = h->pc_type;
    a.multi = h->nranks > 1 ? 1 : 0; a.has_off = (a.multi && h->nnz_off > 0) ? 1 : 0;
    a.pstride = h->sm_count * 16; a.n_send = a.multi ? h->send_displs[h->nranks] : 0;
    a.slice_off = h->A.slice_off.p; a.scol = h->A.col.p; a.sval = h->A.val.p;
    a.off_ptr = h->off_ptr.p; a.bcol = h->bcol.p; a.bval = h->bval.p; a.ghost = h->ghost_buf.p;
    a.rowptr = h->rowptr.p; a.col = h->col.p; a.val = h->val.p; a.b = h->rhs.p;
    a.x = h->x.p; a.r = h->r.p; a.z = h->z.p; a.p = h->p.p; a.w = h->w.p; a.dinv = h->dinv.p;
    a.partials = h->partials.p; a.st = h->cg.p; a.ctx = a.multi ? h->p2p_ctx.p : nullptr;
    a.send_idx = h->send_idx.p; a.send_dst = h->send_dst.p;
    {
        // halo flavour: tag-validated 16-byte entries (default) or values + per-neighbour flags (PFEM_PCG_HALO=flag)
        const char *henv = getenv("PFEM_PCG_HALO");
        a.halo_tag = (a.multi && h->send_dst_t.p && !(henv && strcmp(henv, "flag") == 0)) ? 1 : 0;
        if (a.halo_tag && henv && strcmp(henv, "tagf") == 0) a.halo_tag = 2;      // all threads push + a system fence by the senders
        if (a.halo_tag && henv && strcmp(henv, "tagall") == 0) a.halo_tag = 3;    // all threads push, no fence (first tag version)
        a.ghost_t = h->ghost_buf.p + h->ghost_tag_off;
        a.send_dst_t = h->send_dst_t.p;
    }
    a.bcast = h->pcg_bcast.p;
    if (!h->pcg_parts.p) PFEM_TRY(h->pcg_parts.alloc((size_t)2 * PCG_REPL * a.pstride * 4));
    a.parts = h->pcg_parts.p;
    a.arrive = reinterpret_cast<unsigned long long *>(h->pcg_bcast.p + 8);
    a.push_ticket = reinterpret_cast<unsigned int *>(h->pcg_bcast.p + 12);
    double *svp = h->sv.p;
    void *params[] = {(void *)&a, (void *)&svp};
    PFEM_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), params, 0, s));
    h->launches++;
    used = true;
    return PFEM_OK;
}

static int grid_for(pfem_solver *h, long long work_items, int per_thread)
{
    long long blocks = (work_items + (long long)CG_THREADS * per_thread - 1) / ((long long)CG_THREADS * per_thread);
    const long long cap = (long long)h->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

static int launch_spmv(pfem_solver *h, int mode)
{
    const int g = grid_for(h, (long long)h->A.nslices * 32, 1);
    const bool prof = h->profile && mode > 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (prof) {
        PFEM_CUDA(cudaEventCreate(&e0));
        PFEM_CUDA(cudaEventCreate(&e1));
        h->prof_ev.push_back(e0);
        h->prof_ev.push_back(e1);
        PFEM_CUDA(cudaEventRecord(e0, h->stream));
    }
    spmv_sell_kernel<<<g, CG_THREADS, 0, h->stream>>>(h->A.nslices, h->size_local, h->A.slice_off.p, h->A.col.p,
                                                      h->A.val.p, h->p.p, h->w.p, h->partials.p + 2 * h->sm_count * 16,
                                                      mode >= 0 ? h->cg.p : nullptr, mode < 0 ? 0 : mode);
    h->launches++;
    if (prof) PFEM_CUDA(cudaEventRecord(e1, h->stream));
    return PFEM_OK;
}

// sum the SpMV event pairs whose kernels actually ran (launches after convergence early-out in ~2 us: skip them)
static int collect_profile(pfem_solver *h, int its)
{
    const size_t pairs = h->prof_ev.size() / 2;
    for (size_t i = 0; i < pairs; i++) {
        if ((int)i < its) {
            float ms = 0.f;
            PFEM_CUDA(cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
            h->prof_spmv_s += ms * 1e-3;
            h->prof_spmv_n++;
        }
        cudaEventDestroy(h->prof_ev[2 * i]);
        cudaEventDestroy(h->prof_ev[2 * i + 1]);
    }
    h->prof_ev.clear();
    return PFEM_OK;
}

int cg_solve(pfem_solver *h)
{
    cudaStream_t s = h->stream;
    const int nloc = h->size_local, P = h->nranks;
    const int pstride = h->sm_count * 16;
    const bool multi = P > 1;
    PFEM_CUDA(cudaEventRecord(h->ev0, s));
    // KSPSetUp: move the assembled values into the solver layout
    if (h->nnz > 0) {
        values_to_solver_kernel<<<grid_for(h, h->nnz, 4), CG_THREADS, 0, s>>>(h->nnz, h->val.p, h->csr2sell.p, h->A.val.p, h->bval.p);
        h->launches++;
    }
    CgState init;
    memset(&init, 0, sizeof init);
    init.rtol = h->rtol; init.abstol = h->abstol; init.dtol = h->dtol; init.max_it = h->max_it;
    init.iter = (multi && !h->p2p) ? -1 : 0;
    init.seq = ++h->solve_seq;
    PFEM_CUDA(cudaMemcpyAsync(h->cg.p, &init, sizeof init, cudaMemcpyHostToDevice, s));
    bool persistent = false;
    PFEM_TRY(cg_solve_persistent(h, persistent));
    if (persistent) {
        PFEM_CUDA(cudaMemcpyAsync(h->cg_host, h->cg.p, sizeof(CgState), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaEventRecord(h->ev1, s));
        PFEM_CUDA(cudaEventSynchronize(h->ev1));
        PFEM_CUDA(cudaGetLastError());
        float msp = 0.f;
        PFEM_CUDA(cudaEventElapsedTime(&msp, h->ev0, h->ev1));
        h->t_solve = msp * 1e-3;
        h->its = h->cg_host->its; h->reason = h->cg_host->reason; h->rnorm = h->cg_host->dp;
        if (h->reason == -101) { set_error("cg: peer exchange timed out"); return PFEM_ERR_NCCL; }
        return PFEM_OK;
    }
    const int gv = grid_for(h, nloc, 2);
    // launch-per-phase path: NCCL exchange by default; its peer-memory variant (the stepping stone towards the persistent
    // kernel) only on request
    const char *kp = getenv("PFEM_KERNELS_P2P");
    const bool p2p = multi && h->p2p && kp && kp[0] == '1';
    const P2pCtx *ctx = p2p ? h->p2p_ctx.p : nullptr;
    const bool ilu = h->pc_type == PFEM_PC_BJACOBI_ILU0;
    if (ilu) {
        // PCSetUp (numeric ILU(0) of the diagonal block), x = 0, r = b, z = M^-1 r, then (z.z, z.r)
        cg_setup_kernel<<<gv, CG_THREADS, 0, s>>>(nloc, h->row_lo, h->pc_type, h->rowptr.p, h->col.p, h->val.p, h->rhs.p, h->x.p,
                                                  h->r.p, h->z.p, h->dinv.p, h->partials.p, pstride, h->cg.p, 0, ctx, 1);
        h->launches++;
        PFEM_TRY(ilu_setup(h));
        PFEM_TRY(ilu_apply(h));
    }
    cg_setup_kernel<<<gv, CG_THREADS, 0, s>>>(nloc, h->row_lo, h->pc_type, h->rowptr.p, h->col.p, h->val.p, h->rhs.p, h->x.p,
                                              h->r.p, h->z.p, h->dinv.p, h->partials.p, pstride, h->cg.p,
                                              p2p ? 2 : (multi ? 0 : 1), ctx, ilu ? 2 : 0);
    h->launches++;
    if (multi && !p2p) {
        PFEM_TRY(comm_allreduce_sum(h, h->cg.p->red, 2, s));
        scalar_after_setup_kernel<<<1, 1, 0, s>>>(h->cg.p);
        h->launches++;
    }
    const int chunk = 16;
    const int n_send = multi ? h->send_displs[P] : 0;
    int reason = 0;
    long long guard = 0;
    while (true) {
        for (int c = 0; c < chunk; c++) {
            cg_direction_kernel<<<gv, CG_THREADS, 0, s>>>(nloc, h->z.p, h->p.p, h->cg.p);
            h->launches++;
            if (!multi) {
                PFEM_TRY(launch_spmv(h, 1));
            } else if (p2p) {
                // halo values go straight into the neighbours' ghost buffers; the diagonal block multiplies meanwhile
                halo_push_kernel<<<grid_for(h, n_send > 0 ? n_send : 1, 1), CG_THREADS, 0, s>>>(n_send, h->send_idx.p, h->p.p, h->send_dst.p,
                                                                                            h->cg.p, ctx);
                h->launches++;
                PFEM_TRY(launch_spmv(h, 2));
                spmv_offdiag_p2p_kernel<<<grid_for(h, h->n_brows > 0 ? h->n_brows : 1, 1), CG_THREADS, 0, s>>>(
                    h->n_brows, h->brow_ids.p, h->brow_ptr.p, h->bcol.p, h->bval.p, h->ghost_buf.p, h->p.p, h->w.p,
                    h->partials.p + 3 * pstride, h->cg.p, ctx);
                h->launches++;
            } else {
                // halo: pack boundary values, exchange on the comm stream while the diagonal block multiplies
                if (n_send > 0) {
                    pack_halo_kernel<<<grid_for(h, n_send, 1), CG_THREADS, 0, s>>>(n_send, h->send_idx.p, h->p.p, h->send_buf.p, h->cg.p);
                    h->launches++;
                }
                PFEM_CUDA(cudaEventRecord(h->ev_pack, s));
                PFEM_CUDA(cudaStreamWaitEvent(h->comm_stream, h->ev_pack, 0));
                PFEM_TRY(comm_halo_exchange(h, h->send_buf.p, h->ghost_buf.p, h->comm_stream));
                PFEM_CUDA(cudaEventRecord(h->ev_halo, h->comm_stream));
                PFEM_TRY(launch_spmv(h, 2));
                PFEM_CUDA(cudaStreamWaitEvent(s, h->ev_halo, 0));
                spmv_offdiag_kernel<<<grid_for(h, h->n_brows > 0 ? h->n_brows : 1, 1), CG_THREADS, 0, s>>>(
                    h->n_brows, h->brow_ids.p, h->brow_ptr.p, h->bcol.p, h->bval.p, h->ghost_buf.p, h->p.p, h->w.p,
                    h->partials.p + 3 * pstride, h->cg.p);
                h->launches++;
                PFEM_TRY(comm_allreduce_sum(h, h->cg.p->red, 1, s));
                scalar_after_spmv_kernel<<<1, 1, 0, s>>>(h->cg.p);
                h->launches++;
            }
            if (ilu) {
                cg_update_kernel<<<gv, CG_THREADS, 0, s>>>(nloc, h->p.p, h->w.p, h->dinv.p, h->x.p, h->r.p, h->z.p, h->partials.p,
                                                           pstride, h->cg.p, 0, ctx, 1);
                h->launches++;
                PFEM_TRY(ilu_apply(h));
            }
            cg_update_kernel<<<gv, CG_THREADS, 0, s>>>(nloc, h->p.p, h->w.p, h->dinv.p, h->x.p, h->r.p, h->z.p, h->partials.p,
                                                       pstride, h->cg.p, p2p ? 2 : (multi ? 0 : 1), ctx, ilu ? 2 : 0);
            h->launches++;
            if (multi && !p2p) {
                PFEM_TRY(comm_allreduce_sum(h, h->cg.p->red, 2, s));
                scalar_after_update_kernel<<<1, 1, 0, s>>>(h->cg.p);
                h->launches++;
            }
        }
        PFEM_CUDA(cudaMemcpyAsync(h->cg_host, h->cg.p, sizeof(CgState), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaStreamSynchronize(s));
        reason = h->cg_host->reason;
        if (reason != 0) break;
        guard += chunk;
        if (guard > (long long)h->max_it + 2 * chunk) { set_error("cg: iteration guard tripped"); return PFEM_ERR_STATE; }
    }
    PFEM_CUDA(cudaEventRecord(h->ev1, s));
    PFEM_CUDA(cudaEventSynchronize(h->ev1));
    PFEM_CUDA(cudaGetLastError());
    float ms = 0.f;
    PFEM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->t_solve = ms * 1e-3;
    h->its = h->cg_host->its; h->reason = reason; h->rnorm = h->cg_host->dp;
    if (h->profile) PFEM_TRY(collect_profile(h, h->its));
    return PFEM_OK;
}

int time_spmv(pfem_solver *h, int reps, double *seconds)
{
    if (reps < 1) reps = 1;
    cudaStream_t s = h->stream;
    if (h->nnz > 0) {
        values_to_solver_kernel<<<grid_for(h, h->nnz, 4), CG_THREADS, 0, s>>>(h->nnz, h->val.p, h->csr2sell.p, h->A.val.p, h->bval.p);
        h->launches++;
    }
    PFEM_CUDA(cudaMemcpyAsync(h->p.p, h->rhs.p, (size_t)h->size_local * sizeof(double), cudaMemcpyDeviceToDevice, s));
    for (int i = 0; i < 3; i++) PFEM_TRY(launch_spmv(h, -1));
    PFEM_CUDA(cudaEventRecord(h->ev0, s));
    for (int i = 0; i < reps; i++) PFEM_TRY(launch_spmv(h, -1));
    PFEM_CUDA(cudaEventRecord(h->ev1, s));
    PFEM_CUDA(cudaEventSynchronize(h->ev1));
    PFEM_CUDA(cudaGetLastError());
    float ms = 0.f;
    PFEM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    *seconds = ms * 1e-3 / reps;
    return PFEM_OK;
}

}  // namespace pfem

extern "C" int pfem_debug_pcg_trace(long long *out64)
{
#ifdef PFEM_PCG_TRACE
    return cudaMemcpyFromSymbol(out64, pfem::g_pcg_trace, 64 * sizeof(long long)) == cudaSuccess ? 0 : PFEM_ERR_CUDA;
#else
    (void)out64;
    return PFEM_ERR_STATE;      // not a trace build
#endif
}
