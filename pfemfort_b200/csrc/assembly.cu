// assembly.cu -- the value pass: fused Ke/Fe + MatSetValues(ADD) + Dirichlet lifting + VecSetValues(ADD).
//
// Replaces the element loop of the *parallelimpl1 drivers (tetrapoissonparallelimpl1.F:828-884,
// tetraelasticityparallelimpl1.F:901-968) and PETSc's MatSetValues/VecSetValues behind it.
//
// Design (row gather, no atomics): one thread owns one matrix row.  It walks the row's incident
// elements in ascending element id, recomputes the element geometry, and adds the row's slice of
// Klocal (read transposed, as PETSc reads the column-major Fortran block) into the row's CSR segment,
// which lives in shared memory for the whole CTA (a CTA owns R consecutive rows = one contiguous CSR
// segment, loaded and stored with coalesced accesses).  Per matrix entry the contributions are thus
// summed in exactly the order of the reference's sequential np=1 run, so the result is deterministic and,
// because the TU is built with -fmad=false, bit-identical to the no-FMA CPU evaluation.
#include <cstdlib>
#include <cstring>

#include "assembly_rows.cuh"
#include "internal.cuh"

namespace pfem {

// error path only: number of elements with Jac < 0 among those that touch an owned row
template <int KIND>
__global__ void count_negative_kernel(int nElem, int rec_ints, const int *__restrict__ erec, const double *__restrict__ xyz,
                                      int row_lo, int row_hi, int *__restrict__ count)
{
    using T = ElemTraits<KIND>;
    constexpr int NPE = T::NPE, NDIM = T::NDIM, NSIZE = T::NPE * T::NDOF;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nElem; e += gridDim.x * blockDim.x) {
        const int *rec = erec + (size_t)e * rec_ints;
        bool owned = false;
        for (int q = 0; q < NSIZE; q++) owned |= rec[NPE + q] >= row_lo && rec[NPE + q] < row_hi;
        if (!owned) continue;
        int nodes[NPE];
        for (int i = 0; i < NPE; i++) nodes[i] = rec[i];
        double x[NPE], y[NPE], z[NPE];
        load_coords<NPE, NDIM>(xyz, nodes, x, y, z);
        ElemOp<KIND> op;
        op.load_geom(x, y, z);
        if (op.g.Jac < 0.0) atomicAdd(count, 1);
    }
}

template <int KIND, int R>
static int launch_assemble_sell(pfem_solver *h, const AsmArgs &args)
{
    const int blocks = ceil_div(h->size_local, R);
    if (blocks == 0) return PFEM_OK;
    constexpr bool POISSON = KIND == POISSON_TRIA || KIND == POISSON_TETRA;
    const size_t smem = h->asm_smem;                                   // includes one sink accumulator per thread
    if (POISSON && args.unit) {
        PFEM_CUDA(cudaFuncSetAttribute(assemble_sell_kernel<KIND, R, POISSON>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        assemble_sell_kernel<KIND, R, POISSON><<<blocks, R, smem, h->stream>>>(args);
    } else {
        PFEM_CUDA(cudaFuncSetAttribute(assemble_sell_kernel<KIND, R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        assemble_sell_kernel<KIND, R, false><<<blocks, R, smem, h->stream>>>(args);
    }
    h->launches++;
    PFEM_CUDA(cudaGetLastError());
    return PFEM_OK;
}

template <int KIND, int R>
static int launch_assemble(pfem_solver *h, const AsmArgs &args)
{
    const int blocks = ceil_div(h->size_local, R);
    if (blocks == 0) return PFEM_OK;
    PFEM_CUDA(cudaFuncSetAttribute(assemble_kernel<KIND, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->asm_smem));
    assemble_kernel<KIND, R><<<blocks, R, h->asm_smem, h->stream>>>(args);
    h->launches++;
    PFEM_CUDA(cudaGetLastError());
    return PFEM_OK;
}

template <int KIND>
static int dispatch_rows(pfem_solver *h, const AsmArgs &args)
{
    if (h->asm_sell) {
        switch (h->asm_rows_per_cta) {
        case 256: return launch_assemble_sell<KIND, 256>(h, args);
        case 128: return launch_assemble_sell<KIND, 128>(h, args);
        case 64: return launch_assemble_sell<KIND, 64>(h, args);
        case 32: return launch_assemble_sell<KIND, 32>(h, args);
        }
    }
    switch (h->asm_rows_per_cta) {
    case 256: return launch_assemble<KIND, 256>(h, args);
    case 128: return launch_assemble<KIND, 128>(h, args);
    case 64: return launch_assemble<KIND, 64>(h, args);
    case 32: return launch_assemble<KIND, 32>(h, args);
    }
    set_error("assembly: no CTA shape fits shared memory");
    return PFEM_ERR_SIZE;
}

// largest CSR / incidence segment over the R-row CTAs, for the four CTA shapes at once
__global__ void segment_max_kernel(int nloc, const int *__restrict__ rowptr, const int *__restrict__ rinc_ptr, int *__restrict__ out)
{
    const int shapes[4] = {256, 128, 64, 32};
    int mx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r0 = (blockIdx.x * blockDim.x + threadIdx.x) * 32; r0 < nloc; r0 += gridDim.x * blockDim.x * 32) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (r0 % shapes[q]) continue;
            const int r1 = min(r0 + shapes[q], nloc);
            mx[q] = max(mx[q], rowptr[r1] - rowptr[r0]);
            mx[4 + q] = max(mx[4 + q], rinc_ptr[r1] - rinc_ptr[r0]);
        }
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
        int v = mx[q];
        for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out + q, v);
    }
}

// Pick the rows-per-CTA so that the largest CTA segment (CSR values + columns + incidences) fits in smem.
int plan_assembly(pfem_solver *h)
{
    const int nloc = h->size_local;
    DevBuf<int> dmx;
    PFEM_TRY(dmx.alloc(8));
    PFEM_CUDA(cudaMemsetAsync(dmx.p, 0, 8 * sizeof(int), h->stream));
    segment_max_kernel<<<h->sm_count * 4, 256, 0, h->stream>>>(nloc, h->rowptr.p, h->rinc_ptr.p, dmx.p);
    h->launches++;
    int mxs[8];
    PFEM_CUDA(cudaMemcpyAsync(mxs, dmx.p, sizeof mxs, cudaMemcpyDeviceToHost, h->stream));
    PFEM_CUDA(cudaStreamSynchronize(h->stream));
    int max_smem = 0;
    PFEM_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    const int shapes[4] = {256, 128, 64, 32};
    const char *env = getenv("PFEM_ASM_ROWS");            // tuning hook: force the rows-per-CTA shape
    const int forced = env ? atoi(env) : 0;
    // prefer the largest shape that still lets two CTAs share an SM
    for (int pass = 0; pass < 2; pass++) {
        const size_t limit = pass == 0 ? (size_t)max_smem / 2 - 1024 : (size_t)max_smem;
        for (int idx = 0; idx < 4; idx++) {
            // the streamed kernel is register-bound and barrier-free at one warp per CTA: prefer the smallest shape
            const int q = h->asm_sell ? 3 - idx : idx;
            const int R = shapes[q];
            if (forced && R != forced) continue;
            int mn = mxs[q];
            const int mi = mxs[4 + q];
            mn = (mn + 1) & ~1;   // keep the int arrays 8-byte aligned
            const size_t bytes = h->asm_sell ? (size_t)mn * 8 + 16 + (size_t)R * 8 : (size_t)mn * 12 + (size_t)mi * 4 + 16;
            if (bytes <= limit) {
                h->asm_rows_per_cta = R;
                h->asm_smem = bytes;
                h->asm_max_seg = mn;
                return PFEM_OK;
            }
        }
    }
    set_error("assembly: a 32-row CSR segment does not fit in %d bytes of shared memory", max_smem);
    return PFEM_ERR_SIZE;
}

// which kernel runs the pass: PFEM_ASM (rows | fast | ctile | tiled | tiled2) overrides the handle's request
// (pfem_solver_set_assembly_mode).  Default = "rows": the streamed row-gather kernel with the reference-order no-FMA element
// operators (sequential summation order: bit-identical to the sequential CPU evaluation); "fast" = the same kernel with the
// FMA cofactor-form operators (1e-12 contract; measured 5 % faster on C5: the kernel is issue-bound on its non-FP64
// instructions, not on the FP64 pipe); "ctile" = the colour-scheduled tile kernel (compute-once, TMA-staged; measured slower
// than the row gather on B200: shared-memory read-modify-write traffic and barriers, profiles/r02_value_pass.md).
int dispatch_rows_fast(pfem_solver *h, const AsmArgs &args);     // assembly_fast.cu

static int pick_mode(pfem_solver *h, int &mode)
{
    const char *env = getenv("PFEM_ASM");
    int want = h->asm_mode_req == PFEM_ASM_FAST ? 0 : 1;            // 0 fast, 1 rows (default), 2 tiled, 3 tiled2, 4 ctile
    if (env) {
        if (!strcmp(env, "rows")) want = 1;
        else if (!strcmp(env, "tiled")) want = 2;
        else if (!strcmp(env, "tiled2")) want = 3;
        else if (!strcmp(env, "ctile")) want = 4;
        else if (!strcmp(env, "fast")) want = 0;
        else if (!strcmp(env, "auto") || !env[0]) want = 1;
        else { set_error("PFEM_ASM=%s: expected fast | rows | ctile | tiled | tiled2 | auto", env); return PFEM_ERR_ARG; }
    }
    if (want == 4) {
        if (h->ndof == 1 && !h->ct_ready && !h->ct_tried) { h->ct_tried = true; PFEM_TRY(build_ctiles(h)); }
        if (h->ct_ready) { mode = 3; return PFEM_OK; }
        set_error("PFEM_ASM=ctile: the tile kernel does not apply to this pattern (kind %d)", h->kind);
        return PFEM_ERR_ARG;
    }
    if (!h->rows_ready) {
        PFEM_TRY(build_asm_streams(h));
        h->asm_rows_per_cta = 0;
        h->rows_ready = true;
    }
    if ((want == 2 || want == 3) && h->ndof == 1) {                // round-1 opt-in tiles (host-built from the row streams)
        const int tile_mode = want - 1;
        if (!h->tiles_ready || h->tile_mode != tile_mode) PFEM_TRY(build_tiles_device(h, tile_mode));
        if (h->asm_tiled) { mode = 2; return PFEM_OK; }
    }
    if (h->asm_rows_per_cta == 0) PFEM_TRY(plan_assembly(h));
    mode = h->asm_sell ? (want == 0 ? 4 : 1) : 0;
    return PFEM_OK;
}

int assemble_values(pfem_solver *h, const double *elemData, const double *timeData, int *n_neg)
{
    int mode = 0;
    PFEM_TRY(pick_mode(h, mode));
    DevBuf<double> dED, dTD;
    PFEM_TRY(dED.alloc(8));
    PFEM_TRY(dTD.alloc(8));
    double ed[8] = {0}, td[8] = {0};
    const int ned = h->kind == PFEM_POISSON_TRIA ? 2 : h->kind == PFEM_POISSON_TETRA ? 3 : h->kind == PFEM_ELASTICITY_TRIA ? 5 : 6;
    for (int i = 0; i < ned; i++) ed[i] = elemData[i];
    td[1] = timeData[1];
    cudaStream_t s = h->stream;
    PFEM_CUDA(cudaMemcpyAsync(dED.p, ed, sizeof ed, cudaMemcpyHostToDevice, s));
    PFEM_CUDA(cudaMemcpyAsync(dTD.p, td, sizeof td, cudaMemcpyHostToDevice, s));
    PFEM_CUDA(cudaMemsetAsync(h->neg_count.p, 0, sizeof(int), s));
    AsmArgs a;
    a.nloc = h->size_local; a.row_lo = h->row_lo; a.row_hi = h->row_hi; a.rec_ints = h->rec_ints;
    a.erec = h->erec.p; a.xyz = h->xyz.p; a.applied = h->applied.p;
    a.rowptr = h->rowptr.p; a.col = h->col.p; a.val = h->val.p; a.rhs = h->rhs.p;
    a.rinc_ptr = h->rinc_ptr.p; a.rinc = h->rinc.p;
    a.elemData = dED.p; a.timeData = dTD.p; a.neg_count = h->neg_count.p;
    a.load_val = h->values_zero ? 0 : 1; a.load_rhs = h->rhs_zero ? 0 : 1;
    a.max_seg_nnz = h->asm_max_seg;
    a.ainc_off = h->ainc_off.p; a.ainc = h->ainc.p; a.conn4 = h->conn4.p; a.neg_flag = h->neg_count.p;
    a.unit = (td[1] == 1.0 && ed[0] == 1.0 && ed[1] == 1.0 && (h->kind == PFEM_POISSON_TRIA || ed[2] == 1.0)) ? 1 : 0;
    PFEM_CUDA(cudaEventRecord(h->ev0, s));
    int st = PFEM_OK;
    h->last_asm_mode = mode;
    if (mode == 3) st = assemble_values_ctile(h, dED.p, dTD.p);
    else if (mode == 4) st = dispatch_rows_fast(h, a);
    else if (mode == 2) st = assemble_values_tiled(h, dED.p, dTD.p, a.unit != 0);
    else switch (h->kind) {
    case PFEM_POISSON_TRIA: st = dispatch_rows<POISSON_TRIA>(h, a); break;
    case PFEM_POISSON_TETRA: st = dispatch_rows<POISSON_TETRA>(h, a); break;
    case PFEM_ELASTICITY_TRIA: st = dispatch_rows<ELASTICITY_TRIA>(h, a); break;
    case PFEM_ELASTICITY_TETRA: st = dispatch_rows<ELASTICITY_TETRA>(h, a); break;
    default: set_error("assemble: mesh kind not set"); return PFEM_ERR_STATE;
    }
    PFEM_TRY(st);
    PFEM_CUDA(cudaEventRecord(h->ev1, s));
    int neg = 0;
    PFEM_CUDA(cudaMemcpyAsync(&neg, h->neg_count.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    if (neg && mode != 0) {        // the streamed / tile kernels only raise a flag: count the offending elements now
        PFEM_CUDA(cudaMemsetAsync(h->neg_count.p, 0, sizeof(int), s));
        const int G = h->sm_count * 8;
        switch (h->kind) {
        case PFEM_POISSON_TRIA: count_negative_kernel<POISSON_TRIA><<<G, 256, 0, s>>>(h->nElem, h->rec_ints, h->erec.p, h->xyz.p, h->row_lo, h->row_hi, h->neg_count.p); break;
        case PFEM_POISSON_TETRA: count_negative_kernel<POISSON_TETRA><<<G, 256, 0, s>>>(h->nElem, h->rec_ints, h->erec.p, h->xyz.p, h->row_lo, h->row_hi, h->neg_count.p); break;
        case PFEM_ELASTICITY_TRIA: count_negative_kernel<ELASTICITY_TRIA><<<G, 256, 0, s>>>(h->nElem, h->rec_ints, h->erec.p, h->xyz.p, h->row_lo, h->row_hi, h->neg_count.p); break;
        default: count_negative_kernel<ELASTICITY_TETRA><<<G, 256, 0, s>>>(h->nElem, h->rec_ints, h->erec.p, h->xyz.p, h->row_lo, h->row_hi, h->neg_count.p); break;
        }
        h->launches++;
        PFEM_CUDA(cudaMemcpyAsync(&neg, h->neg_count.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        PFEM_CUDA(cudaStreamSynchronize(s));
    }
    float ms = 0.f;
    PFEM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->t_assemble = ms * 1e-3;
    h->values_zero = false;
    h->rhs_zero = false;
    if (n_neg) *n_neg = neg;
    if (neg) {
        set_error("assembly: %d element(s) with negative Jacobian (the reference STOPs)", neg);
        return PFEM_ERR_NEG_JACOBIAN;
    }
    return PFEM_OK;
}

// ---- slow-path adds: MatSetValues / VecSetValues / MatSetValue mirrors ----------------------------------------
// One thread walks the block in PETSc's order (row by row, column by column).  Rows of this rank are added in place;
// rows owned by another rank are NOT dropped: like PETSc's stash they are kept on the host and shipped to their owner
// at the next assembly point (stash_flush, called by pfem_solver_solve = MatAssemblyBegin/End of solverpetsc.F:447-468),
// where they are added in arrival order (source rank ascending, call order within a rank).  A location that the pattern
// pass did not create cannot be inserted later (the structure is fixed; PETSc would allocate it under
// MAT_NEW_NONZERO_LOCATIONS): it is counted and reported as PFEM_ERR_PATTERN, never skipped silently.
__global__ void add_entries_kernel(int n, const int *__restrict__ rows, const int *__restrict__ cols,
                                   const double *__restrict__ vals, int transposed, const double *__restrict__ F,
                                   int row_lo, int row_hi, const int *__restrict__ rowptr, const int *__restrict__ col,
                                   double *__restrict__ val, double *__restrict__ rhs, int *__restrict__ off_pattern)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int missed = 0;
    for (int i = 0; i < n; i++) {
        const int r = rows[i];
        if (r < row_lo || r >= row_hi) continue;   // negative rows are dropped like PETSc does; other ranks' rows went to the stash
        const int lr = r - row_lo;
        if (F) rhs[lr] = rhs[lr] + F[i];
        if (!vals) continue;
        for (int j = 0; j < n; j++) {
            const int c = cols[j];
            if (c < 0) continue;
            int lo = rowptr[lr], hi = rowptr[lr + 1];
            const int end = hi;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (col[mid] < c) lo = mid + 1; else hi = mid;
            }
            if (lo < end && col[lo] == c) {
                const double v = transposed ? vals[j + n * i] : vals[i + n * j];
                val[lo] = val[lo] + v;
            } else {
                missed++;
            }
        }
    }
    if (missed) atomicAdd(off_pattern, missed);
}

// stashed (row, col, value) triples received from the other ranks, in arrival order; col < 0: right-hand side entry
__global__ void add_triples_kernel(int n, const int *__restrict__ t, int row_lo, int row_hi, const int *__restrict__ rowptr,
                                   const int *__restrict__ col, double *__restrict__ val, double *__restrict__ rhs,
                                   int *__restrict__ off_pattern)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int missed = 0;
    for (int k = 0; k < n; k++) {
        const int r = t[4 * k], c = t[4 * k + 1];
        const double v = __hiloint2double(t[4 * k + 3], t[4 * k + 2]);
        if (r < row_lo || r >= row_hi) { missed++; continue; }
        const int lr = r - row_lo;
        if (c < 0) { rhs[lr] = rhs[lr] + v; continue; }
        int lo = rowptr[lr], hi = rowptr[lr + 1];
        const int end = hi;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (col[mid] < c) lo = mid + 1; else hi = mid;
        }
        if (lo < end && col[lo] == c) val[lo] = val[lo] + v; else missed++;
    }
    if (missed) atomicAdd(off_pattern, missed);
}

static void stash_push(pfem_solver *h, int owner, int row, int col, double v)
{
    long long bits;
    memcpy(&bits, &v, sizeof bits);
    std::vector<int> &st = h->stash[owner];
    st.push_back(row); st.push_back(col);
    st.push_back((int)(bits & 0xffffffffLL)); st.push_back((int)((bits >> 32) & 0xffffffffLL));
}

int add_entries(pfem_solver *h, int n, const int *rows, const int *cols, const double *vals, bool transposed,
                const double *F)
{
    if (n <= 0 || !rows) { set_error("add: bad argument"); return PFEM_ERR_ARG; }
    if (vals && !cols) { set_error("add: cols missing"); return PFEM_ERR_ARG; }
    for (int i = 0; i < n; i++)
        if (rows[i] >= h->size_global || (vals && cols[i] >= h->size_global)) {      // PETSc: "Row too large" / "Column too large"
            set_error("add: index %d outside the global system (size %d)", rows[i] >= h->size_global ? rows[i] : cols[i], h->size_global);
            return PFEM_ERR_ARG;
        }
    // rows of other ranks: stash (MatSetValues / VecSetValues on an off-process row)
    if (h->nranks > 1) {
        if ((int)h->stash.size() != h->nranks) h->stash.assign(h->nranks, std::vector<int>());
        for (int i = 0; i < n; i++) {
            const int r = rows[i];
            if (r < 0 || (r >= h->row_lo && r < h->row_hi)) continue;
            int owner = 0;
            while (owner + 1 < h->nranks && r >= h->row_starts[owner + 1]) owner++;
            if (F) stash_push(h, owner, r, -1, F[i]);
            if (vals)
                for (int j = 0; j < n; j++)
                    if (cols[j] >= 0) stash_push(h, owner, r, cols[j], transposed ? vals[j + (size_t)n * i] : vals[i + (size_t)n * j]);
        }
    }
    DevBuf<int> dr, dc;
    DevBuf<double> dv, df;
    cudaStream_t s = h->stream;
    PFEM_TRY(dr.alloc(n));
    PFEM_CUDA(cudaMemcpyAsync(dr.p, rows, n * sizeof(int), cudaMemcpyHostToDevice, s));
    if (vals) {
        PFEM_TRY(dc.alloc(n));
        PFEM_TRY(dv.alloc((size_t)n * n));
        PFEM_CUDA(cudaMemcpyAsync(dc.p, cols, n * sizeof(int), cudaMemcpyHostToDevice, s));
        PFEM_CUDA(cudaMemcpyAsync(dv.p, vals, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, s));
    }
    if (F) {
        PFEM_TRY(df.alloc(n));
        PFEM_CUDA(cudaMemcpyAsync(df.p, F, n * sizeof(double), cudaMemcpyHostToDevice, s));
    }
    PFEM_CUDA(cudaMemsetAsync(h->neg_count.p, 0, sizeof(int), s));
    add_entries_kernel<<<1, 32, 0, s>>>(n, dr.p, vals ? dc.p : nullptr, vals ? dv.p : nullptr, transposed ? 1 : 0,
                                        F ? df.p : nullptr, h->row_lo, h->row_hi, h->rowptr.p, h->col.p, h->val.p, h->rhs.p,
                                        h->neg_count.p);
    h->launches++;
    PFEM_CUDA(cudaGetLastError());
    int missed = 0;
    PFEM_CUDA(cudaMemcpyAsync(&missed, h->neg_count.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    if (vals) h->values_zero = false;
    if (F) h->rhs_zero = false;
    if (missed) {
        h->off_pattern_total += missed;
        set_error("add: %d matrix location(s) are not in the pattern built by the pattern pass (new nonzero locations cannot be "
                  "inserted after it; entries inside the pattern were added)", missed);
        return PFEM_ERR_PATTERN;
    }
    return PFEM_OK;
}

// MatAssemblyBegin/End + VecAssemblyBegin/End (solverpetsc.F:447-468) for the slow-path adds: collective over the ranks.
int stash_flush(pfem_solver *h)
{
    if (h->nranks == 1) return PFEM_OK;
    if ((int)h->stash.size() != h->nranks) h->stash.assign(h->nranks, std::vector<int>());
    int mine = 0;
    for (int q = 0; q < h->nranks; q++) mine += (int)(h->stash[q].size() / 4);
    std::vector<int> all;
    PFEM_TRY(comm_allgather_int(h, mine, all));
    long long total = 0;
    for (int v : all) total += v;
    if (total == 0) return PFEM_OK;
    std::vector<int> sendbuf, sendcounts(h->nranks, 0), recvbuf, recvcounts;
    for (int q = 0; q < h->nranks; q++) {
        sendcounts[q] = (int)h->stash[q].size();
        sendbuf.insert(sendbuf.end(), h->stash[q].begin(), h->stash[q].end());
        h->stash[q].clear();
    }
    PFEM_TRY(comm_alltoallv_int(h, sendbuf, sendcounts, recvbuf, recvcounts));
    const int ntrip = (int)(recvbuf.size() / 4);
    h->stash_received += ntrip;
    if (ntrip == 0) return PFEM_OK;
    cudaStream_t s = h->stream;
    DevBuf<int> dt;
    PFEM_TRY(dt.alloc(recvbuf.size()));
    PFEM_CUDA(cudaMemcpyAsync(dt.p, recvbuf.data(), recvbuf.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    PFEM_CUDA(cudaMemsetAsync(h->neg_count.p, 0, sizeof(int), s));
    add_triples_kernel<<<1, 32, 0, s>>>(ntrip, dt.p, h->row_lo, h->row_hi, h->rowptr.p, h->col.p, h->val.p, h->rhs.p, h->neg_count.p);
    h->launches++;
    int missed = 0;
    PFEM_CUDA(cudaMemcpyAsync(&missed, h->neg_count.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    PFEM_CUDA(cudaStreamSynchronize(s));
    h->values_zero = false; h->rhs_zero = false;
    if (missed) {
        h->off_pattern_total += missed;
        set_error("assembly: %d stashed entr(ies) from other ranks are not in this rank's pattern", missed);
        return PFEM_ERR_PATTERN;
    }
    return PFEM_OK;
}

}  // namespace pfem
