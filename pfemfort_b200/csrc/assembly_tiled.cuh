// assembly_tiled.cuh -- the tiled (compute-once) value pass kernel for the one-dof-per-node kinds (Poisson tria/tet).
//
// Same job as assemble_sell_kernel (assembly.cu): the element loop of tetrapoissonparallelimpl1.F:828-884 /
// triapoissonparallelimpl1.F:849-905 with PETSc's MatSetValues/VecSetValues(ADD) behind it, fused, without atomics,
// summing every matrix entry in ascending element id (the reference's sequential np=1 order) with the no-FMA
// arithmetic of elements.cuh => bit-identical to the row-gather kernel and to the CPU oracle.
//
// What changes is WHO computes an element.  The row-gather kernel recomputes the geometry of an element once per
// incident row (4x for a tet) and is FP64-pipe/issue bound at ~22 % of the HBM roof.  Here a CTA owns a tile of
// spatially close rows (tiles.hpp) and
//   phase A: its threads compute each element that touches the tile ONCE and stage the columns Klocal(:,k) and the
//            lifted Flocal(k) of the tile's own local dofs k in shared memory;
//   phase B: one thread per tile row walks the row's incidence stream (coalesced {staged column, slot bytes}
//            entries, ascending element id) and adds the staged column into the row's FP64 accumulators in shared
//            memory at the precomputed slots (Dirichlet columns go to a sink: branch-free), and Flocal into the RHS;
//   phase C: the accumulators are streamed to the CSR value array, one warp per row.
//
// This header is also compiled for the host by tests/emu (PFEM_EMULATE + a small CUDA shim), so it uses no warp
// intrinsics; inline PTX is confined to tiled_ld_xyz.
#pragma once
#include "elements.cuh"
#include "tiles.hpp"

namespace pfem {

struct TiledArgs {
    const int *tdesc;                 // [ntiles][TILE_DESC_INTS]
    const int4 *trows;                // { local row or -1, accumulator offset, rowptr[row], row length }
    const int2 *tel;                  // { e | dbc<<31, base | mask<<24 }
    const long long *tslice_off;
    const int2 *tinc;                 // { staged column or -1, slot bytes }
    const int2 *crec;                 // scatter kernel: per staged column { row base | p0<<16 | p1<<24, p2 | p3<<8 | pF<<16 }
    const uint4 *cnt;                 // scatter kernel: run lengths, 16 per chunk
    const int4 *ts2;                  // scatter kernel: per tile slice { buffer offset, first chunk, chunks, width }
    const int4 *conn4;                // [nElem] 0-based NEW node ids
    const int *erec;                  // [nElem][rec_ints] conn + dofs (Dirichlet elements only)
    int rec_ints;
    const double *xyz, *applied;
    const int *rowptr;
    double *val, *rhs;
    const double *elemData, *timeData;
    int *neg_flag;
    int load_val, load_rhs;
};

// 256-bit read-only load of one node's (x, y, z, pad)
__device__ __forceinline__ void tiled_ld_xyz(const double *p, double &x, double &y, double &z)
{
#if defined(__CUDA_ARCH__)
    double w;                                                // the pad lane of the 256-bit load
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(p));
    (void)w;
#else
    x = p[0]; y = p[1]; z = p[2];
#endif
}

__device__ __forceinline__ double tiled_kcoef(const Params<POISSON_TRIA> &p, int d) { return d == 0 ? p.kx : p.ky; }
__device__ __forceinline__ double tiled_kcoef(const Params<POISSON_TETRA> &p, int d) { return d == 0 ? p.kx : (d == 1 ? p.ky : p.kz); }

// Staging layout in shared memory.  A staged column is 4 doubles = two 16-byte chunks; four columns share a 128-byte
// line.  In phase A consecutive lanes write columns 4 apart (one element each), i.e. the same chunk of consecutive lines:
// the chunk index is XOR-swizzled with the line number so that the 8 lanes of a quarter warp hit 8 different bank
// groups.  Flocal is stored structure-of-arrays by (column & 3) for the same reason.
__device__ __forceinline__ int tiled_kst_off(int col, int half)      // in doubles
{
    const int line = col >> 2, chunk = ((((col & 3) << 1) | half) ^ line) & 7;
    return line * 16 + chunk * 2;
}
__device__ __forceinline__ int tiled_fst_off(int col, int nlines) { return (col & 3) * nlines + (col >> 2); }

#ifndef PFEM_DYN_SMEM
#define PFEM_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// Phase A of both tiled kernels: every element of the tile's list once.  emit(col, kc, f, pre) receives, for each local
// dof k of the element whose row belongs to the tile, the running column index of the tile (base + rank of k), the column
// kc[j] = Klocal(j, k) that MatSetValues(ADD) adds into row k, and the lifted Flocal(k).  prefetch(base) runs at the top
// of the element's iteration, before the geometry arithmetic: whatever per-column metadata emit needs is requested there
// and handed to emit as `pre` (emit consumes it front to back, one column at a time).
template <int KIND, int THREADS, bool UNIT, class Prefetch, class Emit>
__device__ __forceinline__ void tiled_phase_a(const TiledArgs &a, int el_off, int nel, Prefetch prefetch, Emit emit)
{
    using T = ElemTraits<KIND>;
    constexpr int NPE = T::NPE, NDIM = T::NDIM, NSIZE = NPE;
    const int tid = threadIdx.x;
    Params<KIND> prm;
    prm.init(a.elemData, a.timeData);
    const int2 *tel = a.tel + el_off;
    auto load_te = [&](int i) { return i < nel ? __ldcs(tel + i) : make_int2(0, 0); };
    auto load_conn = [&](const int2 &te) { return __ldg(a.conn4 + (te.x & 0x7fffffff)); };   // element 0 past the end
    auto load_xyz = [&](const int4 &c, double (&x)[NPE], double (&y)[NPE], double (&z)[NPE]) {
        const int nd[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int i = 0; i < NPE; i++) {
            if (NDIM == 3) tiled_ld_xyz(a.xyz + (size_t)nd[i] * 4, x[i], y[i], z[i]);
            else {
                const double2 t = __ldg(reinterpret_cast<const double2 *>(a.xyz + (size_t)nd[i] * 2));
                x[i] = t.x; y[i] = t.y; z[i] = 0.0;
            }
        }
    };
    // software pipeline: tile entries three iterations ahead, node ids two ahead, next coordinates in flight
    int2 te0 = load_te(tid), te1 = load_te(tid + THREADS), te2 = load_te(tid + 2 * THREADS);
    int4 c0 = load_conn(te0), c1 = load_conn(te1);
    double xq[NPE], yq[NPE], zq[NPE];
    load_xyz(c0, xq, yq, zq);
    for (int i = tid; i < nel; i += THREADS) {
        const int2 te3 = load_te(i + 3 * THREADS);
        const int4 c2 = load_conn(te2);
        double x[NPE], y[NPE], z[NPE];
#pragma unroll
        for (int q = 0; q < NPE; q++) { x[q] = xq[q]; y[q] = yq[q]; z[q] = zq[q]; }
        load_xyz(c1, xq, yq, zq);                          // next element: in flight during this one's arithmetic
        const int2 te = te0;
        const int4 cn = c0;
        te0 = te1; te1 = te2; te2 = te3; c0 = c1; c1 = c2;

        const int e = te.x & 0x7fffffff;
        const unsigned int mask = ((unsigned int)te.y >> 24) & 15u;
        int col = te.y & 0xffffff;
        auto pre = prefetch(col);
        ElemOp<KIND> op;
        op.load_geom(x, y, z);
        const bool neg = op.g.Jac < 0.0;                   // the reference STOPs here: flag it, stage zeros
        if (neg) atomicOr(a.neg_flag, 1);
        op.set_dvol(prm);
        // b_d(j) = dN_d(j) * dvol (poisson.F:87-89,177-179): shared by every column of the element
        double bd[NDIM][NPE];
#pragma unroll
        for (int d = 0; d < NDIM; d++)
#pragma unroll
            for (int j = 0; j < NPE; j++) bd[d][j] = op.g.dN[d][j] * op.dvol;
        // Dirichlet data of the element (rare): which local dofs are fixed, and their applied values
        bool fixed[NSIZE];
        double gval[NSIZE];
#pragma unroll
        for (int j = 0; j < NSIZE; j++) { fixed[j] = false; gval[j] = 0.0; }
        if (te.x < 0) {
            const int nd[4] = {cn.x, cn.y, cn.z, cn.w};
            const int *dof = a.erec + (size_t)e * a.rec_ints + NPE;
#pragma unroll
            for (int j = 0; j < NSIZE; j++) {
                fixed[j] = dof[j] == -1;
                if (fixed[j]) gval[j] = a.applied[nd[j]];
            }
        }
#pragma unroll
        for (int k = 0; k < NSIZE; k++) {
            if (!((mask >> k) & 1u)) continue;
            // MatSetValues(ADD) reads the column-major block row-major: entry (row k, col j) += Klocal(j, k),
            // Klocal(j,k) = af*(b1(j)*(kx*dNx(k)) + b2(j)*(ky*dNy(k)) [+ b3(j)*(kz*dNz(k))])   (poisson.F:93-95,183-187)
            double pk[NDIM];
#pragma unroll
            for (int d = 0; d < NDIM; d++) pk[d] = UNIT ? op.g.dN[d][k] : tiled_kcoef(prm, d) * op.g.dN[d][k];
            double kc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int j = 0; j < NSIZE; j++) {
                double s = bd[0][j] * pk[0] + bd[1][j] * pk[1];
                if (NDIM == 3) s = s + bd[NDIM - 1][j] * pk[NDIM - 1];
                kc[j] = UNIT ? s : prm.af * s;
            }
            // Flocal(k) with valC = 0, then lifting in ascending Dirichlet local index: F_k -= Klocal(k, ii) * g_ii
            double f = (op.g.N[k] * op.dvol) * prm.force;
            if (te.x < 0) {
#pragma unroll
                for (int ii = 0; ii < NSIZE; ii++) {
                    if (!fixed[ii]) continue;
                    double s = bd[0][k] * (tiled_kcoef(prm, 0) * op.g.dN[0][ii]) + bd[1][k] * (tiled_kcoef(prm, 1) * op.g.dN[1][ii]);
                    if (NDIM == 3) s = s + bd[NDIM - 1][k] * (tiled_kcoef(prm, NDIM - 1) * op.g.dN[NDIM - 1][ii]);
                    f = f - (prm.af * s) * gval[ii];
                }
            }
            if (neg) { kc[0] = kc[1] = kc[2] = kc[3] = 0.0; f = 0.0; }
            emit(col, kc, f, pre);
            col++;
        }
    }
}

template <int KIND, int THREADS, int MINB, bool UNIT>
__global__ void __launch_bounds__(THREADS, MINB) assemble_tiled_kernel(const TiledArgs a)
{
    using T = ElemTraits<KIND>;
    constexpr int NPE = T::NPE, NSIZE = NPE;
    static_assert(T::NDOF == 1, "tiled value pass: one dof per node");
    constexpr int NWARPS = THREADS / 32;

    PFEM_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int *td = a.tdesc + (size_t)blockIdx.x * TILE_DESC_INTS;
    const int row_off = td[TD_ROW_OFF], nrows_pad = td[TD_NROWS_PAD], el_off = td[TD_EL_OFF], nel = td[TD_NEL];
    const int slice0 = td[TD_SLICE0], nnz = td[TD_NNZ], ncols = td[TD_NCOLS];
    const int nlines = (ncols + 3) >> 2;
    double *Kst = reinterpret_cast<double *>(smem_raw);      // [nlines][16]: Klocal(0..NSIZE-1, k) of the staged columns (swizzled)
    double *Fst = Kst + (size_t)nlines * 16;                 // [4][nlines] : lifted Flocal(k)
    double *acc = Fst + (size_t)nlines * 4;                  // [nnz]       : the tile's CSR values, tile-row order
    double *sink = acc + nnz;                                // [THREADS] : Dirichlet columns land here (never used)

    // ---- accumulators: zero, or the current values when the matrix was not zeroed since the last pass ----
    sink[tid] = 0.0;
    if (!a.load_val) {
        for (int q = tid; q < nnz; q += THREADS) acc[q] = 0.0;
    } else {
        for (int i = warp; i < nrows_pad; i += NWARPS) {
            const int4 tr = __ldg(a.trows + row_off + i);
            if (tr.x < 0) continue;
            for (int j = lane; j < tr.w; j += 32) acc[tr.y + j] = a.val[tr.z + j];
        }
    }
    // the row threads' own stream heads: issued now, consumed in phase B (in flight during phase A)
    int4 my_tr = make_int4(-1, 0, 0, 0);
    long long my_o0 = 0;
    int my_width = 0;
    if (tid < nrows_pad) {
        my_tr = __ldg(a.trows + row_off + tid);
        my_o0 = a.tslice_off[slice0 + warp];
        my_width = (int)((a.tslice_off[slice0 + warp + 1] - my_o0) >> 5);
    }

    // ---- phase A: every element of the tile once ----
    tiled_phase_a<KIND, THREADS, UNIT>(a, el_off, nel, [](int) { return 0; }, [&](int col, const double (&kc)[4], double f, int &) {
        *reinterpret_cast<double2 *>(Kst + tiled_kst_off(col, 0)) = make_double2(kc[0], kc[1]);
        *reinterpret_cast<double2 *>(Kst + tiled_kst_off(col, 1)) = make_double2(kc[2], kc[3]);
        Fst[tiled_fst_off(col, nlines)] = f;
    });
    // ---- phase B: one thread per tile row gathers its staged columns in ascending element id ----
    // The incidence entries are streamed in batches of NB per row, the next batch in flight while the current one is
    // consumed (an entry costs ~60 cycles of shared-memory work, far less than one HBM latency); the first batch is
    // issued before the barrier.
    constexpr int NB = 8;
    const int2 *ip = a.tinc + my_o0 + lane;
    auto load_entry = [&](int m) { return m < my_width ? __ldcs(ip + (size_t)m * 32) : make_int2(-1, 0); };
    int2 buf[NB];
#pragma unroll
    for (int q = 0; q < NB; q++) buf[q] = load_entry(q);
    __syncthreads();
    if (tid < nrows_pad) {
        const int4 tr = my_tr;
        const bool live = tr.x >= 0;
        double *racc = acc + tr.y;
        double *dummy = sink + tid;
        double facc = (live && a.load_rhs) ? a.rhs[tr.x] : 0.0;
        for (int m0 = 0; m0 < my_width; m0 += NB) {
            int2 nxt[NB];
#pragma unroll
            for (int q = 0; q < NB; q++) nxt[q] = load_entry(m0 + NB + q);
#pragma unroll
            for (int q = 0; q < NB; q++) {
                const int2 cur = buf[q];
                if (cur.x >= 0) {                              // (< 0: slice padding)
                    const double2 k01 = *reinterpret_cast<const double2 *>(Kst + tiled_kst_off(cur.x, 0));
                    const double2 k23 = *reinterpret_cast<const double2 *>(Kst + tiled_kst_off(cur.x, 1));
                    const double kc[4] = {k01.x, k01.y, k23.x, k23.y};
                    const unsigned int sw = (unsigned int)cur.y;
                    // the free dofs of an element are distinct columns of the row, so the NSIZE read-modify-writes of
                    // one incidence are independent: issue every load before the first store (Dirichlet columns share
                    // the sink, whose value is never used)
                    double *dst[NSIZE];
                    double cur_v[NSIZE];
#pragma unroll
                    for (int j = 0; j < NSIZE; j++) {
                        const unsigned int sl = (sw >> (8 * j)) & 255u;
                        dst[j] = sl == 255u ? dummy : racc + sl;
                    }
                    const double fk = Fst[tiled_fst_off(cur.x, nlines)];
#pragma unroll
                    for (int j = 0; j < NSIZE; j++) cur_v[j] = *dst[j];
#pragma unroll
                    for (int j = 0; j < NSIZE; j++) *dst[j] = cur_v[j] + kc[j];
                    facc = facc + fk;                          // VecSetValues(ADD)
                }
            }
#pragma unroll
            for (int q = 0; q < NB; q++) buf[q] = nxt[q];
        }
        if (live) a.rhs[tr.x] = facc;
    }
    __syncthreads();

    // ---- phase C: accumulators -> CSR values, one warp per row (rows of a tile are runs of consecutive rows) ----
    // (the row descriptors carry rowptr[row] and the row length: no dependent global load per row)
#pragma unroll 4
    for (int i = warp; i < nrows_pad; i += NWARPS) {
        const int4 tr = __ldg(a.trows + row_off + i);
        if (tr.x < 0) continue;
        for (int j = lane; j < tr.w; j += 32) a.val[tr.z + j] = acc[tr.y + j];
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Scatter variant ("deterministic segmented reduction by slot").  Phase A is the same; instead of staging whole columns
// for the rows to gather, every contribution is stored straight at its final position in a run-ordered buffer: the
// contributions of one row lie at positions 0,1,2,... in the order (slot 0: ascending element), (slot 1: ...), ...,
// (Flocal: ascending element), with position p of the row in lane l of a slice at  slice base + p*33 + l  (33 = 32 lanes
// + 1: stores of one row to different positions and loads of different rows at one position are both bank-conflict
// free).  Phase B is then a plain run sum per row: no slot decode, no read-modify-write chains, half the shared-memory
// traffic of the gather.  Same per-entry summation order => same bits.
// ------------------------------------------------------------------------------------------------------------------
template <int KIND, int THREADS, int MINB, bool UNIT>
__global__ void __launch_bounds__(THREADS, MINB) assemble_tiled2_kernel(const TiledArgs a)
{
    using T = ElemTraits<KIND>;
    constexpr int NPE = T::NPE, NSIZE = NPE;
    static_assert(T::NDOF == 1, "tiled value pass: one dof per node");
    constexpr int NWARPS = THREADS / 32;

    PFEM_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int *td = a.tdesc + (size_t)blockIdx.x * TILE_DESC_INTS;
    const int row_off = td[TD_ROW_OFF], nrows_pad = td[TD_NROWS_PAD], el_off = td[TD_EL_OFF], nel = td[TD_NEL];
    const int slice0 = td[TD_SLICE0], crec_off = td[TD2_CREC_OFF], cbuf = td[TD2_CBUF];
    double *Cb = reinterpret_cast<double *>(smem_raw);       // [cbuf]: run-ordered contributions
    double *acc = Cb + cbuf;                                 // [td[TD_NNZ]]: the tile's CSR values, tile-row order

    // the row threads' descriptors and first run-length chunk: issued now, consumed in phase B
    int4 my_tr = make_int4(-1, 0, 0, 0), my_s2 = make_int4(0, 0, 0, 0);
    uint4 my_cnt = make_uint4(0u, 0u, 0u, 0u);
    if (tid < nrows_pad) {
        my_tr = __ldg(a.trows + row_off + tid);
        my_s2 = __ldg(a.ts2 + slice0 + warp);
        if (my_s2.z > 0) my_cnt = __ldcs(a.cnt + (size_t)my_s2.y + lane);
    }

    // ---- phase A: every element of the tile once; contributions go straight to their run positions ----
    const int2 *crec = a.crec + crec_off;
    struct Recs { int2 r[NSIZE]; };
    tiled_phase_a<KIND, THREADS, UNIT>(a, el_off, nel,
        [&](int base) {                                      // the element's column records (at most NSIZE; the array is padded)
            Recs q;
#pragma unroll
            for (int j = 0; j < NSIZE; j++) q.r[j] = __ldg(crec + base + j);
            return q;
        },
        [&](int, const double (&kc)[4], double f, Recs &q) {
            const unsigned int u0 = (unsigned int)q.r[0].x, u1 = (unsigned int)q.r[0].y;
#pragma unroll
            for (int j = 0; j + 1 < NSIZE; j++) q.r[j] = q.r[j + 1];            // next column's record moves to the front
            double *rowp = Cb + (u0 & 0xffffu);
            const unsigned int p[4] = {(u0 >> 16) & 255u, u0 >> 24, u1 & 255u, (u1 >> 8) & 255u};
#pragma unroll
            for (int j = 0; j < NSIZE; j++)
                if (p[j] != 255u) rowp[p[j] * TILE_CB_STRIDE] = kc[j];         // 255: Dirichlet column, dropped
            rowp[((u1 >> 16) & 255u) * TILE_CB_STRIDE] = f;
        });
    __syncthreads();

    // ---- phase B: one thread per tile row sums its runs in order ----
    if (tid < nrows_pad && my_tr.x >= 0) {
        const int4 tr = my_tr;
        const double *src = Cb + my_s2.x + lane;
        const uint4 *cp = a.cnt + (size_t)my_s2.y + lane;
        const int len = tr.w, nchunks = (len + 1 + 15) >> 4;
        uint4 cur = my_cnt;
        for (int c = 0; c < nchunks; c++) {
            const uint4 nxt = c + 1 < nchunks ? __ldcs(cp + (size_t)(c + 1) * 32) : make_uint4(0u, 0u, 0u, 0u);
            const int q1 = min(16, len + 1 - 16 * c);
#pragma unroll 1
            for (int q = 0; q < q1; q++) {                                     // compact code: 16 copies would not fit the i-cache
                const int slot = 16 * c + q;
                const unsigned int word = q < 8 ? (q < 4 ? cur.x : cur.y) : (q < 12 ? cur.z : cur.w);
                const int n = (int)((word >> (8 * (q & 3))) & 255u);
                double sum;
                if (slot < len) sum = a.load_val ? a.val[tr.z + slot] : 0.0;
                else sum = a.load_rhs ? a.rhs[tr.x] : 0.0;
#pragma unroll 4
                for (int i = 0; i < n; i++) { sum = sum + *src; src += TILE_CB_STRIDE; }
                if (slot < len) acc[tr.y + slot] = sum;
                else a.rhs[tr.x] = sum;                                        // VecSetValues(ADD)
            }
            cur = nxt;
        }
    }
    __syncthreads();

    // ---- phase C: accumulators -> CSR values, one warp per row ----
#pragma unroll 4
    for (int i = warp; i < nrows_pad; i += NWARPS) {
        const int4 tr = __ldg(a.trows + row_off + i);
        if (tr.x < 0) continue;
        for (int j = lane; j < tr.w; j += 32) a.val[tr.z + j] = acc[tr.y + j];
    }
}

}  // namespace pfem
