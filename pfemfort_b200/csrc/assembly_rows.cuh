// assembly_rows.cuh -- the row-gather value-pass kernels (the default value pass; host side: assembly.cu).
//
// Replaces the element loop of the *parallelimpl1 drivers (tetrapoissonparallelimpl1.F:828-884,
// tetraelasticityparallelimpl1.F:901-968) and PETSc's MatSetValues/VecSetValues behind it; see assembly.cu for the design.
// Kept in a header so that tests/emu can compile the very same kernel source for the host (PFEM_EMULATE + CUDA shim)
// and check it against the oracle in the CPU test suite; inline PTX is confined to ld_xyz.
#pragma once
#include "elements.cuh"

#ifndef PFEM_DYN_SMEM
#define PFEM_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace pfem {

struct AsmArgs {
    int nloc, row_lo, row_hi, rec_ints;
    const int *erec;
    const double *xyz;
    const double *applied;
    const int *rowptr, *col;
    double *val, *rhs;
    const int *rinc_ptr, *rinc;
    const double *elemData, *timeData;
    int *neg_count;
    int load_val, load_rhs;
    int max_seg_nnz;     // smem carve-up
    const long long *ainc_off;
    const int *ainc;
    const int *conn4;
    int *neg_flag;
    int unit;            // kx = ky = kz = af = 1.0 exactly
};

template <int NPE, int NDIM>
__device__ __forceinline__ void load_coords(const double *__restrict__ xyz, const int nodes[NPE], double x[NPE],
                                            double y[NPE], double z[NPE])
{
#pragma unroll
    for (int i = 0; i < NPE; i++) {
        if (NDIM == 3) {
            const double2 *p = reinterpret_cast<const double2 *>(xyz + (size_t)nodes[i] * 4);
            const double2 a = __ldg(p), b = __ldg(p + 1);
            x[i] = a.x; y[i] = a.y; z[i] = b.x;
        } else {
            const double2 a = __ldg(reinterpret_cast<const double2 *>(xyz + (size_t)nodes[i] * 2));
            x[i] = a.x; y[i] = a.y; z[i] = 0.0;
        }
    }
}

template <int KIND, int R>
__global__ void __launch_bounds__(R) assemble_kernel(const AsmArgs a)
{
    using T = ElemTraits<KIND>;
    constexpr int NPE = T::NPE, NDOF = T::NDOF, NDIM = T::NDIM, NSIZE = NPE * NDOF;
    constexpr int REC4 = (NPE + NSIZE + 3) / 4;

    PFEM_DYN_SMEM(smem_raw);
    double *acc = reinterpret_cast<double *>(smem_raw);
    int *scol = reinterpret_cast<int *>(acc + a.max_seg_nnz);
    int *sinc = scol + a.max_seg_nnz;

    const int r0 = blockIdx.x * R;
    const int rend = min(r0 + R, a.nloc);
    const int nnz0 = a.rowptr[r0], nnz1 = a.rowptr[rend];
    const int inc0 = a.rinc_ptr[r0], inc1 = a.rinc_ptr[rend];

    // stage the CTA's CSR segment and incidence segment through shared memory (coalesced)
    for (int k = threadIdx.x; k < nnz1 - nnz0; k += R) {
        scol[k] = a.col[nnz0 + k];
        acc[k] = a.load_val ? a.val[nnz0 + k] : 0.0;
    }
    for (int m = threadIdx.x; m < inc1 - inc0; m += R) sinc[m] = a.rinc[inc0 + m];
    __syncthreads();

    const int r = r0 + threadIdx.x;
    if (r < rend) {
        Params<KIND> prm;
        prm.init(a.elemData, a.timeData);
        const int seg = a.rowptr[r] - nnz0, len = a.rowptr[r + 1] - a.rowptr[r];
        const int *rc = scol + seg;
        double *racc = acc + seg;
        double facc = a.load_rhs ? a.rhs[r] : 0.0;
        const double du0[3] = {0.0, 0.0, 0.0};    // valC = 0 in the drivers (tetrapoissonparallelimpl1.F:824)
        const int m1 = a.rinc_ptr[r + 1] - inc0;
        for (int m = a.rinc_ptr[r] - inc0; m < m1; m++) {
            const int code = sinc[m];
            const int e = code / NSIZE, k = code - e * NSIZE;
            int rec[REC4 * 4];
            const int4 *rp = reinterpret_cast<const int4 *>(a.erec + (size_t)e * a.rec_ints);
#pragma unroll
            for (int q = 0; q < REC4; q++) {
                const int4 v = __ldg(rp + q);
                rec[4 * q] = v.x; rec[4 * q + 1] = v.y; rec[4 * q + 2] = v.z; rec[4 * q + 3] = v.w;
            }
            const int *nodes = rec, *dofs = rec + NPE;
            double x[NPE], y[NPE], z[NPE];
            load_coords<NPE, NDIM>(a.xyz, nodes, x, y, z);
            ElemOp<KIND> op;
            op.load_geom(x, y, z);
            if (op.g.Jac < 0.0) {
                // the reference STOPs here; report once per element (from its first owned local dof)
                int first = 0;
#pragma unroll
                for (int q = NSIZE - 1; q >= 0; q--)
                    if (dofs[q] >= a.row_lo && dofs[q] < a.row_hi) first = q;
                if (first == k) atomicAdd(a.neg_count, 1);
                continue;
            }
            op.set_dvol(prm);
            // MatSetValues(ADD): entry (row k, col j) += Klocal(j, k)
            op.col_setup(prm, k);
#pragma unroll
            for (int j = 0; j < NSIZE; j++) {
                const int c = dofs[j];
                if (c < 0) continue;
                const double v = op.K(prm, j);
                int lo = 0, hi = len;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (rc[mid] < c) lo = mid + 1; else hi = mid;
                }
                racc[lo] = racc[lo] + v;
            }
            // Flocal(k), then lifting in ascending Dirichlet local index: F_k -= Klocal(k, ii) * g_ii
            double f = op.F(prm, k, du0);
#pragma unroll
            for (int ii = 0; ii < NSIZE; ii++) {
                if (dofs[ii] != -1) continue;
                const double gval = a.applied[(size_t)nodes[ii / NDOF] * NDOF + (ii % NDOF)];
                op.col_setup(prm, ii);
                f = f - op.K(prm, k) * gval;
            }
            facc = facc + f;      // VecSetValues(ADD)
        }
        a.rhs[r] = facc;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nnz1 - nnz0; k += R) a.val[nnz0 + k] = acc[k];
}

// 256-bit read-only load of one node's (x, y, z, pad)
__device__ __forceinline__ void ld_xyz(const double *p, double &x, double &y, double &z)
{
#if defined(__CUDA_ARCH__)
    double w;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(p));
#else
    x = p[0]; y = p[1]; z = p[2];
#endif
}

// Streamed row-gather value pass.  Same arithmetic and summation order as assemble_kernel, but every input is a
// coalesced stream: a warp owns a 32-row slice and reads its incidence entries column-major ({code, slot bytes},
// slots precomputed at pattern time: no search), the element record is one int4 of node ids, coordinates are 256-bit
// loads, and only the FP64 accumulators of the CTA's contiguous CSR segment live in shared memory.
// The loop is software-pipelined: incidence entries three iterations ahead, node ids two ahead, and the next
// iteration's coordinates already in flight into registers while the current element is computed.
// UNIT: kx = ky = kz = af = 1.0 exactly (the drivers' constants): multiplications by 1.0 are skipped (bit-identical).
// OP: the element operator -- ElemOp<KIND> (reference evaluation order; its translation unit is built with -fmad=false) or
// FastOp<KIND> (elements_fast.cuh: FMA-contracted cofactor form; 1e-12 contract).
template <int KIND, int R, bool UNIT, class OP = ElemOp<KIND>>
__global__ void __launch_bounds__(R) assemble_sell_kernel(const AsmArgs a)
{
    using T = ElemTraits<KIND>;
    constexpr int NPE = T::NPE, NDOF = T::NDOF, NDIM = T::NDIM, NSIZE = NPE * NDOF;
    constexpr int WORDS = NSIZE <= 4 ? 2 : 4;
    constexpr int NW = (NSIZE + 3) / 4;          // slot words in use
    constexpr int XYZ = NDIM == 3 ? 4 : 2;       // doubles per node record

    PFEM_DYN_SMEM(smem_raw);
    double *acc = reinterpret_cast<double *>(smem_raw);

    const int r0 = blockIdx.x * R;
    const int rend = min(r0 + R, a.nloc);
    const int nnz0 = a.rowptr[r0], nnz1 = a.rowptr[rend];
    for (int k = threadIdx.x; k < nnz1 - nnz0; k += R) acc[k] = a.load_val ? a.val[nnz0 + k] : 0.0;
    __syncthreads();

    const int r = r0 + threadIdx.x;
    const int slice = r >> 5, lane = threadIdx.x & 31;
    if (slice * 32 < a.nloc) {
        Params<KIND> prm;
        prm.init(a.elemData, a.timeData);
        const bool live = r < rend;
        double *racc = acc + (live ? a.rowptr[r] - nnz0 : 0);
        double *dummy = acc + a.max_seg_nnz + threadIdx.x;      // sink for Dirichlet columns (never read)
        double facc = (live && a.load_rhs) ? a.rhs[r] : 0.0;
        const long long o0 = a.ainc_off[slice];
        const int width = (int)((a.ainc_off[slice + 1] - o0) >> 5);
        const int *ip = a.ainc + (size_t)(o0 + lane) * WORDS;
        const int4 *conn = reinterpret_cast<const int4 *>(a.conn4);

        struct Ent { int code; unsigned int sw[3]; };
        auto load_entry = [&](int m) {
            Ent t;
            t.code = -1; t.sw[0] = t.sw[1] = t.sw[2] = 0u;
            if (m < width) {
                const int *q = ip + (size_t)m * 32 * WORDS;
                if (WORDS == 2) {
                    const int2 v = __ldcs(reinterpret_cast<const int2 *>(q));
                    t.code = v.x; t.sw[0] = (unsigned int)v.y;
                } else {
                    const int4 v = __ldcs(reinterpret_cast<const int4 *>(q));
                    t.code = v.x; t.sw[0] = (unsigned int)v.y; t.sw[1] = (unsigned int)v.z; t.sw[2] = (unsigned int)v.w;
                }
            }
            return t;
        };
        auto load_conn = [&](const Ent &t) { return t.code >= 0 ? __ldg(conn + t.code / NSIZE) : make_int4(0, 0, 0, 0); };
        auto load_xyz = [&](const int4 &c, double (&x)[NPE], double (&y)[NPE], double (&z)[NPE]) {
            const int nd[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int i = 0; i < NPE; i++) {
                if (NDIM == 3) ld_xyz(a.xyz + (size_t)nd[i] * 4, x[i], y[i], z[i]);
                else {
                    const double2 t = __ldg(reinterpret_cast<const double2 *>(a.xyz + (size_t)nd[i] * 2));
                    x[i] = t.x; y[i] = t.y; z[i] = 0.0;
                }
            }
        };
        Ent e0 = load_entry(0), e1 = load_entry(1), e2 = load_entry(2);
        int4 c0 = load_conn(e0), c1 = load_conn(e1);
        double xq[NPE], yq[NPE], zq[NPE];                  // coordinates of the element of the NEXT iteration
        load_xyz(c0, xq, yq, zq);
#pragma unroll 2      // two copies: the coordinate / node-id register sets ping-pong instead of being copied (unroll 6: slower)
        for (int m = 0; m < width; m++) {
            const Ent e3 = load_entry(m + 3);
            const int4 c2 = load_conn(e2);
            double x[NPE], y[NPE], z[NPE];
#pragma unroll
            for (int i = 0; i < NPE; i++) { x[i] = xq[i]; y[i] = yq[i]; z[i] = zq[i]; }
            load_xyz(c1, xq, yq, zq);                      // in flight during this iteration's arithmetic
            const Ent cur = e0;
            const int4 cn = c0;
            e0 = e1; e1 = e2; e2 = e3; c0 = c1; c1 = c2;
            if (cur.code < 0) continue;                    // slice padding
            const int k = cur.code % NSIZE;
            const int nodes[4] = {cn.x, cn.y, cn.z, cn.w};
            OP op;
            op.load_geom(x, y, z);
            if (op.g.Jac < 0.0) { atomicOr(a.neg_flag, 1); continue; }   // the reference STOPs here
            op.set_dvol(prm);
            // MatSetValues(ADD): entry (row k, col j) += Klocal(j, k); Dirichlet columns (slot byte 0xFF) are dropped
            if (UNIT) op.col_setup_unit(k); else op.col_setup(prm, k);
#pragma unroll
            for (int j = 0; j < NSIZE; j++) {
                const unsigned int sl = (cur.sw[j >> 2] >> (8 * (j & 3))) & 255u;
                double *dst = sl == 255u ? dummy : racc + sl;
                *dst = *dst + (UNIT ? op.K_unit(j) : op.K(prm, j));
            }
            // Flocal(k), then lifting in ascending Dirichlet local index: F_k -= Klocal(k, ii) * g_ii
            double f = op.F0(prm, k);
            bool any_dbc = false;
#pragma unroll
            for (int q = 0; q < NW; q++) {
                const unsigned int v = ~cur.sw[q];         // a 0xFF slot byte becomes a zero byte (unused bytes hold 0x00)
                any_dbc |= ((v - 0x01010101u) & ~v & 0x80808080u) != 0u;
            }
            if (any_dbc) {
#pragma unroll
                for (int ii = 0; ii < NSIZE; ii++) {
                    const unsigned int sl = (cur.sw[ii >> 2] >> (8 * (ii & 3))) & 255u;
                    if (sl != 255u) continue;
                    const double gval = a.applied[(size_t)nodes[ii / NDOF] * NDOF + (ii % NDOF)];
                    op.col_setup(prm, ii);
                    f = f - op.K(prm, k) * gval;
                }
            }
            facc = facc + f;      // VecSetValues(ADD)
        }
        if (live) a.rhs[r] = facc;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nnz1 - nnz0; k += R) a.val[nnz0 + k] = acc[k];
}

}  // namespace pfem
