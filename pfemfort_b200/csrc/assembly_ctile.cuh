// assembly_ctile.cuh -- the colour-scheduled tile value pass (default for the one-dof-per-node kinds).
//
// Same job as the row-gather kernels (assembly_rows.cuh): the element loop of tetrapoissonparallelimpl1.F:828-884 /
// triapoissonparallelimpl1.F:849-905 with PETSc's MatSetValues(ADD) / VecSetValues(ADD) and the Dirichlet lifting behind it,
// fused into one kernel without atomics.  What differs is the schedule:
//
//   * a persistent CTA (one per SM) owns a TILE of up to ~1000 spatially close matrix rows (Morton order of the rows'
//     nodes, built on the GPU at pattern time: assembly_ctile.cu).  The tile's FP64 accumulators (its CSR value
//     segment, fixed stride per row, plus its RHS entries) live in shared memory for the whole tile; the coordinates
//     and applied values of the tile's nodes (own + halo) are staged into shared memory by ONE bulk asynchronous copy
//     (cp.async.bulk, completion on an mbarrier), the next tile's table is prefetched into L2 meanwhile;
//   * every element that touches the tile is computed ONCE per tile (1.3-1.4 visits per element instead of one
//     geometry evaluation per incident row = 4 per tetrahedron), by one thread, from shared-memory coordinates;
//   * the tile's element visits are grouped into ROUNDS at pattern time (greedy colouring, assembly_ctile.cu) such that
//     within a round no two visits have the same owned row at the same local position k.  A round commits column k of
//     every visit with plain shared-memory read-modify-writes (no two threads touch the same row => no atomics), then
//     __syncthreads, then k+1.  The order in which contributions reach a matrix entry is therefore FIXED by the
//     schedule: run-to-run deterministic, independent of the SM count -- but it is not the sequential element order of
//     the reference, and the element arithmetic below uses FMAs and the symmetric form of Klocal, so results agree with
//     the sequential no-FMA evaluation to rounding (tested at 1e-12 relative), not bit for bit.  PFEM_ASM=rows selects
//     the bit-identical row-gather kernels instead.
//
// Element arithmetic (P1, one Gauss point): with the edge vectors from local node 3 (tet: basisfuncs.F:493-509) the
// unnormalised gradient of N_i is the cofactor vector c_i, grad N_i = c_i / Jac, so
//   Klocal(i,j) = af * dvol * sum_d k_d gradN_i[d] gradN_j[d] = (af * gw / Jac) * sum_d k_d c_i[d] c_j[d]
// (elementutilitiespoisson.F:87-95,177-187 with dvol = gw * Jac), Flocal(i) = N_i * dvol * force (valC = 0 in the
// drivers), followed by the lifting F_i -= Klocal(i,j) * g_j over the Dirichlet local nodes j
// (tetrapoissonparallelimpl1.F:856-872).
#pragma once
#include "elements.cuh"

namespace pfem {

constexpr int CT_DESC = 8;     // ints per tile descriptor
enum { CT_ROW0 = 0, CT_NROWS = 1, CT_NODE0 = 2, CT_NNODES = 3, CT_ROUND0 = 4, CT_NROUNDS = 5, CT_STRIDE = 6, CT_NVISITS = 7 };
constexpr int CT_MAX_ROUNDS = 64;

struct CtileArgs {
    int ntiles;
    const int *tdesc;          // [ntiles][CT_DESC]
    const int4 *trow;          // per tile row (tile order): { local row, rowptr[row], row length, 0 }
    const double4 *tnode;      // per-tile node tables { x, y, z, applied value }, own rows' nodes first
    const int *round_off;      // per tile NROUNDS+1 absolute offsets into the visit arrays
    const uint2 *vnode;        // per visit: 4 x u16 tile-local node ids (local dof k -> node)
    const uint4 *vslot;        // per visit: one slot word per local dof k (4 slot bytes; 0xFFFFFFFF: row k not in this tile)
    double *val, *rhs;
    const double *elemData, *timeData;
    int *neg_flag;
    int load_val, load_rhs;
    int node_cap;              // shared-memory carve-up: nodes (double4), row descriptors (int4), accumulators, RHS, sinks
    int row_cap;
    int acc_cap;
};

// ---- Blackwell/Hopper async-copy plumbing (inline PTX; SASS: UBLKCP / SYNCS) ------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src_gmem, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- element arithmetic of the tile kernel ------------------------------------------------------------------------------
// K(i,j), i <= j, in packed upper-triangular order; F(i) before lifting.
template <int KIND> struct TileElem;

template <> struct TileElem<POISSON_TETRA> {
    static constexpr int NPE = 4, NK = 10;
    double K[NK], F0, Jac;
    __device__ __forceinline__ static int idx(int i, int j) { return i <= j ? (i * (9 - i)) / 2 + (j - i) : (j * (9 - j)) / 2 + (i - j); }
    __device__ __forceinline__ void compute(const double4 (&P)[4], const Params<POISSON_TETRA> &p, bool unit)
    {
        // rows of B: nodes 0, 1, 3 relative to node 2 (basisfuncs.F:493-509)
        const double ax = P[0].x - P[2].x, ay = P[0].y - P[2].y, az = P[0].z - P[2].z;
        const double bx = P[1].x - P[2].x, by = P[1].y - P[2].y, bz = P[1].z - P[2].z;
        const double cx = P[3].x - P[2].x, cy = P[3].y - P[2].y, cz = P[3].z - P[2].z;
        double g[3][4];
        g[0][0] = by * cz - bz * cy; g[1][0] = bz * cx - bx * cz; g[2][0] = bx * cy - by * cx;   // b x c
        g[0][1] = cy * az - cz * ay; g[1][1] = cz * ax - cx * az; g[2][1] = cx * ay - cy * ax;   // c x a
        g[0][3] = ay * bz - az * by; g[1][3] = az * bx - ax * bz; g[2][3] = ax * by - ay * bx;   // a x b
#pragma unroll
        for (int d = 0; d < 3; d++) g[d][2] = -((g[d][0] + g[d][1]) + g[d][3]);
        Jac = ax * g[0][0] + ay * g[1][0] + az * g[2][0];
        const double s = (p.af * p.gw) / Jac;
        double h[3][4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            h[0][j] = unit ? g[0][j] : p.kx * g[0][j];
            h[1][j] = unit ? g[1][j] : p.ky * g[1][j];
            h[2][j] = unit ? g[2][j] : p.kz * g[2][j];
        }
        int q = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = i; j < 4; j++) K[q++] = s * (g[0][i] * h[0][j] + g[1][i] * h[1][j] + g[2][i] * h[2][j]);
        F0 = (0.25 * (p.gw * Jac)) * p.force;            // N_i = 1/4 at the Gauss point for every i
    }
    __device__ __forceinline__ double Fk(int) const { return F0; }
};

template <> struct TileElem<POISSON_TRIA> {
    static constexpr int NPE = 3, NK = 6;
    double K[NK], F0, Jac;
    __device__ __forceinline__ static int idx(int i, int j) { return i <= j ? (i * (7 - i)) / 2 + (j - i) : (j * (7 - j)) / 2 + (i - j); }
    __device__ __forceinline__ void compute(const double4 (&P)[4], const Params<POISSON_TRIA> &p, bool unit)
    {
        const double ax = P[1].x - P[0].x, ay = P[1].y - P[0].y;      // basisfuncs.F:208-217
        const double bx = P[2].x - P[0].x, by = P[2].y - P[0].y;
        Jac = ax * by - ay * bx;
        double g[2][3];
        g[0][1] = by;  g[1][1] = -bx;
        g[0][2] = -ay; g[1][2] = ax;
        g[0][0] = -(g[0][1] + g[0][2]); g[1][0] = -(g[1][1] + g[1][2]);
        const double s = (p.af * p.gw) / Jac;
        double h[2][3];
#pragma unroll
        for (int j = 0; j < 3; j++) { h[0][j] = unit ? g[0][j] : p.kx * g[0][j]; h[1][j] = unit ? g[1][j] : p.ky * g[1][j]; }
        int q = 0;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = i; j < 3; j++) K[q++] = s * (g[0][i] * h[0][j] + g[1][i] * h[1][j]);
        F0 = (p.gw * Jac) * p.force;                      // times N_i below
    }
    __device__ __forceinline__ double Fk(int i) const
    {
        const double xi = third_f();
        return (i == 0 ? 1.0 - xi - xi : xi) * F0;
    }
};

// SUBSYNC = true : the rounds were built with the per-position rule (no two visits of a round share an owned row at the
//                   same local position k): column k of every visit is committed, then __syncthreads, then k + 1.
// SUBSYNC = false: the rounds were built with the full rule (no two visits of a round share an owned row at all): every
//                   visit commits all its columns at once, one __syncthreads per chunk.
template <int KIND, int B, bool SUBSYNC>
__global__ void __launch_bounds__(B, 1) assemble_ctile_kernel(const CtileArgs a)
{
    using E = TileElem<KIND>;
    constexpr int NPE = E::NPE;
    extern __shared__ __align__(128) unsigned char ct_smem[];
    double4 *snode = reinterpret_cast<double4 *>(ct_smem);
    int4 *strow = reinterpret_cast<int4 *>(snode + a.node_cap);
    double *acc = reinterpret_cast<double *>(strow + a.row_cap);
    double *frhs = acc + a.acc_cap;
    double *sink = frhs + a.row_cap + threadIdx.x;          // Dirichlet columns are added here (never read back)
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ int s_roff[CT_MAX_ROUNDS + 2];

    const int tid = threadIdx.x;
    Params<KIND> prm;
    prm.init(a.elemData, a.timeData);
    bool unit = prm.af == 1.0 && prm.kx == 1.0 && prm.ky == 1.0;
    if constexpr (KIND == POISSON_TETRA) unit = unit && prm.kz == 1.0;

    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    *sink = 0.0;
    __syncthreads();
    unsigned parity = 0;

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int *td = a.tdesc + (size_t)tile * CT_DESC;
        const int row0 = td[CT_ROW0], nrows = td[CT_NROWS], node0 = td[CT_NODE0], nnodes = td[CT_NNODES];
        const int round0 = td[CT_ROUND0], nrounds = td[CT_NROUNDS], stride = td[CT_STRIDE];
        // ---- stage the tile's node table and row descriptors (bulk async copies) while the accumulators are zeroed ----
        if (tid == 0) {
            fence_proxy_async();                 // the previous tile's generic-proxy reads of the staged tables precede these writes
            const unsigned nb = (unsigned)nnodes * 32u, rb = (unsigned)nrows * 16u;
            mbar_expect_tx(&mbar, nb + rb);
            bulk_g2s(snode, a.tnode + node0, nb, &mbar);
            bulk_g2s(strow, a.trow + row0, rb, &mbar);
            const int nt = tile + gridDim.x;
            if (nt < a.ntiles) {
                const int *nd = a.tdesc + (size_t)nt * CT_DESC;
                bulk_prefetch_l2(a.tnode + nd[CT_NODE0], (unsigned)nd[CT_NNODES] * 32u);
                bulk_prefetch_l2(a.trow + nd[CT_ROW0], (unsigned)nd[CT_NROWS] * 16u);
            }
        }
        for (int q = tid; q <= nrounds; q += B) s_roff[q] = a.round_off[round0 + q];
        const int nacc = nrows * stride;
        if (!a.load_val) for (int q = tid; q < nacc; q += B) acc[q] = 0.0;
        if (!a.load_rhs) for (int i = tid; i < nrows; i += B) frhs[i] = 0.0;
        mbar_wait(&mbar, parity);
        parity ^= 1u;
        if (a.load_val | a.load_rhs) {           // ADD on top of the current values (no setZero since the last pass)
            for (int i = tid >> 4; i < nrows; i += B >> 4) {
                const int4 tr = strow[i];
                if (a.load_val) for (int j = tid & 15; j < tr.z; j += 16) acc[i * stride + j] = a.val[tr.y + j];
                if (a.load_rhs && (tid & 15) == 0) frhs[i] = a.rhs[tr.x];
            }
        }
        __syncthreads();

        // ---- rounds ----
        int r = 0, base = s_roff[0], vend = nrounds > 0 ? s_roff[1] : base;
        uint2 vn = make_uint2(0u, 0u);
        uint4 vs = make_uint4(~0u, ~0u, ~0u, ~0u);
        if (nrounds > 0 && base + tid < vend) { vn = __ldcs(a.vnode + base + tid); vs = __ldcs(a.vslot + base + tid); }
        while (r < nrounds) {
            // next chunk: same round, or the head of the next one
            int nbase = base + B, nr = r, nend = vend;
            if (nbase >= vend) { nr = r + 1; nbase = vend; nend = nr < nrounds ? s_roff[nr + 1] : vend; }
            uint2 vn2 = make_uint2(0u, 0u);
            uint4 vs2 = make_uint4(~0u, ~0u, ~0u, ~0u);
            if (nr < nrounds && nbase + tid < nend) { vn2 = __ldcs(a.vnode + nbase + tid); vs2 = __ldcs(a.vslot + nbase + tid); }

            const unsigned w[4] = {vs.x, vs.y, vs.z, vs.w};
            const bool active = (vs.x & vs.y & vs.z & vs.w) != ~0u;
            E el;
            double gval[4] = {0.0, 0.0, 0.0, 0.0};
            if (active) {
                const unsigned nl[4] = {vn.x & 0xFFFFu, vn.x >> 16, vn.y & 0xFFFFu, vn.y >> 16};
                double4 P[4];
#pragma unroll
                for (int i = 0; i < NPE; i++) { P[i] = snode[nl[i]]; gval[i] = P[i].w; }
                if (NPE == 3) P[3] = P[2];
                el.compute(P, prm, unit);
                if (el.Jac < 0.0) atomicOr(a.neg_flag, 1);         // the reference STOPs here
            }
#pragma unroll
            for (int k = 0; k < NPE; k++) {
                if (active && w[k] != ~0u) {
                    const unsigned rl = (k < 2 ? (vn.x >> (16 * k)) : (vn.y >> (16 * (k - 2)))) & 0xFFFFu;
                    double *ra = acc + rl * stride;
                    // MatSetValues(ADD): entry (row k, col j) += Klocal(j, k) (= Klocal(k, j) in this symmetric form).
                    // The free columns of an element are distinct entries of the row: all loads first, then all stores.
                    // A Dirichlet column (slot byte 0xFF) lands in the sink and goes to the lifting instead.
                    double *dst[NPE];
                    double cur[NPE];
#pragma unroll
                    for (int j = 0; j < NPE; j++) {
                        const unsigned sl = (w[k] >> (8 * j)) & 255u;
                        dst[j] = sl == 255u ? sink : ra + sl;
                    }
#pragma unroll
                    for (int j = 0; j < NPE; j++) cur[j] = *dst[j];
                    const double fr = frhs[rl];
                    double f = el.Fk(k);
                    const unsigned nw = ~w[k];                     // a 0xFF slot byte becomes a zero byte
                    if (((nw - 0x01010101u) & ~nw & 0x80808080u) != 0u) {
                        // lifting: F_k -= Klocal(k, j) * g_j over the Dirichlet local nodes j (tetrapoissonparallelimpl1.F:856-872)
#pragma unroll
                        for (int j = 0; j < NPE; j++)
                            if (((w[k] >> (8 * j)) & 255u) == 255u) f = fma(-el.K[E::idx(k, j)], gval[j], f);
                    }
#pragma unroll
                    for (int j = 0; j < NPE; j++) *dst[j] = cur[j] + el.K[E::idx(k, j)];
                    frhs[rl] = fr + f;                              // VecSetValues(ADD)
                }
                if (SUBSYNC) __syncthreads();
            }
            if (!SUBSYNC) __syncthreads();
            vn = vn2; vs = vs2; base = nbase; r = nr; vend = nend;
        }

        // ---- write-out: accumulators -> CSR values / RHS, half a warp per row ----
        for (int i = tid >> 4; i < nrows; i += B >> 4) {
            const int4 tr = strow[i];
            for (int j = tid & 15; j < tr.z; j += 16) a.val[tr.y + j] = acc[i * stride + j];
            if ((tid & 15) == 0) a.rhs[tr.x] = frhs[i];
        }
        __syncthreads();
    }
}

}  // namespace pfem
