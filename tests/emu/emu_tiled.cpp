// emu_tiled.cpp -- TEST INFRASTRUCTURE ONLY: runs the tiled value-pass kernel (pfemfort_b200/csrc/assembly_tiled.cuh)
// and its host tile builder (tiles.hpp) on the CPU, CTA by CTA, through tests/emu/cuda_shim.h.
//
// The harness first restates, on the host, the device-side array formats the pattern pass produces (element records,
// row incidence lists, the row-gather incidence streams with their slot bytes: pattern.cu), hands them to the very
// same build_tiles() the product calls, and then executes the very same kernel source.  tests/test_tiled_emu.py
// compares the resulting CSR values / RHS bit for bit with the oracle.  Built by that test with g++.
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "cuda_shim.h"

thread_local uint3 threadIdx, blockIdx;
uint3 blockDim, gridDim;
EmuBarrier emu_barrier;
unsigned char *emu_smem = nullptr;

#include "assembly_tiled.cuh"

using namespace pfem;

template <int KIND, int THREADS, bool UNIT>
static void run_grid(const TiledArgs &args, int ntiles, size_t smem_bytes, int mode)
{
    std::vector<unsigned char> smem(smem_bytes + 64);
    unsigned char *base = smem.data();
    base += (16 - ((uintptr_t)base & 15)) & 15;
    emu_smem = base;
    std::memset(base, 0xFF, smem_bytes);          // NaN poison: a read of unstaged data shows up in the results
    emu_barrier.reset(THREADS);
    blockDim = uint3{(unsigned)THREADS, 1, 1};
    gridDim = uint3{(unsigned)ntiles, 1, 1};
    std::vector<std::thread> pool;
    for (int t = 0; t < THREADS; t++)
        pool.emplace_back([&, t] {
            threadIdx = uint3{(unsigned)t, 0, 0};
            for (int b = 0; b < ntiles; b++) {
                blockIdx = uint3{(unsigned)b, 0, 0};
                if (mode == 2) assemble_tiled2_kernel<KIND, THREADS, 1, UNIT>(args);
                else assemble_tiled_kernel<KIND, THREADS, 1, UNIT>(args);
                emu_barrier.wait();
                if (t == 0) std::memset(base, 0xFF, smem_bytes);
                emu_barrier.wait();
            }
        });
    for (auto &th : pool) th.join();
}

// conn0: [npe][nElem] 0-based NEW node ids; edof: [nsize][nElem] global dof ids (-1 Dirichlet); xyz: [ndim][nNode] NEW
// numbering; rowptr/col: CSR pattern of the owned rows [row_lo, row_lo+nloc) with global columns.
// stats: { ntiles, elem_visits, elems_touched, max_smem, neg_flag }
extern "C" int emu_assemble_tiled(int kind, int nElem, int nNode, const int *conn0, const int *edof, const double *xyz_soa,
                                  const double *applied, int row_lo, int nloc, const int *rowptr, const int *col,
                                  const double *elemData, const double *timeData, int tile_rows, int smem_budget, int threads,
                                  int load, double *val, double *rhs, long long *stats, int mode)
{
    if (kind != POISSON_TRIA && kind != POISSON_TETRA) return 2;
    const bool build_only = threads < 0;          // tile statistics only (large meshes: the emulated kernel is slow)
    if (build_only) threads = -threads;
    const int npe = kind == POISSON_TRIA ? 3 : 4, nsize = npe, ndim = kind == POISSON_TRIA ? 2 : 3;
    const int rec_ints = ((npe + nsize + 3) / 4) * 4, stride = ndim == 3 ? 4 : 2;
    // --- device formats, restated (pattern.cu: pack_conn/pack_dof, pack_xyz, conn4, inc sort, fill_asm_inc) ---
    std::vector<int> erec((size_t)nElem * rec_ints, -1), conn4((size_t)nElem * 4);
    for (int e = 0; e < nElem; e++) {
        for (int i = 0; i < npe; i++) erec[(size_t)e * rec_ints + i] = conn0[(size_t)i * nElem + e];
        for (int k = 0; k < nsize; k++) erec[(size_t)e * rec_ints + npe + k] = edof[(size_t)k * nElem + e];
        for (int i = 0; i < 4; i++) conn4[(size_t)e * 4 + i] = conn0[(size_t)(i < npe ? i : npe - 1) * nElem + e];
    }
    std::vector<double> xyz((size_t)nNode * stride, 0.0);
    for (int n = 0; n < nNode; n++)
        for (int d = 0; d < ndim; d++) xyz[(size_t)n * stride + d] = xyz_soa[(size_t)d * nNode + n];
    std::vector<int> rinc_ptr(nloc + 1, 0);
    for (int e = 0; e < nElem; e++)
        for (int k = 0; k < nsize; k++) {
            const int d = edof[(size_t)k * nElem + e];
            if (d >= row_lo && d < row_lo + nloc) rinc_ptr[d - row_lo + 1]++;
        }
    for (int r = 0; r < nloc; r++) rinc_ptr[r + 1] += rinc_ptr[r];
    std::vector<int> rinc(rinc_ptr[nloc] > 0 ? rinc_ptr[nloc] : 1), cur(rinc_ptr.begin(), rinc_ptr.end() - 1);
    for (int e = 0; e < nElem; e++)
        for (int k = 0; k < nsize; k++) {
            const int d = edof[(size_t)k * nElem + e];
            if (d >= row_lo && d < row_lo + nloc) rinc[cur[d - row_lo]++] = e * nsize + k;
        }
    const int nslices = (nloc + 31) / 32;
    std::vector<long long> ainc_off(nslices + 1, 0);
    for (int s = 0; s < nslices; s++) {
        int w = 0;
        for (int l = 0; l < 32 && s * 32 + l < nloc; l++) w = std::max(w, rinc_ptr[s * 32 + l + 1] - rinc_ptr[s * 32 + l]);
        ainc_off[s + 1] = ainc_off[s] + (long long)w * 32;
    }
    std::vector<int> ainc((size_t)ainc_off[nslices] * 2 + 2, -1);
    for (int r = 0; r < nloc; r++) {
        const int c0 = rowptr[r], len = rowptr[r + 1] - c0;
        int m = 0;
        for (int q = rinc_ptr[r]; q < rinc_ptr[r + 1]; q++, m++) {
            const int code = rinc[q], e = code / nsize;
            unsigned int w = 0;
            for (int j = 0; j < nsize; j++) {
                const int c = edof[(size_t)j * nElem + e];
                unsigned int sl = 255u;
                if (c >= 0) sl = (unsigned int)(std::lower_bound(col + c0, col + c0 + len, c) - (col + c0));
                w |= (sl & 255u) << (8 * j);
            }
            int *out = ainc.data() + (size_t)(ainc_off[r >> 5] + (r & 31) + (long long)m * 32) * 2;
            out[0] = code;
            out[1] = (int)w;
        }
    }
    // --- the product's tile builder ---
    TileInput in;
    in.nloc = nloc; in.row_lo = row_lo; in.nElem = nElem; in.npe = npe; in.nsize = nsize; in.rec_ints = rec_ints;
    in.ndim = ndim; in.xyz_stride = stride; in.erec = erec.data(); in.xyz = xyz.data(); in.rowptr = rowptr;
    in.rinc_ptr = rinc_ptr.data(); in.rinc = rinc.data(); in.ainc_off = ainc_off.data(); in.ainc = ainc.data();
    in.ainc_words = 2; in.max_rows = tile_rows; in.smem_budget = (size_t)smem_budget; in.cta_threads = threads;
    in.mode = mode == 2 ? 2 : 1;
    TileSet ts;
    if (build_tiles(in, ts) != 0) return 1;
    if (stats) { stats[0] = ts.ntiles; stats[1] = ts.elem_visits; stats[2] = ts.elems_touched; stats[3] = (long long)ts.max_smem; stats[4] = 0; }
    if (build_only) return 0;
    // --- the product's kernel ---
    double ed[8] = {0}, td[8] = {0};
    const int ned = kind == POISSON_TRIA ? 2 : 3;
    for (int i = 0; i < ned; i++) ed[i] = elemData[i];
    td[1] = timeData[1];
    int neg_flag = 0;
    TiledArgs a;
    a.tdesc = ts.tdesc.data();
    std::vector<int4> trows4(ts.trows.size() / 4 + 1);
    std::memcpy(trows4.data(), ts.trows.data(), ts.trows.size() * sizeof(int));
    a.trows = trows4.data();
    a.tel = reinterpret_cast<const int2 *>(ts.tel.data());
    a.tslice_off = ts.tslice_off.data();
    a.tinc = reinterpret_cast<const int2 *>(ts.tinc.data());
    std::vector<uint4> cnt4(ts.cnt.size() / 16 + 1);
    if (!ts.cnt.empty()) std::memcpy(cnt4.data(), ts.cnt.data(), ts.cnt.size());
    a.crec = reinterpret_cast<const int2 *>(ts.crec.data());
    a.cnt = cnt4.data();
    std::vector<int4> ts24(ts.ts2.size() / 4 + 1);
    if (!ts.ts2.empty()) std::memcpy(ts24.data(), ts.ts2.data(), ts.ts2.size() * sizeof(int));
    a.ts2 = ts24.data();
    std::vector<int4> c4(nElem);
    std::memcpy(c4.data(), conn4.data(), (size_t)nElem * sizeof(int4));
    a.conn4 = c4.data();
    a.erec = erec.data(); a.rec_ints = rec_ints; a.xyz = xyz.data(); a.applied = applied; a.rowptr = rowptr;
    a.val = val; a.rhs = rhs; a.elemData = ed; a.timeData = td; a.neg_flag = &neg_flag;
    a.load_val = load; a.load_rhs = load;
    const bool unit = td[1] == 1.0 && ed[0] == 1.0 && ed[1] == 1.0 && (kind == POISSON_TRIA || ed[2] == 1.0);
    const size_t smem = ts.max_smem;
#define RUN(K, T)                                                      \
    do {                                                               \
        if (unit) run_grid<K, T, true>(a, ts.ntiles, smem, mode);      \
        else run_grid<K, T, false>(a, ts.ntiles, smem, mode);          \
    } while (0)
    if (threads == 128) { if (kind == POISSON_TRIA) RUN(POISSON_TRIA, 128); else RUN(POISSON_TETRA, 128); }
    else if (threads == 256) { if (kind == POISSON_TRIA) RUN(POISSON_TRIA, 256); else RUN(POISSON_TETRA, 256); }
    else if (threads == 512) { if (kind == POISSON_TRIA) RUN(POISSON_TRIA, 512); else RUN(POISSON_TETRA, 512); }
    else return 3;
    if (stats) {
        stats[0] = ts.ntiles; stats[1] = ts.elem_visits; stats[2] = ts.elems_touched; stats[3] = (long long)ts.max_smem;
        stats[4] = neg_flag;
    }
    return 0;
}
