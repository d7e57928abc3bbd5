// emu_tiled.cpp -- TEST INFRASTRUCTURE ONLY: runs the tiled value-pass kernel (pfemfort_b200/csrc/assembly_tiled.cuh)
// and its host tile builder (tiles.hpp) on the CPU, CTA by CTA, through tests/emu/cuda_shim.h.
//
// The harness first restates, on the host, the device-side array formats the pattern pass produces (element records,
// row incidence lists, the row-gather incidence streams with their slot bytes: pattern.cu), hands them to the very
// same build_tiles() the product calls, and then executes the very same kernel source.  tests/test_tiled_emu.py
// compares the resulting CSR values / RHS bit for bit with the oracle.  Built by that test with g++.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "cuda_shim.h"
#include "emu_formats.hpp"

thread_local uint3 threadIdx, blockIdx;
uint3 blockDim, gridDim;
EmuBarrier emu_barrier;
unsigned char *emu_smem = nullptr;

#include "assembly_tiled.cuh"

using namespace pfem;

template <int KIND, int THREADS, bool UNIT>
static void run_grid(const TiledArgs &args, int ntiles, size_t smem_bytes, int mode)
{
    // exactly smem_bytes, 16-byte aligned: an overrun of the CTA's shared memory is a heap overflow AddressSanitizer sees
    void *raw = nullptr;
    if (posix_memalign(&raw, 16, smem_bytes ? smem_bytes : 16) != 0) return;
    unsigned char *base = static_cast<unsigned char *>(raw);
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
    free(raw);
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
    EmuFormats F;
    emu_build_formats(kind, nElem, nNode, conn0, edof, xyz_soa, row_lo, nloc, rowptr, col, F);
    const int npe = F.npe, nsize = F.nsize, ndim = F.ndim, rec_ints = F.rec_ints, stride = F.stride;
    std::vector<int> &erec = F.erec, &conn4 = F.conn4, &rinc_ptr = F.rinc_ptr, &rinc = F.rinc, &ainc = F.ainc;
    std::vector<long long> &ainc_off = F.ainc_off;
    std::vector<double> &xyz = F.xyz;
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
