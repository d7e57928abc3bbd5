// emu_rows.cpp -- TEST INFRASTRUCTURE ONLY: runs the DEFAULT value-pass kernels (pfemfort_b200/csrc/assembly_rows.cuh:
// assemble_sell_kernel, the streamed row gather, and assemble_kernel, its binary-search fallback) on the CPU, CTA by CTA,
// through tests/emu/cuda_shim.h, for all four element kinds.  tests/test_rows_emu.py compares the CSR values / RHS bit
// for bit with the oracle.  The product never loads this.
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

#include "assembly_rows.cuh"

using namespace pfem;

template <class Kernel>
static void run_grid(Kernel kernel, int nblocks, int threads, size_t smem_bytes)
{
    // exactly smem_bytes, 16-byte aligned: an overrun of the CTA's shared memory is a heap overflow AddressSanitizer sees
    void *raw = nullptr;
    if (posix_memalign(&raw, 16, smem_bytes ? smem_bytes : 16) != 0) return;
    unsigned char *base = static_cast<unsigned char *>(raw);
    emu_smem = base;
    std::memset(base, 0xFF, smem_bytes);          // NaN poison
    emu_barrier.reset(threads);
    blockDim = uint3{(unsigned)threads, 1, 1};
    gridDim = uint3{(unsigned)nblocks, 1, 1};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t] {
            threadIdx = uint3{(unsigned)t, 0, 0};
            for (int b = 0; b < nblocks; b++) {
                blockIdx = uint3{(unsigned)b, 0, 0};
                kernel();
                emu_barrier.wait();
                if (t == 0) std::memset(base, 0xFF, smem_bytes);
                emu_barrier.wait();
            }
        });
    for (auto &th : pool) th.join();
    free(raw);
}

template <int KIND, int R>
static void run_kind(const AsmArgs &a, int nblocks, size_t smem, bool streamed, bool unit_ok)
{
    constexpr bool POISSON = KIND == POISSON_TRIA || KIND == POISSON_TETRA;
    if (!streamed) run_grid([&] { assemble_kernel<KIND, R>(a); }, nblocks, R, smem);
    else if (POISSON && a.unit && unit_ok) run_grid([&] { assemble_sell_kernel<KIND, R, POISSON>(a); }, nblocks, R, smem);
    else run_grid([&] { assemble_sell_kernel<KIND, R, false>(a); }, nblocks, R, smem);
}

// streamed != 0: assemble_sell_kernel, else assemble_kernel; R in {32, 128}.  flags[0] receives the kernel's negative-
// Jacobian indicator (count for the generic kernel, flag for the streamed one).
extern "C" int emu_assemble_rows(int kind, int nElem, int nNode, const int *conn0, const int *edof, const double *xyz_soa,
                                 const double *applied, int row_lo, int nloc, const int *rowptr, const int *col,
                                 const double *elemData, const double *timeData, int R, int streamed, int load, double *val,
                                 double *rhs, int *flags)
{
    if (kind < 0 || kind > 3 || (R != 32 && R != 128)) return 2;
    EmuFormats F;
    emu_build_formats(kind, nElem, nNode, conn0, edof, xyz_soa, row_lo, nloc, rowptr, col, F);
    const int nblocks = (nloc + R - 1) / R;
    int mn = 0, mi = 0;
    for (int b = 0; b < nblocks; b++) {
        const int r0 = b * R, r1 = std::min(r0 + R, nloc);
        mn = std::max(mn, rowptr[r1] - rowptr[r0]);
        mi = std::max(mi, F.rinc_ptr[r1] - F.rinc_ptr[r0]);
    }
    mn = (mn + 1) & ~1;                                           // plan_assembly (assembly.cu)
    const size_t smem = streamed ? (size_t)mn * 8 + 16 + (size_t)R * 8 : (size_t)mn * 12 + (size_t)mi * 4 + 16;
    double ed[8] = {0}, td[8] = {0};
    const int ned = kind == 0 ? 2 : kind == 1 ? 3 : kind == 2 ? 5 : 6;
    for (int i = 0; i < ned; i++) ed[i] = elemData[i];
    td[1] = timeData[1];
    int neg = 0;
    AsmArgs a;
    a.nloc = nloc; a.row_lo = row_lo; a.row_hi = row_lo + nloc; a.rec_ints = F.rec_ints;
    a.erec = F.erec.data(); a.xyz = F.xyz.data(); a.applied = applied; a.rowptr = rowptr; a.col = col; a.val = val; a.rhs = rhs;
    a.rinc_ptr = F.rinc_ptr.data(); a.rinc = F.rinc.data(); a.elemData = ed; a.timeData = td; a.neg_count = &neg;
    a.load_val = load; a.load_rhs = load; a.max_seg_nnz = mn; a.ainc_off = F.ainc_off.data(); a.ainc = F.ainc.data();
    std::vector<int4> c4(nElem + 1);
    std::memcpy(c4.data(), F.conn4.data(), (size_t)nElem * sizeof(int4));
    a.conn4 = reinterpret_cast<const int *>(c4.data());
    a.neg_flag = &neg;
    a.unit = (td[1] == 1.0 && ed[0] == 1.0 && ed[1] == 1.0 && (kind == 0 || ed[2] == 1.0)) ? 1 : 0;   // assemble_values
#define RUNK(K)                                                              \
    do {                                                                     \
        if (R == 32) run_kind<K, 32>(a, nblocks, smem, streamed != 0, true); \
        else run_kind<K, 128>(a, nblocks, smem, streamed != 0, true);        \
    } while (0)
    switch (kind) {
    case 0: RUNK(POISSON_TRIA); break;
    case 1: RUNK(POISSON_TETRA); break;
    case 2: RUNK(ELASTICITY_TRIA); break;
    default: RUNK(ELASTICITY_TETRA); break;
    }
    if (flags) flags[0] = neg;
    return 0;
}
