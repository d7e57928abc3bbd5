// emu_explicit.cpp -- TEST INFRASTRUCTURE ONLY: runs the explicit-dynamics gather kernels of
// pfemfort_b200/csrc/explicit.cuh (ex_mass_kernel, ex_step_kernel) on the CPU through tests/emu/cuda_shim.h, thread by thread
// (the kernels have no shared memory and no barrier, so CUDA threads run one after the other).  tests/test_explicit_emu.py
// compares lumped mass and time-loop state bit for bit with the oracle.  The product never loads this.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "cuda_shim.h"

#ifndef PFEM_EMU_HAVE_DOUBLE4
struct __attribute__((aligned(32))) double4 { double x, y, z, w; };
#endif

thread_local uint3 threadIdx, blockIdx;
uint3 blockDim, gridDim;
EmuBarrier emu_barrier;
unsigned char *emu_smem = nullptr;

#include "explicit.cuh"

using namespace pfem;

template <class Kernel> static void run_grid(Kernel kernel, int nblocks, int threads)
{
    blockDim = uint3{(unsigned)threads, 1, 1};
    gridDim = uint3{(unsigned)nblocks, 1, 1};
    for (int b = 0; b < nblocks; b++)
        for (int t = 0; t < threads; t++) {
            blockIdx = uint3{(unsigned)b, 0, 0};
            threadIdx = uint3{(unsigned)t, 0, 0};
            kernel();
        }
}

// kind: 2 = ELASTICITY_TRIA, 3 = ELASTICITY_TETRA.  xyz: AoS double2 (2-D) / double4 (3-D) like the device layout.
extern "C" void emu_explicit_mass(int kind, int nNode, const int *inc_ptr, const int *inc, const int *conn4, const double *xyz,
                                  const double *prm, double *M, int *neg)
{
    const int nblocks = (nNode + 127) / 128;
    if (kind == ELASTICITY_TRIA) run_grid([&] { ex_mass_kernel<ELASTICITY_TRIA>(nNode, inc_ptr, inc, conn4, xyz, prm, M, neg); }, nblocks, 128);
    else run_grid([&] { ex_mass_kernel<ELASTICITY_TETRA>(nNode, inc_ptr, inc, conn4, xyz, prm, M, neg); }, nblocks, 128);
}

extern "C" void emu_explicit_step(int kind, int nNode, const int *inc_ptr, const int *inc, const int *conn4, const double *xyz,
                                  const double *prm, const double *M, const unsigned char *free_mask, const double *d1, const double *d2,
                                  double *d0, double *velo, double *acce, double dt, int *neg)
{
    const int nblocks = (nNode + 127) / 128;
    if (kind == ELASTICITY_TRIA)
        run_grid([&] { ex_step_kernel<ELASTICITY_TRIA>(nNode, inc_ptr, inc, conn4, xyz, prm, M, free_mask, d1, d2, d0, velo, acce, dt, neg); }, nblocks, 128);
    else
        run_grid([&] { ex_step_kernel<ELASTICITY_TETRA>(nNode, inc_ptr, inc, conn4, xyz, prm, M, free_mask, d1, d2, d0, velo, acce, dt, neg); }, nblocks, 128);
}
