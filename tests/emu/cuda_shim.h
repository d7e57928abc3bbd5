// cuda_shim.h -- TEST INFRASTRUCTURE ONLY: the handful of CUDA names the tiled value-pass kernel uses, for the host.
//
// tests/emu compiles pfemfort_b200/csrc/assembly_tiled.cuh with g++ (-DPFEM_EMULATE -ffp-contract=off) and runs
// every CTA as a group of OS threads joined by a barrier, so that the kernel's index logic, staging, summation order
// and arithmetic can be checked against the oracle on a machine without a GPU.  Nothing in the product loads this.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <mutex>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned int x, y, z, w; };
static inline uint4 make_uint4(unsigned int x, unsigned int y, unsigned int z, unsigned int w) { return uint4{x, y, z, w}; }
struct uint3 { unsigned int x, y, z; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

static inline int min(int a, int b) { return a < b ? a : b; }

extern thread_local uint3 threadIdx, blockIdx;
extern uint3 blockDim, gridDim;

template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
static inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

// CTA-wide barrier
struct EmuBarrier {
    std::mutex m;
    std::condition_variable cv;
    int count = 0, waiting = 0;
    unsigned long long gen = 0;
    void reset(int n) { count = n; waiting = 0; }
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        const unsigned long long g = gen;
        if (++waiting == count) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};
extern EmuBarrier emu_barrier;
extern unsigned char *emu_smem;
#ifdef PFEM_EMU_BREAK_SYNC          // negative control for the ThreadSanitizer run: barriers removed, races must be reported
static inline void __syncthreads() {}
#else
static inline void __syncthreads() { emu_barrier.wait(); }
#endif
#define PFEM_DYN_SMEM(name) unsigned char *name = emu_smem

// Sanitizer run (not part of the default suite; needs LD_PRELOAD of libasan):
//   g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer ... (same flags as tests/test_*_emu.py) -o tests/emu/_build/libemu_X.so
//   ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) python -m pytest tests/test_tiled_emu.py tests/test_rows_emu.py
// The CTA's shared memory is an exact-size heap block, so shared-memory overruns are caught as well as global ones.
// Race detection: the same with -fsanitize=thread and libtsan.so (CUDA threads are OS threads, __syncthreads is a real
// barrier, so a missing barrier is a data race ThreadSanitizer reports; -DPFEM_EMU_BREAK_SYNC is the negative control).
