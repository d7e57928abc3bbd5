// sync_micro.cu -- microbenchmark of grid-barrier building blocks on one B200 (cooperative launch, one CTA per SM).
// Prints ns per operation for: CTA barrier, gpu-scope fences, L2 atomics, and five complete grid-barrier designs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/sync_micro tools/sync_micro.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

struct Slot { double v; unsigned long long tag; };

__device__ __forceinline__ void st_slot(Slot *p, double v, unsigned long long t)
{
    asm volatile("st.global.relaxed.gpu.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(t) : "memory");
}
__device__ __forceinline__ void ld_slot(const Slot *p, double &v, unsigned long long &t)
{
    long long b;
    asm volatile("ld.global.relaxed.gpu.v2.b64 {%0, %1}, [%2];" : "=l"(b), "=l"(t) : "l"(p) : "memory");
    v = __longlong_as_double(b);
}
__device__ __forceinline__ unsigned long long ld_acq(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.global.acquire.gpu.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_rlx(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.global.relaxed.gpu.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long atom_rel(unsigned long long *p, unsigned long long v)
{
    unsigned long long o;
    asm volatile("atom.add.release.gpu.u64 %0, [%1], %2;" : "=l"(o) : "l"(p), "l"(v) : "memory");
    return o;
}
__device__ __forceinline__ void red_rel(unsigned long long *p, unsigned long long v)
{
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_rlx(unsigned long long *p, unsigned long long v)
{
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_acqrel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_sc() { asm volatile("fence.sc.gpu;" ::: "memory"); }

// mode: which experiment; reps iterations; out[mode] = elapsed clock64 of CTA 0
__global__ void __launch_bounds__(1024, 1)
micro(int mode, int reps, unsigned long long *ctr, Slot *slots, double *data, long long *out)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[64];
    const int tid = threadIdx.x, G = gridDim.x;
    unsigned long long epoch = 0;
    double acc = 0.0;
    grid.sync();
    const long long t0 = clock64();
    for (int it = 0; it < reps; it++) {
        switch (mode) {
        case 0: __syncthreads(); break;                                  // CTA barrier alone
        case 1: if (tid == 0) __threadfence(); __syncthreads(); break;   // + one thread fences (nothing pending)
        case 2: __threadfence(); __syncthreads(); break;                 // + every thread fences
        case 3: data[blockIdx.x * 1024 + tid] = acc; __syncthreads(); if (tid == 0) __threadfence(); __syncthreads(); break;  // pending stores
        case 4: if (tid == 0) acc += (double)atomicAdd(ctr + 8 + blockIdx.x * 16, 1ULL); __syncthreads(); break;   // private-address atomic RT
        case 5: if (tid == 0) acc += (double)ld_rlx(ctr + 8 + blockIdx.x * 16); __syncthreads(); break;            // L2 load RT
        case 6: if (tid == 0) fence_acqrel(); __syncthreads(); break;
        case 7: if (tid == 0) fence_sc(); __syncthreads(); break;
        case 10: grid.sync(); break;                                     // cooperative groups
        case 11: {                                                       // counter barrier: fence + relaxed red + acquire poll
            epoch++;
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                red_rlx(ctr, 1ULL);
                while (ld_rlx(ctr) < epoch * G) {}
                __threadfence();
            }
            __syncthreads();
            break;
        }
        case 12: {                                                       // counter barrier: release red + acquire load poll
            epoch++;
            __syncthreads();
            if (tid == 0) {
                red_rel(ctr, 1ULL);
                while (ld_acq(ctr) < epoch * G) {}
            }
            __syncthreads();
            break;
        }
        case 13: {                                                       // the r01 design: atom.release, last CTA publishes a slot
            epoch++;
            __syncthreads();
            __shared__ int s_last;
            if (tid == 0) {
                const unsigned long long t = atom_rel(ctr, 1ULL);
                s_last = (t == (unsigned long long)G * epoch - 1ULL);
            }
            __syncthreads();
            if (s_last) {
                __threadfence();
                if (tid == 0) { __threadfence(); st_slot(slots + 4096, 1.0, epoch); }
            } else if (tid == 0) {
                double v; unsigned long long tg;
                ld_slot(slots + 4096, v, tg);
                while (tg != epoch) { __nanosleep(32); ld_slot(slots + 4096, v, tg); }
                __threadfence();
            }
            __syncthreads();
            break;
        }
        case 14: {                                                       // all-to-all slots, every poller fences
            epoch++;
            Slot *sl = slots + (epoch & 1) * 2048;
            __syncthreads();
            if (tid == 0) { __threadfence(); st_slot(sl + blockIdx.x, 1.0, epoch); }
            if (tid < G) {
                double v; unsigned long long tg;
                ld_slot(sl + tid, v, tg);
                while (tg != epoch) ld_slot(sl + tid, v, tg);
                acc += v;
                __threadfence();
            }
            __syncthreads();
            break;
        }
        case 15: {                                                       // all-to-all slots, ONE fence after the CTA barrier
            epoch++;
            Slot *sl = slots + (epoch & 1) * 2048;
            __syncthreads();
            if (tid == 0) { __threadfence(); st_slot(sl + blockIdx.x, 1.0, epoch); }
            if (tid < G) {
                double v; unsigned long long tg;
                ld_slot(sl + tid, v, tg);
                while (tg != epoch) ld_slot(sl + tid, v, tg);
                acc += v;
            }
            __syncthreads();
            if (tid == 0) __threadfence();
            __syncthreads();
            break;
        }
        case 16: {                                                       // all-to-all slots, no fences at all (lower bound)
            epoch++;
            Slot *sl = slots + (epoch & 1) * 2048;
            __syncthreads();
            if (tid == 0) st_slot(sl + blockIdx.x, 1.0, epoch);
            if (tid < G) {
                double v; unsigned long long tg;
                ld_slot(sl + tid, v, tg);
                while (tg != epoch) ld_slot(sl + tid, v, tg);
                acc += v;
            }
            __syncthreads();
            break;
        }
        case 17: {                                                       // counter, no fences (lower bound)
            epoch++;
            __syncthreads();
            if (tid == 0) {
                red_rlx(ctr, 1ULL);
                while (ld_rlx(ctr) < epoch * G) {}
            }
            __syncthreads();
            break;
        }
        case 18: {                                                       // counter + fence, then a dependent 3-load chain from cold L1
            epoch++;
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                red_rlx(ctr, 1ULL);
                while (ld_rlx(ctr) < epoch * G) {}
                __threadfence();
            }
            __syncthreads();
            int i = (blockIdx.x * 1024 + tid) & 4095;
            i = (int)data[i] & 4095; i = (int)data[i + 1] & 4095; acc += data[i + 2];
            break;
        }
        }
    }
    const long long t1 = clock64();
    sh[0] = acc;
    if (blockIdx.x == 0 && tid == 0) out[mode] = t1 - t0;
    if (acc == 12345.678) data[0] = sh[0];
}

int main()
{
    int dev = 0, sms = 0, khz = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    unsigned long long *ctr; Slot *slots; double *data; long long *out;
    cudaMalloc(&ctr, 8 * (8 + 16 * 256)); cudaMalloc(&slots, sizeof(Slot) * 8192); cudaMalloc(&data, 8 * 1024 * 256);
    cudaMallocManaged(&out, 8 * 64);
    const int modes[] = {0, 1, 2, 3, 4, 5, 6, 7, 10, 11, 12, 13, 14, 15, 16, 17, 18};
    const char *names[] = {"syncthreads", "+fence by thread 0 (idle)", "+fence by all 1024 threads", "1024 stores, sync, fence t0, sync",
                           "atomicAdd RT (private addr)", "ld.relaxed.gpu RT", "fence.acq_rel.gpu t0", "fence.sc.gpu t0",
                           "cg grid.sync", "counter: fence+red+poll+fence", "counter: red.release + ld.acquire poll",
                           "r01 design: atom.release, last CTA slot", "a2a slots, every poller fences", "a2a slots, one acquire fence",
                           "a2a slots, no fence (bound)", "counter, no fence (bound)", "counter+fences + 3 dependent cold loads"};
    for (int threads : {1024, 256}) {
        for (size_t m = 0; m < sizeof(modes) / sizeof(int); m++) {
            int mode = modes[m], reps = 2000;
            cudaMemset(ctr, 0, 8 * (8 + 16 * 256)); cudaMemset(slots, 0, sizeof(Slot) * 8192); cudaMemset(data, 0, 8 * 1024 * 256);
            void *args[] = {&mode, &reps, &ctr, &slots, &data, &out};
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            cudaError_t e = cudaLaunchCooperativeKernel((void *)micro, dim3(sms), dim3(threads), args, 0, 0);
            cudaEventRecord(e1);
            cudaError_t e2 = cudaDeviceSynchronize();
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("threads %4d mode %2d %-44s : %8.1f ns/op (clock64 %lld cyc/op)  %s %s\n", threads, mode, names[m], 1e6 * ms / reps,
                   out[mode] / reps, e ? cudaGetErrorString(e) : "", e2 ? cudaGetErrorString(e2) : "");
        }
    }
    return 0;
}
