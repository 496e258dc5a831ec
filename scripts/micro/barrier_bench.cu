// Microbenchmark: grid barrier cost with 1 CTA per SM (developer tool).
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release(unsigned *p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_relaxed(unsigned *p, unsigned v) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

template <int MODE>
__global__ void k(unsigned *bar, int iters, long long *out, double *junk) {
    extern __shared__ double sm[];
    unsigned target = 0;
    long long t0 = clock64();
    cg::grid_group g = cg::this_grid();
    for (int it = 0; it < iters; it++) {
        if (MODE == 4) { junk[blockIdx.x * blockDim.x + threadIdx.x] = it; }  // some global writes before the barrier
        if (MODE == 0 || MODE == 4) {
            __syncthreads();
            target += gridDim.x;
            if (threadIdx.x == 0) { red_release(bar, 1u); while (ld_acquire(bar) < target) {} }
            __syncthreads();
        } else if (MODE == 1) {
            __syncthreads();
            target += gridDim.x;
            if (threadIdx.x == 0) { __threadfence(); red_relaxed(bar, 1u); while (ld_relaxed(bar) < target) {} __threadfence(); }
            __syncthreads();
        } else if (MODE == 2) {
            g.sync();
        } else if (MODE == 3) {  // per-CTA flags: each CTA writes its own slot, everybody reads all slots
            __syncthreads();
            target += 1;
            if (threadIdx.x == 0) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + blockIdx.x * 32), "r"(target) : "memory"); }
            if (threadIdx.x < gridDim.x) { while (ld_acquire(bar + threadIdx.x * 32) < target) {} }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (clock64() - t0) / iters;
}

int main() {
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    int nsm = pr.multiProcessorCount;
    unsigned *bar; long long *out; double *junk;
    cudaMalloc(&bar, 4096 * 4 * 32); cudaMalloc(&out, 8); cudaMalloc(&junk, 148 * 1024 * 8);
    int iters = 2000;
    const char *names[5] = {"red.release+ld.acquire", "fence+relaxed", "cg grid.sync", "per-CTA flags", "release/acquire + stores"};
    for (int threads : {512, 1024}) {
        for (int mode = 0; mode < 5; mode++) {
            cudaMemset(bar, 0, 4096 * 4 * 32);
            void *args[] = {&bar, &iters, &out, &junk};
            const void *f = mode == 0 ? (const void *)k<0> : mode == 1 ? (const void *)k<1> : mode == 2 ? (const void *)k<2> : mode == 3 ? (const void *)k<3> : (const void *)k<4>;
            cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            cudaError_t e = cudaLaunchCooperativeKernel(f, dim3(nsm), dim3(threads), args, 200 * 1024, 0);
            cudaEventRecord(e1);
            cudaError_t e2 = cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long cyc; cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
            printf("threads %4d  %-28s: %6lld cycles/barrier  %.3f us/barrier (%s %s)\n", threads, names[mode], cyc, ms * 1e3 / iters, cudaGetErrorString(e), cudaGetErrorString(e2));
        }
    }
    return 0;
}
