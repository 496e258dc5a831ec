// Microbenchmark: 64-bit gathers from a shared-memory slice, the inner operation of the dot phase (developer tool).
// One CTA of 512 threads per SM holds L doubles; every lane owns 4 words of 4 x u16 indices per step (= one
// 128-word sub-chunk per warp) and does the 16 gathers + the adds of dot_sub.  Index patterns:
//   0 conflict-free (lane l of a half-warp hits bank pair l)     1 uniformly random
//   2 random, but every aligned group of 16 lanes gets 16 distinct residues mod 16 (ideal bank order)
//   3 as 2 for 80 % of the gathers, random for the rest
// Variants: A = plain dot_sub; B = 2 accumulators; reports cycles per sub-chunk step per warp and LDS.64 / clk / SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

constexpr int kThreads = 512;
constexpr int L = 23872;
constexpr int kSteps = 64;  // sub-chunks per warp, indices preloaded per step from global (L2 resident)

__device__ __forceinline__ double gather4(unsigned long long x, const double *E) {
    return (E[x & 0xFFFFu] + E[(x >> 16) & 0xFFFFu]) + (E[(x >> 32) & 0xFFFFu] + E[x >> 48]);
}

template <int VAR>
__global__ void __launch_bounds__(kThreads, 1) k(const unsigned long long *idx, int reps, long long *cyc, double *out) {
    extern __shared__ double E[];
    for (int i = threadIdx.x; i <= L; i += kThreads) E[i] = 1.0 + i * 1e-6;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long *base = idx + (size_t)warp * kSteps * 128;
    double acc = 0.0, acc2 = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
        unsigned long long xc[4], xn[4];
#pragma unroll
        for (int t = 0; t < 4; t++) xc[t] = base[lane + 32 * t];
        for (int s = 0; s < kSteps; s++) {
            const unsigned long long *nx = base + (size_t)((s + 1) % kSteps) * 128;
#pragma unroll
            for (int t = 0; t < 4; t++) xn[t] = nx[lane + 32 * t];
            if (VAR == 0) {
#pragma unroll
                for (int t = 0; t < 4; t++) acc = fma(1.0 + t, gather4(xc[t], E), acc);
            } else {
                acc = fma(1.0, gather4(xc[0], E), acc);
                acc2 = fma(2.0, gather4(xc[1], E), acc2);
                acc = fma(1.0, gather4(xc[2], E), acc);
                acc2 = fma(2.0, gather4(xc[3], E), acc2);
            }
            if (VAR == 2) {  // with the warp reduction of a unit end
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            }
#pragma unroll
            for (int t = 0; t < 4; t++) xc[t] = xn[t];
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * kThreads + threadIdx.x] = acc + acc2;
}

static void make_indices(int pattern, std::vector<unsigned long long> &v) {
    // layout: [warp][step][word 0..127], lane l reads words l, l+32, l+64, l+96; gather q of the 16 lanes of a half-warp =
    // index q of words w0 + 16h + {0..15}
    v.assign((size_t)16 * kSteps * 128, 0);
    srand(12345);
    for (size_t blk = 0; blk < v.size() / 16; blk++) {  // 16 consecutive words
        for (int q = 0; q < 4; q++) {
            int perm[16];
            for (int i = 0; i < 16; i++) perm[i] = i;
            for (int i = 15; i > 0; i--) { int j = rand() % (i + 1); int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
            for (int l = 0; l < 16; l++) {
                unsigned id;
                const unsigned rnd = (unsigned)(rand() % (L / 16));
                if (pattern == 0) id = rnd * 16 + l;
                else if (pattern == 1) id = (unsigned)(rand() % L);
                else if (pattern == 2) id = rnd * 16 + perm[l];
                else id = (rand() % 100 < 80) ? rnd * 16 + perm[l] : (unsigned)(rand() % L);
                v[blk * 16 + l] |= (unsigned long long)id << (16 * q);
            }
        }
    }
}

int main() {
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = (size_t)(L + 2) * 8;
    unsigned long long *d_idx;
    long long *d_cyc;
    double *d_out;
    cudaMalloc(&d_idx, (size_t)16 * kSteps * 128 * 8);
    cudaMalloc(&d_cyc, nsm * 8);
    cudaMalloc(&d_out, (size_t)nsm * kThreads * 8);
    const int reps = 20;
    for (int var = 0; var < 3; var++) {
        for (int pat = 0; pat < 4; pat++) {
            std::vector<unsigned long long> h;
            make_indices(pat, h);
            cudaMemcpy(d_idx, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
            auto launch = [&](int rp) {
                if (var == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<0><<<nsm, kThreads, smem>>>(d_idx, rp, d_cyc, d_out); }
                if (var == 1) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<1><<<nsm, kThreads, smem>>>(d_idx, rp, d_cyc, d_out); }
                if (var == 2) { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<2><<<nsm, kThreads, smem>>>(d_idx, rp, d_cyc, d_out); }
            };
            launch(2);
            launch(reps);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<long long> c(nsm);
            cudaMemcpy(c.data(), d_cyc, nsm * 8, cudaMemcpyDeviceToHost);
            double mean = 0;
            for (auto x : c) mean += x;
            mean /= nsm;
            const double per_step = mean / (reps * kSteps);           // cycles per sub-chunk step (all 16 warps in parallel)
            const double lds_per_clk = 16.0 * 16.0 / per_step;        // warp-level LDS.64 per clock per SM
            printf("variant %d pattern %d: %.0f cycles per step of 16 warps x 16 LDS.64 -> %.3f LDS.64/clk/SM = %.2f gathers/clk/SM (%.2f clk per LDS.64)\n",
                   var, pat, per_step, lds_per_clk, lds_per_clk * 32, 1.0 / lds_per_clk);
        }
    }
    return 0;
}
