// Fixed effects ("--covariates", reference src/BayesRRm.cpp:2648-2681): once per Gibbs iteration, between the group
// hyper-parameters and sigmaE, every covariate f (in the iteration's shuffled order) gets
//     num   = sum_k X(k,f) * (eps_k + gamma_old * X(k,f))            (:2666-2668)
//     gamma = N(num / denom, sigmaE / denom),  denom = (N-1) + sigmaE / sigmaF, sigmaF = s02F = 1     (:2670-2671, BayesRRm.h:34)
//     eps  += (gamma_old - gamma) * X(:,f)                            (:2673-2676)
// One cooperative launch for all covariates: every CTA owns a fixed range of individuals (its part of eps never leaves the
// CTA's hands, so the update of covariate f and the partial sum of covariate f+1 need no synchronisation between them);
// one grid barrier per covariate for the N-sum, whose per-CTA partials are added in a fixed order by every CTA (all CTAs --
// and all GPUs -- compute the same gamma bit for bit). The slice statistics that the marker kernel leaves for the host
// (sum, sum of squares, sum |.|, max |.| per slice) are refreshed at the end.
#pragma once
#include "common.cuh"

namespace hb {

struct CovParams {
    uint32_t N, S, L, F;
    double *E;              // [S*L] stored residual (eps of task 0 = E + shift), slice layout of the marker kernel
    double shift;
    const double *X;        // [F][S*L] covariate columns on the same layout (individuals >= N: 0)
    const int32_t *xI;      // [F] order of this iteration
    const double *z;        // [F] standard normals of this iteration
    double *gamma;          // [F]
    double sigmaE, denom;
    double *part;           // [2][gridDim] partial sums (double buffered over the covariates)
    uint32_t *bar;          // grid barrier counter, zero at launch
    double *slice_sum, *slice_sq, *slice_abs, *slice_max;   // [S]
};

__device__ __forceinline__ double cov_block_sum(double v, double *red) {
    v = warp_sum(v);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (uint32_t w = 0; w < (blockDim.x >> 5); w++) s += red[w];
    return s;
}

__global__ void __launch_bounds__(256) k_cov_gibbs(const CovParams P) {
    __shared__ double red[32];
    const uint32_t nb = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
    const size_t n_all = (size_t)P.S * P.L;
    const size_t per = ((n_all + nb - 1) / nb + 255) & ~(size_t)255;
    const size_t i0 = ((size_t)b * per < n_all) ? (size_t)b * per : n_all, i1 = (i0 + per < n_all) ? i0 + per : n_all;
    uint32_t bar_target = 0;
    for (uint32_t i = 0; i < P.F; i++) {
        const int32_t f = P.xI[i];
        const double *x = P.X + (size_t)f * n_all;
        const double g_old = P.gamma[f];
        double v = 0.0;
        for (size_t k = i0 + tid; k < i1; k += blockDim.x) {
            const double xk = x[k];
            v += xk * ((P.E[k] + P.shift) + g_old * xk);      // padded individuals have x = 0
        }
        const double s = cov_block_sum(v, red);
        double *part = P.part + (size_t)(i & 1u) * nb;
        if (tid == 0) part[b] = s;
        grid_barrier(P.bar, bar_target, nb);
        // every CTA adds the partials in the same order
        double num = 0.0;
        if (tid < 32) {
            double a = 0.0;
            for (uint32_t j = tid; j < nb; j += 32) a += __ldcg(part + j);
            num = warp_sum(a);
        }
        if (tid == 0) red[0] = num;
        __syncthreads();
        num = red[0];
        __syncthreads();
        const double g_new = num / P.denom + sqrt(P.sigmaE / P.denom) * P.z[i];
        const double d = g_old - g_new;
        for (size_t k = i0 + tid; k < i1; k += blockDim.x) P.E[k] += d * x[k];
        if (b == 0 && tid == 0) P.gamma[f] = g_new;
        // (gamma[f] is read again only in a later iteration / launch)
    }
    grid_barrier(P.bar, bar_target, nb);
    // slice statistics, one CTA per slice
    for (uint32_t c = b; c < P.S; c += nb) {
        double v = 0.0, v2 = 0.0, va = 0.0, vm = 0.0;
        for (uint32_t k = tid; k < P.L; k += blockDim.x) {
            const uint32_t gi = c * P.L + k;
            if (gi < P.N) {
                const double e = __ldcg(P.E + gi);
                v += e; v2 += e * e; va += fabs(e); vm = fmax(vm, fabs(e));
            }
        }
        const double s1 = cov_block_sum(v, red), s2 = cov_block_sum(v2, red), sa = cov_block_sum(va, red);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vm = fmax(vm, __shfl_xor_sync(0xffffffffu, vm, o));
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = vm;
        __syncthreads();
        if (tid == 0) {
            double mx = 0.0;
            for (uint32_t w = 0; w < (blockDim.x >> 5); w++) mx = fmax(mx, red[w]);
            P.slice_sum[c] = s1; P.slice_sq[c] = s2; P.slice_abs[c] = sa; P.slice_max[c] = mx;
        }
        __syncthreads();
    }
}

}  // namespace hb
