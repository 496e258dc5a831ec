// fh_kernels.cuh -- bayesFHMPI (horseshoe-type local scales) around the marker kernel.
//
// The reference draws two inverse-gamma variates per marker inside its marker loop (src/BayesRRm.cpp:1729 nu_var before the
// mixture draw, :1952 lambda_var after it) and builds the marker's own prior variance lambda_tilde from them (:1730). Neither
// draw depends on another marker of the same iteration: nu_var(m) and lambda_tilde(m) use lambda_var(m) of the PREVIOUS
// iteration, tau and c_slab; lambda_var(m) uses the marker's final effect. So both leave the latency chain of the window loop:
//   k_fh_prepare  (before the marker loop)  nu_var, lambda_tilde and the three numbers the mixture draw needs per marker
//                                           { denom (:1748), 0.5*log(...) (:1871), sqrt(sigmaE/denom) (:1901) }
//   k_fh_finish   (after the marker loop)   lambda_var (:1952) and the chunk sums of beta^2 / lambda_var (:2506-2509)
// k_brr_iteration reads the per-marker triple instead of the per-(group, component) tables.
#pragma once
#include "brr_kernel.cuh"
#include "common.cuh"

namespace hb {

constexpr uint32_t kTagFhNu = 0x46484E55u;  // 'FHNU'
constexpr uint32_t kTagFhLa = 0x46484C41u;  // 'FHLA'

// Standard Gamma(a, 1) variate of RNG spec v1 (oracle: ho_fh_gamma): Marsaglia-Tsang, one Philox block per attempt, keyed by
// (seed, global marker), counter (attempt, iteration, tag)
__device__ __forceinline__ double fh_gamma(uint32_t seed, uint32_t marker, uint32_t iteration, uint32_t tag, double a) {
    const double a1 = (a < 1.0) ? a + 1.0 : a;
    const double d = a1 - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double g = d;
    for (uint32_t attempt = 0; attempt < 4096u; attempt++) {
        uint32_t w[4];
        philox4x32(attempt, iteration, tag, 0u, seed, marker, w);
        const double u1 = ((double)w[0] + 0.5) * (1.0 / 4294967296.0);
        const double u2 = ((double)w[1] + 0.5) * (1.0 / 4294967296.0);
        const double x = sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        const double u = 1.0 - ((double)(w[2] >> 5) * 67108864.0 + (double)(w[3] >> 6)) * (1.0 / 9007199254740992.0);
        g = d * v;
        if (u < 1.0 - 0.0331 * (x * x) * (x * x)) break;
        if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) break;
    }
    if (a < 1.0) {
        uint32_t w[4];
        philox4x32(0xFFFFFFFFu, iteration, tag, 0u, seed, marker, w);
        g *= pow(((double)w[0] + 0.5) * (1.0 / 4294967296.0), 1.0 / a);
    }
    return g;
}

struct FhParams {
    uint32_t M, m_start, seed, iteration;
    double shape;            // 0.5 + 0.5*v0L
    double v0L, tau, sigmaE, dNm1;
    const double *c_slab;    // [G]
    const int32_t *grp;      // [M]
    const double *g_tape;    // [M] standard gamma variates (replay) or NULL = spec v1
    const double *beta;      // [M] (finish)
    double *lambda, *nu;     // [M]
    double *par;             // [3*M] denom, chalf, sd (prepare)
    double *part;            // [gridDim.x] chunk sums of beta^2/lambda (finish)
};

// src/distributions_boost.cpp:97-103: inv_gamma_rate_rng(shape, rate) = 1 / rgamma(shape, 1/rate)
__device__ __forceinline__ double inv_gamma_rate_from(double g, double rate) { return __drcp_rn(__dmul_rn(g, __drcp_rn(rate))); }

__global__ void __launch_bounds__(256) k_fh_prepare(const FhParams Q) {
    for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < Q.M; m += gridDim.x * blockDim.x) {
        const double g = Q.g_tape ? Q.g_tape[m] : fh_gamma(Q.seed, Q.m_start + m, Q.iteration, kTagFhNu, Q.shape);
        const double lam = Q.lambda[m];
        Q.nu[m] = inv_gamma_rate_from(g, __dadd_rn(Q.v0L / lam, 1.0));                       // :1729
        const double cs = Q.c_slab[Q.grp[m]];
        const double lt = __dmul_rn(Q.tau, cs) / __dadd_rn(Q.tau, __dmul_rn(cs, lam));       // :1730
        const double den = __dadd_rn(Q.dNm1, Q.sigmaE / lt);                                 // :1748
        Q.par[3 * (size_t)m + 0] = den;
        Q.par[3 * (size_t)m + 1] = 0.5 * log(__dadd_rn(__dmul_rn(lt / Q.sigmaE, Q.dNm1), 1.0));  // :1871
        Q.par[3 * (size_t)m + 2] = sqrt(Q.sigmaE / den);                                     // :1901
    }
}

// chunk b takes a contiguous range of markers; a second launch (k_beta_sqnorm_fin with G = 1) adds the chunk sums in order
__global__ void __launch_bounds__(256) k_fh_finish(const FhParams Q) {
    __shared__ double red[32];
    const uint32_t per = (Q.M + gridDim.x - 1) / gridDim.x, m0 = blockIdx.x * per, m1 = min(Q.M, m0 + per);
    double v = 0.0;
    for (uint32_t m = m0 + threadIdx.x; m < m1; m += blockDim.x) {
        const double g = Q.g_tape ? Q.g_tape[m] : fh_gamma(Q.seed, Q.m_start + m, Q.iteration, kTagFhLa, Q.shape);
        const double b = Q.beta[m];
        const double rate = __dadd_rn(__dmul_rn(__dmul_rn(0.5, b), b) / Q.tau, Q.v0L / Q.nu[m]);   // :1952
        const double lam = inv_gamma_rate_from(g, rate);
        Q.lambda[m] = lam;
        v += __dmul_rn(b, b) / lam;                                                           // :2508
    }
    const double s = block_sum(v, red);
    if (threadIdx.x == 0) Q.part[blockIdx.x] = s;
}

}  // namespace hb
