// Scalar mathematics of BayesW's per-marker update (reference src/BayesW.cpp:145-169, 174-726), for device and host.
#pragma once
#include <math.h>

#include "arms.cuh"

namespace hb {

#define HB_GH_CONST static const
#define HB_GH_NAME(x) gh_h_##x
#include "gh_tables.inc"
#undef HB_GH_CONST
#undef HB_GH_NAME
#ifdef __CUDACC__
#define HB_GH_CONST __device__ const
#define HB_GH_NAME(x) gh_d_##x
#include "gh_tables.inc"
#undef HB_GH_CONST
#undef HB_GH_NAME
#endif
#ifdef __CUDA_ARCH__
#define HB_GH(x) gh_d_##x
#else
#define HB_GH(x) gh_h_##x
#endif

constexpr double kBwEuMasc = 0.577215664901532;  // src/BayesW.cpp:42
constexpr double kBwSqrtPi = 1.77245385090552;   // :41

// index of the n-point rule (--quad_points; :706-709 exits for anything else), -1 if unsupported
HB_HD inline int bw_gh_rule(int n) {
    for (int i = 0; i < HB_GH_NRULES; i++)
        if (HB_GH(n)[i] == n) return i;
    return -1;
}

// what the per-marker densities need (struct pars_beta_sparse of the reference, src/BayesW.hpp)
struct BwMarker {
    double alpha, sigmaG, sum_failure;            // used_data_beta
    double vi_sum, vi_0, vi_1, vi_2;              // sums of vi = exp(alpha*eps - EuMasc) by genotype class
    double mean, sd, mean_sd_ratio;               // mave, mstd (the SD in BayesW), mave/mstd
};

// :161-169
HB_HD inline double bw_gh_integrand(double s, const BwMarker &m, double sqrt_2Ck_sigmaG) {
    const double temp = -m.alpha * s * m.sum_failure * sqrt_2Ck_sigmaG + m.vi_sum -
                        exp(m.alpha * m.mean_sd_ratio * s * sqrt_2Ck_sigmaG) *
                            (m.vi_0 + m.vi_1 * exp(-m.alpha * s * sqrt_2Ck_sigmaG / m.sd) + m.vi_2 * exp(-2 * m.alpha * s * sqrt_2Ck_sigmaG / m.sd)) -
                        s * s;
    return exp(temp);
}

// :174-712: sigma * (sum_i w_i * integrand(sigma * x_i) + w_centre), summed in the reference's order
HB_HD inline double bw_gh_integral(int rule, double C_k, double sigma, const BwMarker &m) {
    const double sq = sqrt(2 * C_k * m.sigmaG);
    const int n = HB_GH(n)[rule], o = HB_GH(off)[rule];
    double temp = 0.0;
    for (int i = 0; i < n - 1; i++) {
        const double t = HB_GH(w)[o + i] * bw_gh_integrand(sigma * HB_GH(x)[o + i], m, sq);
        temp = (i == 0) ? t : temp + t;
    }
    temp = temp + HB_GH(wc)[rule];
    return sigma * temp;
}

// :716-726: post[k] for k = 1..km1 (post[0] = pi_0 * sqrt(pi) is set by the caller, :1473, 1496)
HB_HD inline void bw_marginal_likelihoods(int rule, const double *prior, const double *cVa /*[km1]*/, int km1, const BwMarker &m,
                                          double *post) {
    const double exp_sum = (m.vi_1 * (1 - 2 * m.mean) + 4 * (1 - m.mean) * m.vi_2 + m.vi_sum * m.mean * m.mean) / (m.sd * m.sd);
    for (int i = 0; i < km1; i++) {
        const double sigma = 1.0 / sqrt(1 + m.alpha * m.alpha * m.sigmaG * cVa[i] * exp_sum);
        post[i + 1] = prior[i + 1] * bw_gh_integral(rule, cVa[i], sigma, m);
    }
}

// :145-156 log-density of beta for mixture variance C_k
struct BwBetaDens {
    BwMarker m;
    double mixture_value;
    HB_HD double operator()(double x) const {
        return -m.alpha * x * m.sum_failure -
               exp(m.alpha * x * m.mean_sd_ratio) * (m.vi_0 + m.vi_1 * exp(-m.alpha * x / m.sd) + m.vi_2 * exp(-2 * m.alpha * x / m.sd)) -
               x * x / (2 * mixture_value * m.sigmaG);
    }
};

// ARMS draw of beta around the previous value (:1562-1582): xinit = b - L/10, b, b + L/20, b + L/10; bounds b -+ L,
// L = 2*sqrt(sumSigmaG * C_k)
template <class URand, class Cumulate = ArmsSerialCumulate>
HB_HD inline int bw_sample_beta(const BwMarker &m, double C_k, double sum_sigmaG, double beta_old, URand &urand, double *beta_new,
                                ArmsEnvelope &env, Cumulate cumulate = Cumulate()) {
    const double safe_limit = 2 * sqrt(sum_sigmaG * C_k);
    const double xinit[4] = {beta_old - safe_limit / 10, beta_old, beta_old + safe_limit / 20, beta_old + safe_limit / 10};
    BwBetaDens d{m, C_k};
    return arms_sample(xinit, 4, beta_old - safe_limit, beta_old + safe_limit, d, urand, beta_new, env, cumulate);
}

}  // namespace hb
