// Host side of RNG spec v1 (DESIGN.md "Draw tape").  The reference draws from
// Boost.Random on boost::mt19937 (src/distributions_boost.cpp:57-136); Boost is not
// available and its streams are unpinned, so the product defines its own streams:
// MT19937 raw words (std::mt19937 == boost::mt19937 bit for bit) with explicitly
// specified transforms (no std:: distributions, which are implementation defined).
#pragma once
#include <cmath>
#include <cstdint>
#include <random>
#include <vector>

namespace hb {

struct HostRng {
    std::mt19937 eng;
    explicit HostRng(uint32_t seed = 0) : eng(seed) {}
    void seed(uint32_t s) { eng.seed(s); }
    double res53() {  // [0,1), 53 bits: (a>>5, b>>6)
        uint32_t a = (uint32_t)eng() >> 5, b = (uint32_t)eng() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
    double normal() {  // Box-Muller, no caching of the second variate
        double u1 = res53(), u2 = res53();
        return std::sqrt(-2.0 * std::log(1.0 - u1)) * std::cos(6.283185307179586476925 * u2);
    }
    double gamma(double a) {  // Marsaglia-Tsang, scale 1
        if (a < 1.0) {
            double g = gamma(a + 1.0);
            double u = 1.0 - res53();
            return g * std::pow(u, 1.0 / a);
        }
        const double d = a - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
        for (;;) {
            double x = normal();
            double v = 1.0 + c * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            double u = 1.0 - res53();
            if (u < 1.0 - 0.0331 * (x * x) * (x * x)) return d * v;
            if (std::log(u) < 0.5 * x * x + d * (1.0 - v + std::log(v))) return d * v;
        }
    }
    // src/distributions_boost.cpp:92-94, 112-114
    double inv_gamma(double shape, double scale) { return 1.0 / (gamma(shape) * (1.0 / scale)); }
    double inv_scaled_chisq(double dof, double scale) { return inv_gamma(0.5 * dof, 0.5 * dof * scale); }
    // src/distributions_boost.cpp:97-103
    double inv_gamma_rate(double shape, double rate) { return 1.0 / (gamma(shape) * (1.0 / rate)); }
    template <class T>
    void shuffle(T *a, int n) {  // Fisher-Yates
        for (int i = n - 1; i >= 1; i--) {
            int k = (int)(res53() * (double)(i + 1));
            T t = a[i]; a[i] = a[k]; a[k] = t;
        }
    }
};

// Philox4x32-10 on the host (same function as the device's philox4x32 in common.cuh)
inline void philox4x32_host(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ARMS uniforms of RNG spec v1: 31-bit integers r, u = (r + 0.5) / 2^31 (the form of src/BayesW_arms.cpp:913-918)
struct ArmsStream {
    uint32_t seed, task, c1, c2, tag, idx;
    double operator()() {
        uint32_t w[4];
        philox4x32_host(idx >> 2, c1, c2, tag, seed, task, w);
        const uint32_t r = w[idx & 3u] >> 1;
        idx++;
        return ((double)r + 0.5) / 2147483648.0;
    }
};

}  // namespace hb
