// BayesW (Weibull survival) kernels -- first correct CUDA path for the per-marker loop of BayesW::runMpiGibbs_bW
// (reference src/BayesW.cpp:1480-1852).  One launch pair per synchronisation window:
//   k_bw_window : one CTA per window position: sums of vi = exp(alpha*eps - EuMasc) over the marker's genotype classes
//                 (partial_sum, :49-65, K8/K9), marginal likelihoods by adaptive Gauss-Hermite (:716-726, lanes = nodes),
//                 component cascade (:1536-1597) and the ARMS draw of beta (:1562-1582) by lane 0
//   k_bw_update : one CTA per slice: applies the changed markers in window order to epsilon (sparse_scaadd :1606-1624 and
//                 the sync :1799-1835) and refreshes the slice's sum of vi (K10)
// plus k_bw_reduce for the N-sums of exp inside the log-densities of mu and alpha (K11, :77-88, :132-142).
// epsilon lives in global memory here (L2 resident); the removal of the marker's own effect when beta_old != 0
// (:1499-1516: N exps per marker in the reference) is done algebraically: exp(alpha*(eps+delta)-g) = vi * exp(alpha*delta),
// with only three distinct deltas per marker.
#pragma once
#include "bayesw_math.cuh"
#include "brr_kernel.cuh"

namespace hb {

enum : int { BW_CHAIN = 0, BW_VISUMS = 1, BW_VECSUMS = 2 };

struct BwParams {
    uint32_t N, S, L, M;
    const uint64_t *rec;
    const double *mave, *sd, *sumfail;
    const int32_t *grp;
    double *E;                 // [S*L] epsilon without the scalar shift
    const double *shift_in;    // device scalar: true eps = E + *shift_in
    double *shift_out;
    double alpha;
    const double *vec;         // BW_VECSUMS: vector to sum over the classes (failure indicators)
    double *Vpart;             // [S] sum of vi over the slice
    double *vi;                // [S*L] vi = exp(alpha*eps - EuMasc) itself, refreshed by k_bw_update with the slice sums (K10)
    const int32_t *order;      // window-ordered local marker or -1
    const double *unif;        // window-ordered U(0,1)
    uint32_t base, W, T, t_first, j0;
    double *beta;
    int32_t *comp, *cass;
    const double *pi_L, *cVa, *sigmaG;  // [G*K], [G*(K-1)], [G]
    double sumSigmaG;
    uint32_t K, rule;
    uint32_t seed, iteration;
    uint32_t *chg_cnt;
    ChgEnt *chg_list;
    uint32_t *err;
    int mode;
    const double *beta_in;     // BW_VISUMS: beta_old per position (unit mode)
    double *out;               // unit modes: [W*4]
    // Window sequence on the device (src/BayesW.cpp:1633-1660: synchronise after sync_rate steps, or after every step while nothing
    // changes): ctl = { j0, since, n, n_sync, n_changed, n_windows }. The kernels of window w read ctl_in and the last one writes
    // ctl_out (two blocks in turn), so the host enqueues several windows without waiting for any of them; a window past the end of
    // the marker loop (j0 >= lmax) does nothing. NULL: base / W / j0 as given (unit modes).
    const uint32_t *ctl_in;
    uint32_t *ctl_out;
    uint32_t lmax, SR;
    double *delta;             // several GPUs: [S*L + 2] this GPU's epsilon change of the window (zeroed), then { its constant term, its
                               // number of changed markers }: summed over the GPUs (the reference's MPI_Allreduce of deltaEps,
                               // src/BayesW.cpp:1799-1835) and applied by k_bw_apply_delta; NULL on one GPU
};

__device__ __forceinline__ double bw_block_sum(double v, double *red) {
    v = warp_sum(v);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (uint32_t w = 0; w < (blockDim.x >> 5); w++) s += red[w];
    return s;
}

struct BwArmsRand {  // RNG spec v1: 31-bit integers from Philox, u = (r + 0.5) / 2^31 (form of src/BayesW_arms.cpp:913-918)
    uint32_t seed, task, iteration, j, idx;
    __device__ double operator()() {
        uint32_t w[4];
        philox4x32(idx >> 2, j, iteration, 0x41524D42u, seed, task, w);
        const uint32_t r = w[idx & 3u] >> 1;
        idx++;
        return ((double)r + 0.5) / 2147483648.0;
    }
};

// Warp-cooperative integration of the ARMS envelope. The draw itself is serial (lane 0), but after every new point the
// reference re-exponentiates and re-integrates the whole envelope (its maximum moves): one exp and one divide per point, the
// bulk of the draw's time on a single thread. Lane 0 lists the points in order and wakes the other lanes of its warp, which
// wait in bw_arms_helpers(); every lane evaluates the reference's expressions for its points; lane 0 adds the areas up in order.
struct BwArmsShared {
    ArmsEnvelope env;
    double area[kArmsPoints];
    int ord[kArmsPoints];
    int n, cmd;
};
__device__ __forceinline__ void bw_coop_ey(BwArmsShared &sh, uint32_t lane) {
    ArmsEnvelope &e = sh.env;
    for (int i = (int)lane; i < sh.n; i += 32) { const int q = sh.ord[i]; e.ey[q] = arms_expshift(e.y[q], e.ymax); }
}
__device__ __forceinline__ void bw_coop_area(BwArmsShared &sh, uint32_t lane) {
    ArmsEnvelope &e = sh.env;
    for (int i = (int)lane + 1; i < sh.n; i += 32) {
        const int q = sh.ord[i], l = sh.ord[i - 1];
        double a;
        if (e.x[l] == e.x[q]) a = 0.0;
        else if (fabs(e.y[q] - e.y[l]) < kArmsYEPS) a = 0.5 * (e.ey[q] + e.ey[l]) * (e.x[q] - e.x[l]);
        else a = ((e.ey[q] - e.ey[l]) / (e.y[q] - e.y[l])) * (e.x[q] - e.x[l]);
        sh.area[q] = a;
    }
}
struct BwCoopCumulate {  // called by lane 0 only
    BwArmsShared *sh;
    __device__ void operator()(ArmsEnvelope &e) const {
        int lm = 0;
        while (e.pl[lm] >= 0) lm = e.pl[lm];
        int n = 0;
        double ymax = e.y[lm];
        for (int q = lm; q >= 0; q = e.pr[q]) { sh->ord[n++] = q; if (e.y[q] > ymax) ymax = e.y[q]; }
        e.ymax = ymax;
        sh->n = n; sh->cmd = 1;
        __syncwarp();
        bw_coop_ey(*sh, 0);
        __syncwarp();
        bw_coop_area(*sh, 0);
        __syncwarp();
        e.cum[lm] = 0.0;
        for (int i = 1; i < n; i++) e.cum[sh->ord[i]] = e.cum[sh->ord[i - 1]] + sh->area[sh->ord[i]];
    }
};
__device__ __forceinline__ void bw_arms_helpers(BwArmsShared &sh, uint32_t lane) {  // lanes 1..31 of the drawing warp
    for (;;) {
        __syncwarp();
        if (sh.cmd == 0) break;
        bw_coop_ey(sh, lane);
        __syncwarp();
        bw_coop_area(sh, lane);
        __syncwarp();
    }
}
__device__ __forceinline__ void bw_arms_release(BwArmsShared &sh) {  // lane 0, when it needs no (more) help
    sh.cmd = 0;
    __syncwarp();
}

constexpr uint32_t kBwMaxS = 160;  // slices (<= CTAs of the sampler grid <= SMs)
__global__ void __launch_bounds__(256) k_bw_window(const BwParams P) {
    __shared__ double red[8];
    __shared__ uint32_t s_cum[kBwMaxS + 1], s_st[kBwMaxS], s_w1[kBwMaxS], s_w12[kBwMaxS];
    __shared__ BwArmsShared s_arms;
    const uint32_t p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t w_base = P.base, w_j0 = P.j0;
    if (P.ctl_in) {   // the window as the device left it
        w_j0 = P.ctl_in[0];
        if (w_j0 >= P.lmax || p >= P.ctl_in[2] * P.T) return;
        w_base = w_j0 * P.T;
    }
    const int32_t m = P.order[w_base + p];
    if (m < 0) return;
    const uint64_t rr = P.rec[m];
    const double shift = *P.shift_in;
    const double bconst = P.alpha * shift - kBwEuMasc;
    // what is summed over the marker's genotype classes: the chain reads vi as refreshed at the last synchronisation
    // (same expression, evaluated once per individual instead of once per non-zero); BW_VISUMS evaluates it for its alpha
    const double *vec = (P.mode == BW_VECSUMS) ? P.vec : ((P.mode == BW_CHAIN) ? P.vi : P.E);
    const bool take_exp = (P.mode == BW_VISUMS);
    double a1 = 0.0, a2 = 0.0, am = 0.0;
    if (rr & 1ull) {  // BED record: all slices are one array of S*L/32 words
        const uint64_t *ptr = reinterpret_cast<const uint64_t *>(rr & ~15ull);
        const uint32_t nw = P.S * (P.L / 32);
        for (uint32_t w = tid; w < nw; w += blockDim.x) {
            const uint64_t bits = ld_stream_u64(ptr + w);
            if (bits == ~0ull) continue;
            const double *e = vec + (size_t)32u * w;  // slice c starts at individual c*L = word c*L/32
            for (uint32_t t = 0; t < 32; t++) {
                const uint32_t code = (uint32_t)(bits >> (2u * t)) & 3u;
                if (code == 3u) continue;
                const double x = e[t];
                const double v = take_exp ? exp(P.alpha * x + bconst) : x;
                if (code == 2u) a1 += v; else if (code == 0u) a2 += v; else am += v;
            }
        }
    } else {  // sparse record: the words of all slice blocks as one index space (no slice-by-slice load chains)
        const uint8_t *bp = reinterpret_cast<const uint8_t *>(rr);
        const uint32_t *dir = reinterpret_cast<const uint32_t *>(bp);
        const uint64_t *payload = reinterpret_cast<const uint64_t *>(bp + dir_bytes(P.S));
        for (uint32_t c = tid; c < P.S; c += blockDim.x) {
            const uint32_t st = __ldg(dir + c * 3), n12 = __ldg(dir + c * 3 + 1), nm = __ldg(dir + c * 3 + 2);
            const uint32_t w1 = ((n12 & 0xFFFFu) + 3) / 4, w2 = ((n12 >> 16) + 3) / 4, wm = (nm + 3) / 4;
            s_st[c] = st; s_w1[c] = w1; s_w12[c] = w1 + w2; s_cum[c + 1] = w1 + w2 + wm;
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t a = 0;
            s_cum[0] = 0;
            for (uint32_t c = 0; c < P.S; c++) { a += s_cum[c + 1]; s_cum[c + 1] = a; }
        }
        __syncthreads();
        const uint32_t total = s_cum[P.S];
        for (uint32_t f = tid; f < total; f += blockDim.x) {
            uint32_t lo = 0, hi = P.S;  // slice c with cum[c] <= f < cum[c+1]
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (s_cum[mid] <= f) lo = mid; else hi = mid; }
            const uint32_t c = lo, w = f - s_cum[c];
            const uint64_t x4 = ld_stream_u64(payload + s_st[c] + w);
            const double *e = vec + (size_t)c * P.L;
            double sacc = 0.0;
#pragma unroll
            for (uint32_t t = 0; t < 4; t++) {
                const uint32_t idx = (uint32_t)(x4 >> (16u * t)) & 0xFFFFu;
                if (idx != P.L) {
                    const double x = e[idx];
                    sacc += take_exp ? exp(P.alpha * x + bconst) : x;
                }
            }
            if (w < s_w1[c]) a1 += sacc; else if (w < s_w12[c]) a2 += sacc; else am += sacc;
        }
    }
    const double S1 = bw_block_sum(a1, red), S2 = bw_block_sum(a2, red), SM = bw_block_sum(am, red);
    if (P.mode == BW_VECSUMS) {
        if (tid == 0) { P.out[p * 4 + 0] = S1; P.out[p * 4 + 1] = S2; P.out[p * 4 + 2] = SM; P.out[p * 4 + 3] = 0.0; }
        return;
    }
    if (warp != 0) return;
    // ---- sum of vi over everybody: slice partials in slice order
    double V = 0.0;
    for (uint32_t c = 0; c < P.S; c++) V += P.Vpart[c];
    const double beta_old = (P.mode == BW_VISUMS) ? P.beta_in[p] : P.beta[m];
    const double mean = P.mave[m], sd = P.sd[m];
    double vi_sum = V, vi_1 = S1, vi_2 = S2;
    if (beta_old != 0.0) {  // remove the marker's own effect (:1499-1516): three distinct shifts of epsilon
        const double sig_inv = 1 / sd;
        const double c0 = exp(P.alpha * (-(mean * sig_inv * beta_old)));
        const double c1 = exp(P.alpha * (beta_old * (1.0 - mean) * sig_inv));
        const double c2 = exp(P.alpha * (beta_old * (2.0 - mean) * sig_inv));
        vi_1 = c1 * S1;
        vi_2 = c2 * S2;
        vi_sum = c0 * (V - S1 - S2 - SM) + vi_1 + vi_2 + SM;
    }
    const double vi_0 = vi_sum - vi_1 - vi_2;  // :1525
    if (P.mode == BW_VISUMS) {
        if (lane == 0) { P.out[p * 4 + 0] = vi_sum; P.out[p * 4 + 1] = vi_1; P.out[p * 4 + 2] = vi_2; P.out[p * 4 + 3] = vi_0; }
        return;
    }
    // ---- marginal likelihoods (:716-726); lane i evaluates quadrature node i, sums in the reference's order
    const int g = P.grp[m];
    const uint32_t K = P.K, km1 = K - 1;
    BwMarker bm{P.alpha, P.sigmaG[g], P.sumfail[m], vi_sum, vi_0, vi_1, vi_2, mean, sd, mean / sd};
    double ML[kMaxMix];
    ML[0] = P.pi_L[g * K] * kBwSqrtPi;  // :1473, 1496
    const double exp_sum = (vi_1 * (1 - 2 * mean) + 4 * (1 - mean) * vi_2 + vi_sum * mean * mean) / (sd * sd);
    const int nq = HB_GH(n)[P.rule], qo = HB_GH(off)[P.rule];
    for (uint32_t k = 0; k < km1; k++) {
        const double Ck = P.cVa[g * km1 + k];
        const double sigma = 1.0 / sqrt(1 + P.alpha * P.alpha * bm.sigmaG * Ck * exp_sum);
        const double sq = sqrt(2 * Ck * bm.sigmaG);
        double t = 0.0;
        if ((int)lane < nq - 1) t = HB_GH(w)[qo + lane] * bw_gh_integrand(sigma * HB_GH(x)[qo + lane], bm, sq);
        double temp = __shfl_sync(0xffffffffu, t, 0);
        for (int i = 1; i < nq - 1; i++) temp += __shfl_sync(0xffffffffu, t, i);
        temp = temp + HB_GH(wc)[P.rule];
        ML[k + 1] = P.pi_L[g * K + k + 1] * (sigma * temp);
    }
    double MLsum = 0.0;
    for (uint32_t k = 0; k < K; k++) MLsum += ML[k];
    const double prob = P.unif[w_base + p];                                       // :1528
    double acum = ML[0] / MLsum;                                                  // :1536
    int comp = -1;
    for (uint32_t k = 0; k < K; k++) {
        if (prob <= acum) { comp = (int)k; break; }
        if ((k + 1) == km1) acum = 1;                                             // :1592-1593
        else acum += ML[k + 1] / MLsum;
    }
    if (lane != 0) { bw_arms_helpers(s_arms, lane); return; }  // the other lanes serve lane 0's envelope integrations
    double beta_new = beta_old;  // the cascade can fall through without a draw (as in the reference)
    if (comp == 0) beta_new = 0.0;
    else if (comp > 0) {
        BwArmsRand ur{P.seed, P.t_first + (p % P.T), P.iteration, w_j0 + p / P.T, 0u};
        const int rc = bw_sample_beta(bm, P.cVa[g * km1 + comp - 1], P.sumSigmaG, beta_old, ur, &beta_new, s_arms.env, BwCoopCumulate{&s_arms});
        if (rc != ARMS_OK) { atomicExch(P.err, (uint32_t)rc); bw_arms_release(s_arms); return; }
    }
    bw_arms_release(s_arms);
    if (comp >= 0) {
        atomicAdd(&P.cass[g * K + comp], 1);
        P.comp[m] = comp;
        P.beta[m] = beta_new;
    }
    const double dbeta = beta_old - beta_new;                                     // :1600
    if (dbeta != 0.0) {
        const uint32_t idx = atomicAdd(P.chg_cnt, 1u);
        ChgEnt en;
        en.p = p; en.m = (uint32_t)m; en.dbs = dbeta * (1 / sd); en.mave = mean; en.rec = rr;
        P.chg_list[idx] = en;
    }
}

// one CTA per slice; entries are applied in window order (global memory epsilon, L2 coherent accesses)
// one thread: the window sequence after a window with `cnt` changed markers (the host loop of round 1, src/BayesW.cpp:1633-1660)
__device__ __forceinline__ void bw_ctl_advance(const BwParams &P, uint32_t cnt) {
    uint32_t j0 = P.ctl_in[0], since = P.ctl_in[1];
    const uint32_t n = P.ctl_in[2];
    uint32_t nsync = P.ctl_in[3], nchg = P.ctl_in[4];
    if (cnt > 0) { since = 0; nsync++; nchg += cnt; } else { since += n; }
    j0 += n;
    P.ctl_out[0] = j0; P.ctl_out[1] = since;
    P.ctl_out[2] = (j0 >= P.lmax) ? 0u : ((since >= P.SR) ? 1u : min(P.SR - since, P.lmax - j0));
    P.ctl_out[3] = nsync; P.ctl_out[4] = nchg; P.ctl_out[5] = P.ctl_in[5] + 1u;
}
// a window past the end: the control block and the shift are handed on unchanged
__device__ __forceinline__ bool bw_ctl_noop(const BwParams &P, bool hand_on) {
    if (!P.ctl_in || P.ctl_in[0] < P.lmax) return false;
    if (hand_on && blockIdx.x == 0 && threadIdx.x == 0) {
        for (int i = 0; i < 6; i++) P.ctl_out[i] = P.ctl_in[i];
        *P.shift_out = *P.shift_in;
    }
    return true;
}

__global__ void __launch_bounds__(512) k_bw_update(const BwParams P) {
    __shared__ uint32_t ps[kMaxMerged];
    __shared__ double red[16];
    __shared__ const uint64_t *s_ptr[64];
    __shared__ uint32_t s_nw[64], s_b1[64], s_b2[64];
    __shared__ double s_dbs[64], s_mave[64];
    const uint32_t c = blockIdx.x, tid = threadIdx.x, L = P.L;
    if (bw_ctl_noop(P, P.delta == nullptr)) return;   // (several GPUs: k_bw_apply_delta hands the control block on)
    const uint32_t n = *P.chg_cnt;
    const double shift_old = *P.shift_in;
    double off = 0.0;
    double *E = (P.delta ? P.delta : P.E) + (size_t)c * L;
    if (n > kMaxMerged) {
        if (c == 0 && tid == 0) atomicExch(P.err, 4u);
        return;
    }
    if (n > 0) {
        ChgEnt en[2];
        uint32_t rank[2] = {0, 0};
        for (uint32_t u = 0; u < 2; u++) {
            const uint32_t i = tid + u * blockDim.x;
            en[u].p = 0xFFFFFFFFu;
            if (i < n) { en[u] = ld_chg_ent(P.chg_list + i); ps[i] = en[u].p; }
        }
        __syncthreads();
        for (uint32_t u = 0; u < 2; u++)
            if (en[u].p != 0xFFFFFFFFu)
                for (uint32_t i = 0; i < n; i++) rank[u] += (ps[i] < en[u].p) ? 1u : 0u;
        for (uint32_t x0 = 0; x0 < n; x0 += 64) {
            const uint32_t nx = min(64u, n - x0);
            for (uint32_t u = 0; u < 2; u++) {
                if (en[u].p != 0xFFFFFFFFu && rank[u] >= x0 && rank[u] < x0 + nx) {
                    const Blk b = decode_block(en[u].rec, c, P.S, L);
                    const uint32_t x = rank[u] - x0;
                    s_ptr[x] = b.ptr; s_nw[x] = b.nw; s_b1[x] = b.b1; s_b2[x] = b.b2; s_dbs[x] = en[u].dbs; s_mave[x] = en[u].mave;
                }
            }
            __syncthreads();
            for (uint32_t x = 0; x < nx; x++) {
                const double dbs = s_dbs[x], mave = s_mave[x];
                for (uint32_t w = tid; w < s_nw[x]; w += blockDim.x) {
                    const uint64_t bits = ld_stream_u64(s_ptr[x] + w);
                    if (s_b1[x] == 0xFFFFFFFFu) {
                        if (bits == ~0ull) continue;
                        for (uint32_t t = 0; t < 32; t++) {
                            const uint32_t code = (uint32_t)(bits >> (2u * t)) & 3u;
                            if (code == 3u) continue;
                            const double d = ((code == 0u) ? 2.0 : ((code == 2u) ? 1.0 : mave)) * dbs;
                            __stcg(E + 32u * w + t, __ldcg(E + 32u * w + t) + d);
                        }
                    } else {
                        const double d = ((w < s_b1[x]) ? 1.0 : ((w < s_b2[x]) ? 2.0 : mave)) * dbs;
                        // the four individuals of a word are distinct: their loads go out together, then the stores
                        uint32_t idx[4];
                        double v[4];
#pragma unroll
                        for (uint32_t t = 0; t < 4; t++) {
                            idx[t] = (uint32_t)(bits >> (16u * t)) & 0xFFFFu;
                            v[t] = (idx[t] != L) ? __ldcg(E + idx[t]) : 0.0;
                        }
#pragma unroll
                        for (uint32_t t = 0; t < 4; t++)
                            if (idx[t] != L) __stcg(E + idx[t], v[t] + d);
                    }
                }
                off = fma(-mave, dbs, off);
                __syncthreads();
            }
        }
    }
    if (P.delta) {   // several GPUs: the change itself is the result; k_bw_apply_delta adds the sum over the GPUs to epsilon
        if (c == 0 && tid == 0) { P.delta[(size_t)P.S * L] = off; P.delta[(size_t)P.S * L + 1] = (double)n; }
        return;
    }
    const double shift_new = shift_old + off;
    if (c == 0 && tid == 0) { *P.shift_out = shift_new; if (P.ctl_in) bw_ctl_advance(P, n); }
    // refresh the slice's sum of vi (:1832-1834)
    double v = 0.0;
    const double bconst = P.alpha * shift_new - kBwEuMasc;
    for (uint32_t i = tid; i < L; i += blockDim.x) {
        if ((size_t)c * L + i < P.N) {
            const double x = exp(P.alpha * __ldcg(E + i) + bconst);
            P.vi[(size_t)c * L + i] = x;  // gathered by k_bw_window until the next synchronisation
            v += x;
        }
    }
    const double s = bw_block_sum(v, red);
    if (tid == 0) P.Vpart[c] = s;
}

// Several GPUs: epsilon += the epsilon changes of all GPUs (P.delta after the all-reduce), then the slice's vi and their sum as
// in k_bw_update. Every GPU adds the same numbers to the same epsilon: the replicas stay bit-identical.
__global__ void __launch_bounds__(512) k_bw_apply_delta(const BwParams P) {
    __shared__ double red[16];
    const uint32_t c = blockIdx.x, tid = threadIdx.x, L = P.L;
    if (bw_ctl_noop(P, true)) return;
    double *E = P.E + (size_t)c * L;
    const double *D = P.delta + (size_t)c * L;
    const double shift_new = *P.shift_in + P.delta[(size_t)P.S * L];
    if (c == 0 && tid == 0) {
        *P.shift_out = shift_new;
        if (P.ctl_in) bw_ctl_advance(P, (uint32_t)llrint(P.delta[(size_t)P.S * L + 1]));   // changed markers over all GPUs
    }
    double v = 0.0;
    const double bconst = P.alpha * shift_new - kBwEuMasc;
    for (uint32_t i = tid; i < L; i += blockDim.x) {
        const double e = __ldcg(E + i) + D[i];
        __stcg(E + i, e);
        if ((size_t)c * L + i < P.N) {
            const double x = exp(P.alpha * e + bconst);
            P.vi[(size_t)c * L + i] = x;
            v += x;
        }
    }
    const double s = bw_block_sum(v, red);
    if (tid == 0) P.Vpart[c] = s;
}

// N-sums for the log-densities of mu / alpha (K11): mode 0: sum exp(a*eps + b); mode 1: sum eps*f
__global__ void __launch_bounds__(512) k_bw_reduce(const double *__restrict__ E, const double *__restrict__ f, uint32_t N, uint32_t L,
                                                   double shift, double a, double b, int mode, double *__restrict__ part) {
    __shared__ double red[16];
    const uint32_t c = blockIdx.x;
    double v = 0.0;
    for (uint32_t i = threadIdx.x; i < L; i += blockDim.x) {
        const size_t gi = (size_t)c * L + i;
        if (gi < N) {
            const double e = E[gi] + shift;
            v += (mode == 0) ? exp(a * e + b) : e * f[gi];
        }
    }
    const double s = bw_block_sum(v, red);
    if (threadIdx.x == 0) part[c] = s;
}

// unit-level entry points for parity tests: one thread evaluates the scalar functions on the device
// Fixed effects (src/BayesW.cpp:119-129, 1366-1413): sum_i exp(a*eps_i + b*x_i + c0) with x = one covariate column on the
// slice layout (gamma_dens evaluated at gamma: a = alpha, b = alpha*(gamma_old - gamma), c0 = -EuMasc), slice partials.
__global__ void __launch_bounds__(512) k_bw_reduce_cov(const double *__restrict__ E, const double *__restrict__ x, uint32_t N, uint32_t L,
                                                       double shift, double a, double b, double c0, double *__restrict__ part) {
    __shared__ double red[16];
    const uint32_t c = blockIdx.x;
    double v = 0.0;
    for (uint32_t i = threadIdx.x; i < L; i += blockDim.x) {
        const size_t gi = (size_t)c * L + i;
        if (gi < N) v += exp(a * (E[gi] + shift) + b * x[gi] + c0);
    }
    const double s = bw_block_sum(v, red);
    if (threadIdx.x == 0) part[c] = s;
}
// eps += d * x (the residual after a fixed effect moved from gamma_old to gamma_old - d, :1408-1410)
__global__ void k_bw_axpy_cov(double *__restrict__ E, const double *__restrict__ x, size_t n, double d) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) E[i] += d * x[i];
}

__global__ void k_bw_unit_marginal(BwMarker m, int rule, const double *prior, const double *cVa, int km1, double *post) {
    bw_marginal_likelihoods(rule, prior, cVa, km1, m, post);
}
__global__ void k_bw_unit_arms(BwMarker m, double Ck, double sumSigmaG, double beta_old, uint32_t seed, uint32_t task, uint32_t iteration,
                               uint32_t j, double *out) {  // one warp: lane 0 draws, the others help (as in k_bw_window)
    __shared__ BwArmsShared sh;
    if (threadIdx.x != 0) { bw_arms_helpers(sh, threadIdx.x); return; }
    BwArmsRand ur{seed, task, iteration, j, 0u};
    double bn = 0.0;
    const int rc = bw_sample_beta(m, Ck, sumSigmaG, beta_old, ur, &bn, sh.env, BwCoopCumulate{&sh});
    bw_arms_release(sh);
    out[0] = bn; out[1] = (double)rc; out[2] = (double)sh.env.neval; out[3] = (double)ur.idx;
}

}  // namespace hb
