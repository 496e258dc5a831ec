// BayesRRm marker-loop kernel (one cooperative, persistent launch per Gibbs
// iteration).  Replaces the `for (j<lmax)` body of BayesRRm::runMpiGibbs
// (reference src/BayesRRm.cpp:1709-2490): sparse_dotprod (:316-342) / LUT dot
// (:1757-1809), the mixture draw (:1744-1921), sparse_scaadd (:250-281) / LUT
// deltaEps (:1976-2010) and the epsilon synchronisation (:2044-2488).
//
// Design: between two synchronisations hydra's tasks all read a
// STALE epsilon, so every marker of a sync window (sync_rate steps x T tasks) is
// independent of the others.  The kernel therefore processes a whole window in
// parallel:
//   * individuals are cut in S slices; CTA (r,c) keeps slice c of the residual in
//     shared memory (r = replica group, R = gridDim/S groups share the window);
//   * phase A: warps stream the slice block of each marker (u16 local indices or
//     2-bit BED bytes, coalesced 64-bit loads) and gather from shared memory;
//   * the CTA that delivers the last of the S slice partials of a marker sums
//     them in slice order and draws the mixture component and beta on the device;
//   * ONE grid barrier per window, after which every CTA applies the window's
//     non-zero deltaBetas to its slice, in window order (deterministic).
#pragma once
#include "common.cuh"

namespace hb {

enum : int { MODE_CHAIN = 0, MODE_DOT = 1, MODE_SCAADD = 2 };

struct BrrParams {
    // layout
    uint32_t N, S, L, R, M;
    const uint64_t *rec;   // [M] record address | BED flag
    const double *mave;    // [M]
    const double *mstd;    // [M] inverse sd (BayesRRm)
    const int32_t *grp;    // [M]
    // epsilon
    const double *E_in;    // [S*L] common residual without base terms
    double *E_out;         // [S*L]
    double shift_in;       // folded into E at load (base terms of the previous launch)
    double *off_out;       // base-term accumulator of this launch (scalar)
    double *slice_sum_out; // [S]  sum of E over the slice (i < N)
    double *slice_sq_out;  // [S]  sum of E^2
    // marker state
    double *beta;          // [M]
    int32_t *comp;         // [M]
    double *acum;          // [M]
    int32_t *cass;         // [G*K]
    // per-iteration inputs, window-ordered: q = j*T + task
    const int32_t *order;  // [lmax*T] local marker or -1
    const double *u;       // [lmax*T]
    const double *z;       // [lmax*T]
    uint32_t T, SR, lmax, K, G;
    // hyper-parameter tables [G*K]
    const double *logPi, *chalf, *denom, *sdk;
    const uint8_t *grp_active; // [G]
    double i_2sigE, dNm1;
    // scratch
    double *partial;       // [Wmax*S]
    uint32_t *cnt;         // [Wmax] arrival counters (monotonic)
    double *dB;            // [2*Wmax] deltaBeta*mstd per window position, double buffered
    uint32_t Wmax;
    uint32_t *bar;         // grid barrier counter
    unsigned long long *stats; // [8]: 0 nsync 1 nwindows 2 nnz dot 3 nnz upd 4 bed markers 5 changed
    // unit modes
    int mode;
    double *num_out;       // MODE_DOT: [W]
};

struct ItemTab {
    const uint64_t *ptr[kTabCap];
    double mave[kTabCap];
    double part[kTabCap];
    uint32_t nw[kTabCap];   // u64 words of the slice block
    uint32_t b1[kTabCap];   // first word of class 2   (0xFFFFFFFF = BED block)
    uint32_t b2[kTabCap];   // first word of class "missing"
    int32_t m[kTabCap];
};

// ---- slice block dot product: sum_w weight(w) * sum_4 E_s[idx] ------------------
__device__ __forceinline__ double dot_sparse_block(const uint64_t *__restrict__ ptr, uint32_t nw, uint32_t b1,
                                                   uint32_t b2, double mave, const double *__restrict__ E_s,
                                                   uint32_t lane) {
    double acc = 0.0;
    uint32_t w = lane;
    // two independent 64-bit loads in flight per lane
    for (; w + 32 < nw; w += 64) {
        uint64_t x0 = ld_stream_u64(ptr + w), x1 = ld_stream_u64(ptr + w + 32);
        double wt0 = (w < b1) ? 1.0 : ((w < b2) ? 2.0 : mave);
        double wt1 = (w + 32 < b1) ? 1.0 : ((w + 32 < b2) ? 2.0 : mave);
        double s0 = (E_s[x0 & 0xFFFFu] + E_s[(x0 >> 16) & 0xFFFFu]) + (E_s[(x0 >> 32) & 0xFFFFu] + E_s[x0 >> 48]);
        double s1 = (E_s[x1 & 0xFFFFu] + E_s[(x1 >> 16) & 0xFFFFu]) + (E_s[(x1 >> 32) & 0xFFFFu] + E_s[x1 >> 48]);
        acc = fma(wt0, s0, acc);
        acc = fma(wt1, s1, acc);
    }
    if (w < nw) {
        uint64_t x0 = ld_stream_u64(ptr + w);
        double wt0 = (w < b1) ? 1.0 : ((w < b2) ? 2.0 : mave);
        double s0 = (E_s[x0 & 0xFFFFu] + E_s[(x0 >> 16) & 0xFFFFu]) + (E_s[(x0 >> 32) & 0xFFFFu] + E_s[x0 >> 48]);
        acc = fma(wt0, s0, acc);
    }
    return acc;
}

// BED block: lane owns 32 consecutive individuals per 64-bit word and walks them in a
// rotated order so that the 16 lanes of a half-warp hit 16 distinct shared-memory banks.
__device__ __forceinline__ double dot_bed_block(const uint64_t *__restrict__ ptr, uint32_t nw, double mave,
                                                const double *__restrict__ E_s, uint32_t lane) {
    double acc = 0.0;
    for (uint32_t w = lane; w < nw; w += 32) {
        const uint64_t bits = ld_stream_u64(ptr + w);
        if (bits == ~0ull) continue;  // 32 x genotype 0
        const double *e = E_s + 32u * w;
#pragma unroll 8
        for (uint32_t t = 0; t < 32; t++) {
            const uint32_t idx = (t + lane) & 31u;
            const uint32_t code = (uint32_t)(bits >> (2u * idx)) & 3u;
            const double v = e[idx];
            // 00 -> 2, 10 -> 1, 01 -> missing (weight mave), 11 -> 0
            const double wt = (code == 0u) ? 2.0 : ((code == 2u) ? 1.0 : ((code == 1u) ? mave : 0.0));
            acc = fma(wt, v, acc);
        }
    }
    return acc;
}

// ---- epsilon update of one marker on the CTA's slice (all threads) --------------
__device__ __forceinline__ void apply_marker(uint64_t r, uint32_t c, uint32_t S, uint32_t L, double dbs,
                                             double mave, double *__restrict__ E_s) {
    if (r & 1ull) {
        const uint64_t *ptr = reinterpret_cast<const uint64_t *>(r & ~15ull) + (size_t)c * (L / 32);
        for (uint32_t i = threadIdx.x; i < L; i += blockDim.x) {
            const uint64_t bits = ptr[i >> 5];
            const uint32_t code = (uint32_t)(bits >> (2u * (i & 31u))) & 3u;
            if (code != 3u) {
                const double wt = (code == 0u) ? 2.0 : ((code == 2u) ? 1.0 : mave);
                E_s[i] += wt * dbs;
            }
        }
    } else {
        const uint8_t *base = reinterpret_cast<const uint8_t *>(r);
        const uint32_t *dir = reinterpret_cast<const uint32_t *>(base) + c * 3;
        const uint32_t st = dir[0], n1 = dir[1] & 0xFFFFu, n2 = dir[1] >> 16, nm = dir[2];
        const uint16_t *blk = reinterpret_cast<const uint16_t *>(base + dir_bytes(S)) + (size_t)st * 4;
        const uint32_t o2 = ((n1 + 3) / 4) * 4, om = o2 + ((n2 + 3) / 4) * 4;
        const double d1 = dbs, d2 = 2.0 * dbs, dm = mave * dbs;
        // indices are unique inside one marker: no write conflicts
        for (uint32_t e = threadIdx.x; e < n1; e += blockDim.x) E_s[blk[e]] += d1;
        for (uint32_t e = threadIdx.x; e < n2; e += blockDim.x) E_s[blk[o2 + e]] += d2;
        for (uint32_t e = threadIdx.x; e < nm; e += blockDim.x) E_s[blk[om + e]] += dm;
    }
}

__device__ __forceinline__ double block_sum(double v, double *red /*[32]*/) {
    v = warp_sum(v);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    const uint32_t nw = blockDim.x >> 5;
    for (uint32_t w = 0; w < nw; w++) s += red[w];  // fixed order: identical in every CTA
    return s;
}

// ---- mixture draw of one marker (src/BayesRRm.cpp:1721-1933) ---------------------
// Executed by the thread that delivered the last slice partial of window position p.
__device__ void draw_marker(const BrrParams &P, uint32_t p, uint32_t q, int32_t m, uint32_t dbuf) {
    // sum the S slice partials in slice order (deterministic)
    const double *pp = P.partial + (size_t)p * P.S;
    double sum = 0.0;
    for (uint32_t c = 0; c < P.S; c++) sum += __ldcg(pp + c);
    const double mstd = P.mstd[m];
    if (P.mode == MODE_DOT) {
        P.num_out[q] = __dmul_rn(mstd, sum);
        P.dB[(size_t)dbuf * P.Wmax + p] = 0.0;
        return;
    }
    const int g = P.grp[m];
    const uint32_t K = P.K;
    const double beta_old = P.beta[m];
    double beta_new = 0.0, acum = 1.0;
    int comp = 0;
    if (P.grp_active[g]) {
        // num = mstd*(...) ; num += beta*(N-1)            (:1809/:316-342, :1855)
        const double num = __dadd_rn(__dmul_rn(mstd, sum), __dmul_rn(beta_old, P.dNm1));
        double logL[kMaxMix], muk[kMaxMix];
        const double *lp = P.logPi + g * K, *ch = P.chalf + g * K, *dn = P.denom + g * K;
        logL[0] = lp[0];
        muk[0] = 0.0;
        for (uint32_t k = 1; k < K; k++) {
            muk[k] = num / dn[k];                                                    // :1859
            logL[k] = __dadd_rn(__dadd_rn(lp[k], -ch[k]), __dmul_rn(__dmul_rn(muk[k], num), P.i_2sigE));  // :1874-1876
        }
        const double prob = P.u[q];                                                  // :1880
        bool big = false;
        for (uint32_t k = 1; k < K; k++) big |= (fabs(logL[k] - logL[0]) > 700.0);   // :1884
        if (big) acum = 0.0;
        else {
            double s = 0.0;
            for (uint32_t k = 0; k < K; k++) s += exp(logL[k] - logL[0]);
            acum = 1.0 / s;
        }
        const double acum0 = acum;
        for (uint32_t k = 0; k < K; k++) {                                           // :1894-1921
            if (prob <= acum || k == K - 1) {
                if (k > 0) beta_new = __dadd_rn(muk[k], __dmul_rn(P.sdk[g * K + k], P.z[q]));  // :1901
                comp = (int)k;
                break;
            } else {
                bool big2 = false;
                for (uint32_t i = k + 1; i < K; i++) big2 |= (fabs(logL[i] - logL[k + 1]) > 700.0);
                if (!big2) {
                    double s = 0.0;
                    for (uint32_t i = 0; i < K; i++) s += exp(logL[i] - logL[k + 1]);
                    acum += 1.0 / s;
                }
            }
        }
        acum = acum0;                                                                // Acum(marker), :1892
        atomicAdd(&P.cass[g * K + comp], 1);                                         // :1904
    }
    P.beta[m] = beta_new;
    if (P.grp_active[g]) P.comp[m] = comp;  // :1924-1925 leave components untouched
    P.acum[m] = acum;
    const double dbeta = beta_old - beta_new;                                        // :1933
    P.dB[(size_t)dbuf * P.Wmax + p] = (dbeta != 0.0) ? __dmul_rn(dbeta, mstd) : 0.0;
    if (dbeta != 0.0) atomicAdd(&P.stats[5], 1ull);
}

// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) k_brr_iteration(const BrrParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *E_s = reinterpret_cast<double *>(smem_raw);                 // [L+1], slot L = dummy for PAD
    ItemTab *tab = reinterpret_cast<ItemTab *>(smem_raw + (((size_t)P.L + 2) * 8 + 15) / 16 * 16);
    __shared__ double red[32];
    __shared__ uint32_t work_ctr;
    __shared__ uint32_t chg_n;
    __shared__ uint32_t chg_p[kThreads];
    __shared__ uint32_t chg_base[33];

    const uint32_t S = P.S, L = P.L, R = P.R;
    const uint32_t c = blockIdx.x % S, r = blockIdx.x / S;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nctas = gridDim.x;

    // ---- load the slice (fold the base terms of the previous launch) -------------
    for (uint32_t i = tid; i < L; i += blockDim.x) {
        const uint32_t gi = c * L + i;
        E_s[i] = (gi < P.N) ? (P.E_in[gi] + P.shift_in) : 0.0;
    }
    if (tid == 0) E_s[L] = 0.0;
    __syncthreads();
    double slice_sum;
    {
        double v = 0.0;
        for (uint32_t i = tid; i < L; i += blockDim.x) v += E_s[i];
        slice_sum = block_sum(v, red);
    }

    uint32_t bar_target = 0;
    double off = 0.0;
    uint32_t j0 = 0, since = 0, win = 0;
    unsigned long long nnz_dot = 0, nnz_upd = 0, n_bed = 0, n_sync = 0;
    const uint32_t SR = (P.SR == 0) ? 1u : P.SR;

    while (j0 < P.lmax) {
        const uint32_t n = (P.mode != MODE_CHAIN) ? (P.lmax - j0)
                                                  : ((since >= SR) ? 1u : min(SR - since, P.lmax - j0));
        const uint32_t W = n * P.T, base = j0 * P.T;
        const uint32_t dbuf = win & 1u;
        const uint32_t n_items = (W > r) ? (W - r + R - 1) / R : 0;

        if (P.mode != MODE_SCAADD) {
            for (uint32_t k0 = 0; k0 < n_items; k0 += kTabCap) {
                const uint32_t nk = min((uint32_t)kTabCap, n_items - k0);
                // ---- 1. item table: one thread per window position of this group ------
                for (uint32_t k = tid; k < nk; k += blockDim.x) {
                    const uint32_t p = r + R * (k0 + k);
                    const int32_t m = P.order[base + p];
                    tab->m[k] = m;
                    tab->nw[k] = 0;
                    if (m >= 0) {
                        const uint64_t rr = P.rec[m];
                        tab->mave[k] = P.mave[m];
                        if (rr & 1ull) {
                            tab->ptr[k] = reinterpret_cast<const uint64_t *>(rr & ~15ull) + (size_t)c * (L / 32);
                            tab->nw[k] = L / 32;
                            tab->b1[k] = 0xFFFFFFFFu;
                            tab->b2[k] = 0xFFFFFFFFu;
                        } else {
                            const uint8_t *bp = reinterpret_cast<const uint8_t *>(rr);
                            const uint32_t *dir = reinterpret_cast<const uint32_t *>(bp) + c * 3;
                            const uint32_t st = dir[0], n12 = dir[1], nm = dir[2];
                            const uint32_t w1 = ((n12 & 0xFFFFu) + 3) / 4, w2 = ((n12 >> 16) + 3) / 4, wm = (nm + 3) / 4;
                            tab->ptr[k] = reinterpret_cast<const uint64_t *>(bp + dir_bytes(S)) + st;
                            tab->b1[k] = w1;
                            tab->b2[k] = w1 + w2;
                            tab->nw[k] = w1 + w2 + wm;
                        }
                    } else if (P.mode == MODE_CHAIN) {
                        P.dB[(size_t)dbuf * P.Wmax + p] = 0.0;  // padded task step (:2029-2034); every slice CTA writes 0
                    }
                }
                if (tid == 0) work_ctr = 0;
                __syncthreads();
                // ---- 2. phase A: warps pull items -------------------------------------
                for (;;) {
                    uint32_t k = 0;
                    if (lane == 0) k = atomicAdd(&work_ctr, 1u);
                    k = __shfl_sync(0xffffffffu, k, 0);
                    if (k >= nk) break;
                    const uint32_t nw = tab->nw[k];
                    double acc = 0.0;
                    if (tab->m[k] >= 0) {
                        if (tab->b1[k] == 0xFFFFFFFFu) {
                            acc = dot_bed_block(tab->ptr[k], nw, tab->mave[k], E_s, lane);
                            if (lane == 0) n_bed++;
                        } else {
                            acc = dot_sparse_block(tab->ptr[k], nw, tab->b1[k], tab->b2[k], tab->mave[k], E_s, lane);
                            if (lane == 0) nnz_dot += 4ull * nw;
                        }
                        acc = warp_sum(acc);
                    }
                    if (lane == 0) tab->part[k] = acc;
                }
                __syncthreads();
                // ---- 3. publish slice partials; last arriver draws ---------------------
                for (uint32_t k = tid; k < nk; k += blockDim.x) {
                    const int32_t m = tab->m[k];
                    if (m < 0) continue;
                    const uint32_t p = r + R * (k0 + k);
                    // partial of num/mstd: sum_1 + 2 sum_2 + mave*sum_M - mave*sum_slice   (:327-339)
                    const double val = fma(-tab->mave[k], slice_sum, tab->part[k]);
                    __stcg(P.partial + (size_t)p * S + c, val);
                    __threadfence();
                    const uint32_t old = atomicAdd(P.cnt + p, 1u);
                    if ((old + 1u) % S == 0u) {
                        __threadfence();
                        draw_marker(P, p, base + p, m, dbuf);
                    }
                }
                __syncthreads();  // table reuse
            }
            if (P.mode == MODE_DOT) break;  // single window, nothing to apply
            // ---- 4. the one grid barrier of the window --------------------------------
            grid_barrier(P.bar, bar_target, nctas);
        }

        // ---- 5. apply the window's non-zero deltaBetas in window order ---------------
        bool any = false;
        const double *dB = P.dB + (size_t)dbuf * P.Wmax;
        for (uint32_t p0 = 0; p0 < W; p0 += blockDim.x) {
            const uint32_t p = p0 + tid;
            const double d = (p < W) ? __ldcg(dB + p) : 0.0;
            const bool ch = (d != 0.0);
            const uint32_t bal = __ballot_sync(0xffffffffu, ch);
            if (lane == 0) chg_base[warp] = __popc(bal);
            __syncthreads();
            if (tid == 0) {
                uint32_t a = 0;
                for (uint32_t w = 0; w < (blockDim.x >> 5); w++) { uint32_t t = chg_base[w]; chg_base[w] = a; a += t; }
                chg_n = a;
            }
            __syncthreads();
            if (ch) chg_p[chg_base[warp] + __popc(bal & ((1u << lane) - 1u))] = p;
            __syncthreads();
            const uint32_t nchg = chg_n;
            for (uint32_t x = 0; x < nchg; x++) {
                const uint32_t pc = chg_p[x];
                const int32_t m = P.order[base + pc];
                const double dbs = __ldcg(dB + pc);
                const double mv = P.mave[m];
                const uint64_t rr = P.rec[m];
                apply_marker(rr, c, S, L, dbs, mv, E_s);
                off = fma(-mv, dbs, off);  // base term -mave*mstd*deltaBeta of every individual (:265-267)
                if (tid == 0 && !(rr & 1ull)) {
                    const uint32_t *dir = reinterpret_cast<const uint32_t *>(rr) + c * 3;
                    nnz_upd += (dir[1] & 0xFFFFu) + (dir[1] >> 16) + dir[2];
                }
                __syncthreads();
            }
            any |= (nchg > 0);
            __syncthreads();
        }
        if (any) {
            if (tid == 0) E_s[L] = 0.0;  // PAD slot
            __syncthreads();
            double v = 0.0;
            for (uint32_t i = tid; i < L; i += blockDim.x) v += E_s[i];
            slice_sum = block_sum(v, red);
            since = 0;
            n_sync++;
        } else {
            since += n;
        }
        j0 += n;
        win++;
    }

    // ---- epilogue: group 0 stores the slices and their sums ----------------------------
    if (P.mode != MODE_DOT) {
        if (r == 0) {
            double v = 0.0, v2 = 0.0;
            for (uint32_t i = tid; i < L; i += blockDim.x) {
                const uint32_t gi = c * L + i;
                const double e = E_s[i];
                P.E_out[gi] = e;
                if (gi < P.N) { v += e; v2 += e * e; }
            }
            const double s1 = block_sum(v, red);
            const double s2 = block_sum(v2, red);
            if (tid == 0) { P.slice_sum_out[c] = s1; P.slice_sq_out[c] = s2; }
        }
        if (blockIdx.x == 0 && tid == 0) {
            *P.off_out = off;
            P.stats[0] = n_sync;
            P.stats[1] = win;
        }
    }
    // traffic counters (lane 0 of every warp holds a share)
    if (lane == 0) {
        if (nnz_dot) atomicAdd(&P.stats[2], nnz_dot);
        if (n_bed && c == 0) atomicAdd(&P.stats[4], n_bed);
    }
    if (tid == 0 && r == 0 && nnz_upd) atomicAdd(&P.stats[3], nnz_upd);
}

// ---------------------------------------------------------------------------------
// Window-ordered inputs of one iteration: order[q], u[q], z[q] with q = j*T + t.
// perm holds the task-local marker order of every local task block (concatenated);
// ut/zt != NULL: positional tape values indexed like perm; else RNG spec v1 (Philox).
__global__ void k_window_order(const int32_t *__restrict__ perm, const double *__restrict__ ut,
                               const double *__restrict__ zt, const int32_t *__restrict__ task_len,
                               const int32_t *__restrict__ task_off, uint32_t T, uint32_t lmax, uint32_t seed,
                               uint32_t iteration, uint32_t task_first, int32_t *__restrict__ order,
                               double *__restrict__ u, double *__restrict__ z) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= lmax * T) return;
    const uint32_t t = q % T, j = q / T;
    if ((int32_t)j >= task_len[t]) { order[q] = -1; u[q] = 0.0; z[q] = 0.0; return; }  // :2029-2034
    const uint32_t o = (uint32_t)task_off[t] + j;
    order[q] = task_off[t] + perm[o];
    if (ut) { u[q] = ut[o]; z[q] = zt[o]; return; }
    uint32_t w[4];
    philox4x32(j, iteration, 0x48594452u, 0u, seed, task_first + t, w);
    u[q] = ((double)(w[0] >> 5) * 67108864.0 + (double)(w[1] >> 6)) * (1.0 / 9007199254740992.0);
    const double u1 = ((double)w[2] + 0.5) * (1.0 / 4294967296.0);
    const double u2 = ((double)w[3] + 0.5) * (1.0 / 4294967296.0);
    z[q] = sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
}

// sum of beta^2 per group over the local markers (src/BayesRRm.cpp:2496-2499); block g = group g
__global__ void __launch_bounds__(1024) k_beta_sqnorm(const double *__restrict__ beta, const int32_t *__restrict__ grp,
                                                      uint32_t M, double *__restrict__ out) {
    __shared__ double red[32];
    const int g = blockIdx.x;
    double v = 0.0;
    for (uint32_t m = threadIdx.x; m < M; m += blockDim.x) {
        const double b = beta[m];
        if (grp[m] == g) v += b * b;
    }
    const double s = block_sum(v, red);
    if (threadIdx.x == 0) out[g] = s;
}

}  // namespace hb
