/*
 * hydra_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the per-marker Gibbs hot path of
 * medical-genomics-group/hydra (BayesRRm + genotype staging), written from the
 * behaviour of the reference sources; every function cites the reference
 * file:line it follows (paths relative to /root/reference).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (hydra_b200/)
 * never links, imports or executes anything in oracle/.
 *
 * PARITY STATUS
 *   - LUT tables: pinned against the reference's own src/dotp_lut.h and the
 *     reference's generator src/mk_lut.cpp (see oracle/build_ref.sh and
 *     tests/golden/lut_ref.npz, made by tests/golden/make_golden.py).
 *   - BED decode / sparse conversion / NA compaction: integer, restated 1:1.
 *   - Chain arithmetic: restated 1:1 from src/BayesRRm.cpp; the reference
 *     cannot be built here (needs MPI, Eigen, Boost -- none installed) and it
 *     ships no golden outputs, and its random streams come from Boost.Random
 *     which is absent: "parity unpinned" for the RNG streams and whole-chain
 *     outputs.  Replay is therefore defined against an explicit draw tape (see
 *     DESIGN.md, "Draw tape").
 *
 * Build: see oracle/Makefile.  Two variants of the same source:
 *   libhydra_oracle.so       -O2, strict IEEE, no OpenMP  (the checker)
 *   libhydra_oracle_fast.so  -Ofast -march=native -fopenmp (CPU baseline timing,
 *                            flags of the reference's src/Makefile_G:11-16)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint32_t uint;

#define HO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Lookup tables  (src/mk_lut.cpp:25-36 -> lut_a, :54-65 -> lut_b)            */
/* lut_a[byte*4+k] : genotype value of individual k in the byte               */
/*                   00->2, 01->0 (missing), 10->1, 11->0                     */
/* lut_b[byte*4+k] : non-missing mask   01->0, else 1                         */
/* ------------------------------------------------------------------------- */
static double g_lut_a[1024], g_lut_b[1024];
static int g_lut_ready = 0;

HO_API void ho_lut_build(double *a, double *b) {
    for (int i = 0; i < 256; i++) {
        for (int k = 0; k < 4; k++) {
            int code = (i >> (2 * k)) & 3;
            a[i * 4 + k] = (code == 0) ? 2.0 : (code == 2) ? 1.0 : 0.0;
            b[i * 4 + k] = (code == 1) ? 0.0 : 1.0;
        }
    }
}

static void lut_init(void) {
    if (!g_lut_ready) {
        ho_lut_build(g_lut_a, g_lut_b);
        g_lut_ready = 1;
    }
}

/* ------------------------------------------------------------------------- */
/* Genotype staging                                                           */
/* ------------------------------------------------------------------------- */

/* src/data.cpp:1166-1216  Data::sparse_data_get_sizes_from_raw
 * rawdata: NC columns of NB bytes; nind = numInds - NA individuals are decoded. */
HO_API void ho_sparse_get_sizes_from_raw(const uint8_t *rawdata, uint NC, uint NB, uint nind,
                                         size_t *N1, size_t *N2, size_t *NM) {
    size_t n1 = 0, n2 = 0, nm = 0;
    for (uint i = 0; i < NC; ++i) {
        const uint8_t *loc = rawdata + (size_t)i * NB;
        for (uint ii = 0; ii < nind; ++ii) {
            int code = (loc[ii >> 2] >> (2 * (ii & 3))) & 3; /* :1191 */
            if (code == 1) nm++;                              /* :1196 */
            else {
                int g = 2 - ((code & 1) + ((code >> 1) & 1)); /* :1199 */
                if (g == 1) n1++;
                else if (g == 2) n2++;
            }
        }
    }
    *N1 = n1; *N2 = n2; *NM = nm;
}

/* src/data.cpp:1224-1290  Data::sparse_data_fill_indices */
HO_API void ho_sparse_fill_indices(const uint8_t *rawdata, uint NC, uint NB, uint nind,
                                   size_t *N1S, size_t *N1L, uint *I1,
                                   size_t *N2S, size_t *N2L, uint *I2,
                                   size_t *NMS, size_t *NML, uint *IM) {
    size_t i1 = 0, i2 = 0, im = 0;
    size_t N1 = 0, N2 = 0, NM = 0;
    for (uint i = 0; i < NC; ++i) {
        const uint8_t *loc = rawdata + (size_t)i * NB;
        size_t n1 = 0, n2 = 0, nm = 0;
        for (uint ii = 0; ii < nind; ++ii) {
            int code = (loc[ii >> 2] >> (2 * (ii & 3))) & 3; /* :1248 */
            if (code == 1) {                                  /* :1253 missing */
                IM[im++] = ii; nm++;
            } else {
                int g = 2 - ((code & 1) + ((code >> 1) & 1)); /* :1256 */
                if (g == 1) { I1[i1++] = ii; n1++; }
                else if (g == 2) { I2[i2++] = ii; n2++; }
            }
        }
        N1S[i] = N1; N1L[i] = n1; N1 += n1;                   /* :1284-1286 */
        N2S[i] = N2; N2L[i] = n2; N2 += n2;
        NMS[i] = NM; NML[i] = nm; NM += nm;
    }
}

/* src/data.cpp:1112-1158  Data::sparse_data_correct_for_missing_phenotype
 * NAsInds ascending; entries equal to an NA index are dropped, others are
 * shifted down by #{na <= idx}; NL shrinks, NS does not move. */
HO_API void ho_sparse_correct_for_missing_phenotype(const size_t *NS, size_t *NL, uint *I, int M,
                                                    const uint8_t *usebed,
                                                    const uint *NAsInds, int numNAs) {
    for (int i = 0; i < M; ++i) {
        if (usebed && usebed[i]) continue;                    /* :1120 */
        const size_t beg = NS[i], len = NL[i];
        size_t k = 0; uint nas = 0;
        if (len > 0) {
            uint *tmp = (uint *)malloc(len * sizeof(uint));
            memcpy(tmp, I + beg, len * sizeof(uint));
            for (size_t iii = 0; iii < len; ++iii) {
                int isna = 0; uint allnas = 0;
                for (int ii = 0; ii < numNAs; ++ii) {
                    if (NAsInds[ii] > tmp[iii]) break;
                    if (NAsInds[ii] <= tmp[iii]) allnas += 1;
                    if (tmp[iii] == NAsInds[ii]) { isna = 1; nas += 1; break; }
                }
                if (isna) continue;
                I[beg + k] = tmp[iii] - allnas;
                k += 1;
            }
            free(tmp);
        }
        NL[i] -= nas;
    }
}

/* src/data.cpp:826-865 Data::get_bed_marker_from_sparse (same bit rule as the
 * mixed-representation loader :1002-1038): 0xFF then XOR 01 (ones -> 10),
 * XOR 11 (twos -> 00), XOR 10 (missing -> 01). nbytes bytes are initialised. */
HO_API void ho_bed_marker_from_sparse(uint8_t *bdat, size_t nbytes,
                                      const uint *I1, size_t L1,
                                      const uint *I2, size_t L2,
                                      const uint *IM, size_t LM) {
    memset(bdat, 0xFF, nbytes);
    for (size_t j = 0; j < L1; j++) bdat[I1[j] / 4] ^= (uint8_t)(1u << ((I1[j] % 4) * 2));
    for (size_t j = 0; j < L2; j++) bdat[I2[j] / 4] ^= (uint8_t)(3u << ((I2[j] % 4) * 2));
    for (size_t j = 0; j < LM; j++) bdat[IM[j] / 4] ^= (uint8_t)(2u << ((IM[j] % 4) * 2));
}

/* src/BayesRRm.cpp:1502-1508 (BayesRRm: mstd is the INVERSE sd) */
HO_API void ho_marker_stats_brr(int M, uint Ntot, const size_t *N1L, const size_t *N2L,
                                const size_t *NML, double *mave, double *mstd) {
    const double dN = (double)Ntot;
    for (int i = 0; i < M; ++i) {
        mave[i] = ((double)N1L[i] + 2.0 * (double)N2L[i]) / (dN - (double)NML[i]);
        double tmp1 = (double)N1L[i] * (1.0 - mave[i]) * (1.0 - mave[i]);
        double tmp2 = (double)N2L[i] * (2.0 - mave[i]) * (2.0 - mave[i]);
        double tmp0 = (double)(Ntot - N1L[i] - N2L[i] - NML[i]) * (0.0 - mave[i]) * (0.0 - mave[i]);
        mstd[i] = sqrt((double)(Ntot - 1) / (tmp0 + tmp1 + tmp2));
    }
}

/* src/BayesRRm.cpp:371-388 center_and_scale */
HO_API void ho_center_and_scale(double *vec, int N) {
    double mean = 0.0;
    for (int i = 0; i < N; ++i) mean += vec[i];
    mean /= N;
    for (int i = 0; i < N; ++i) vec[i] -= mean;
    double sqn = 0.0;
    for (int i = 0; i < N; ++i) sqn += vec[i] * vec[i];
    sqn = sqrt((double)(N - 1) / sqn);
    for (int i = 0; i < N; ++i) vec[i] *= sqn;
}

/* ------------------------------------------------------------------------- */
/* Numeric kernels                                                            */
/* ------------------------------------------------------------------------- */

/* src/BayesRRm.cpp:129-144 */
static double sum_vector_elements_f64(const double *vec, int N) {
    double sum = 0.0;
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : sum)
#endif
    for (int i = 0; i < N; i++) sum += vec[i];
    return sum;
}

/* src/BayesRRm.cpp:284-313 */
static double partial_sparse_dotprod(const double *vec, const uint *IX, size_t NXS, size_t NXL, double fac) {
    double dp = 0.0;
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : dp)
#endif
    for (size_t i = NXS; i < NXS + NXL; i++) dp += vec[IX[i]];
    dp *= fac;
    return dp;
}

/* src/BayesRRm.cpp:316-342  BayesRRm::sparse_dotprod */
HO_API double ho_sparse_dotprod(const double *vin1,
                                const uint *I1, size_t N1S, size_t N1L,
                                const uint *I2, size_t N2S, size_t N2L,
                                const uint *IM, size_t NMS, size_t NML,
                                double mu, double sig_inv, int N) {
    double dp = 0.0;
    dp += partial_sparse_dotprod(vin1, I1, N1S, N1L, 1.0);
    dp += partial_sparse_dotprod(vin1, I2, N2S, N2L, 2.0);
    double syt = sum_vector_elements_f64(vin1, N);            /* :331 full rescan */
    double dsyt = partial_sparse_dotprod(vin1, IM, NMS, NML, 1.0);
    syt -= dsyt;
    dp -= (mu * syt);
    dp *= sig_inv;
    return dp;
}

/* src/BayesRRm.cpp:1766-1809  LUT dot product on raw BED bytes (dp_version 1).
 * The reference accumulates in a 4-lane AVX2 register (lane k = individual
 * 4*ii+k) and sums the lanes at the end (:1794-1795); restated with 4 scalars. */
HO_API double ho_lut_dotprod(const uint8_t *rawdata, const double *epsilon, int Ntot,
                             double mave, double mstd) {
    lut_init();
    const int fullb = Ntot / 4;
    double v1[4] = {0, 0, 0, 0}, v2[4] = {0, 0, 0, 0};
    for (int ii = 0; ii < fullb; ++ii) {
        const double *c1 = &g_lut_a[rawdata[ii] * 4];
        const double *c2 = &g_lut_b[rawdata[ii] * 4];
        for (int k = 0; k < 4; k++) {
            double p = c2[k] * epsilon[ii * 4 + k];           /* :1785 */
            v2[k] += p;                                       /* :1787 */
            v1[k] += p * c1[k];                               /* :1789-1790 */
        }
    }
    double s1 = v1[0] + v1[1] + v1[2] + v1[3];
    double s2 = v2[0] + v2[1] + v2[2] + v2[3];
    if (Ntot % 4 != 0) {                                      /* :1798-1807 */
        int ii = fullb;
        for (int iii = 0; iii < Ntot - fullb * 4; iii++) {
            int idx = rawdata[ii] * 4 + iii;
            double c1 = g_lut_a[idx], c2 = g_lut_b[idx];
            s1 += c1 * (c2 * epsilon[ii * 4 + iii]);
            s2 += (c2 * epsilon[ii * 4 + iii]);
        }
    }
    return mstd * (s1 - mave * s2);                           /* :1809 */
}

/* src/BayesRRm.cpp:73-91 */
static void set_vector_f64(double *vec, double val, int N) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (int i = 0; i < N; i++) vec[i] = val;
}

/* src/BayesRRm.cpp:212-227 */
static void sparse_set(double *vec, double val, const uint *IX, size_t NXS, size_t NXL) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (size_t i = NXS; i < NXS + NXL; ++i) vec[IX[i]] = val;
}

/* src/BayesRRm.cpp:250-281  BayesRRm::sparse_scaadd */
HO_API void ho_sparse_scaadd(double *vout, double dMULT,
                             const uint *I1, size_t N1S, size_t N1L,
                             const uint *I2, size_t N2S, size_t N2L,
                             const uint *IM, size_t NMS, size_t NML,
                             double mu, double sig_inv, int N) {
    if (dMULT == 0.0) {
        set_vector_f64(vout, 0.0, N);
    } else {
        double aux = mu * sig_inv * dMULT;
        set_vector_f64(vout, -aux, N);
        sparse_set(vout, 0.0, IM, NMS, NML);
        aux = dMULT * (1.0 - mu) * sig_inv;
        sparse_set(vout, aux, I1, N1S, N1L);
        aux = dMULT * (2.0 - mu) * sig_inv;
        sparse_set(vout, aux, I2, N2S, N2L);
    }
}

/* src/BayesRRm.cpp:1976-2010  deltaEps from raw BED bytes */
HO_API void ho_lut_scaadd(double *deltaEps, const uint8_t *rawdata, double deltaBeta,
                          double mave, double mstd, int Ntot) {
    lut_init();
    const double sigdb = mstd * deltaBeta;                    /* :1982 */
    const int fullb = Ntot / 4;
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (int ii = 0; ii < fullb; ++ii) {
        for (int iii = 0; iii < 4; iii++) {
            int idx = rawdata[ii] * 4 + iii;
            deltaEps[ii * 4 + iii] = (g_lut_a[idx] - mave) * g_lut_b[idx] * sigdb; /* :1997 */
        }
    }
    if (Ntot % 4 != 0) {
        int ii = fullb;
        for (int iii = 0; iii < Ntot - fullb * 4; iii++) {
            int idx = rawdata[ii] * 4 + iii;
            deltaEps[ii * 4 + iii] = (g_lut_a[idx] - mave) * g_lut_b[idx] * sigdb;
        }
    }
}

/* src/BayesRRm.cpp:163-182 */
static void sum_vectors_f64_inplace(double *out, const double *in1, int N) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (int i = 0; i < N; i++) out[i] += in1[i];
}
/* src/BayesRRm.cpp:147-160 */
static void sum_vectors_f64(double *out, const double *in1, const double *in2, int N) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (int i = 0; i < N; i++) out[i] = in1[i] + in2[i];
}
/* src/BayesRRm.cpp:94-105 */
static void copy_vector_f64(double *dest, const double *source, int N) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (int i = 0; i < N; i++) dest[i] = source[i];
}

/* src/BayesRRm.cpp:396-413  mpi_define_blocks_of_markers */
HO_API void ho_define_blocks_of_markers(int Mtot, int *MrankS, int *MrankL, uint nblocks) {
    const uint modu = Mtot % nblocks;
    uint start = 0;
    for (uint i = 0; i < nblocks; ++i) {
        MrankL[i] = (int)(Mtot / nblocks);
        if (modu != 0 && i < modu) MrankL[i] += 1;
        MrankS[i] = start;
        start += MrankL[i];
    }
}

/* ------------------------------------------------------------------------- */
/* RNG spec v1 (our own; the reference's Boost.Random streams are unpinned)   */
/* ------------------------------------------------------------------------- */
typedef struct { uint32_t mt[624]; int mti; } ho_mt;

HO_API void ho_mt_seed(ho_mt *s, uint32_t seed) { /* MT19937 init_genrand */
    s->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->mti = 624;
}
HO_API uint32_t ho_mt_u32(ho_mt *s) {
    if (s->mti >= 624) {
        uint32_t *mt = s->mt;
        for (int k = 0; k < 624; k++) {
            uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s->mti = 0;
    }
    uint32_t y = s->mt[s->mti++];
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
    return y;
}
HO_API double ho_mt_res53(ho_mt *s) { /* [0,1) with 53-bit resolution */
    uint32_t a = ho_mt_u32(s) >> 5, b = ho_mt_u32(s) >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}
HO_API double ho_mt_normal(ho_mt *s) { /* Box-Muller, no caching */
    double u1 = ho_mt_res53(s), u2 = ho_mt_res53(s);
    return sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586476925 * u2);
}
HO_API double ho_mt_gamma(ho_mt *s, double a) { /* Marsaglia-Tsang, scale 1 */
    if (a < 1.0) {
        double g = ho_mt_gamma(s, a + 1.0);
        double u = 1.0 - ho_mt_res53(s);
        return g * pow(u, 1.0 / a);
    }
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        double x = ho_mt_normal(s);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        double u = 1.0 - ho_mt_res53(s);
        if (u < 1.0 - 0.0331 * (x * x) * (x * x)) return d * v;
        if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return d * v;
    }
}
/* src/distributions_boost.cpp:92-94,112-114: inv_scaled_chisq(dof,scale) =
 * inv_gamma(0.5*dof, 0.5*dof*scale) = 1/rgamma(shape, 1/scale') */
static double ho_inv_scaled_chisq(ho_mt *s, double dof, double scale) {
    double shape = 0.5 * dof, sc = 0.5 * dof * scale;
    return 1.0 / (ho_mt_gamma(s, shape) * (1.0 / sc));
}
HO_API void ho_shuffle(ho_mt *s, int *a, int n) { /* Fisher-Yates, spec v1 */
    for (int i = n - 1; i >= 1; i--) {
        int k = (int)(ho_mt_res53(s) * (double)(i + 1));
        int t = a[i]; a[i] = a[k]; a[k] = t;
    }
}

/* Philox4x32-10 (Salmon et al. 2011) */
HO_API void ho_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
#define HO_TAG_MARKER 0x48594452u /* 'HYDR' */

/* Per-marker positional draws of spec v1: u in [0,1), z standard normal. */
HO_API void ho_marker_draws(uint32_t seed, uint32_t task, uint32_t iteration, uint32_t j,
                            double *u, double *z) {
    uint32_t ctr[4] = {j, iteration, HO_TAG_MARKER, 0}, key[2] = {seed, task}, w[4];
    ho_philox4x32(ctr, key, w);
    *u = ((double)(w[0] >> 5) * 67108864.0 + (double)(w[1] >> 6)) * (1.0 / 9007199254740992.0);
    double u1 = ((double)w[2] + 0.5) * (1.0 / 4294967296.0);
    double u2 = ((double)w[3] + 0.5) * (1.0 / 4294967296.0);
    *z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
}

/* BayesFH per-marker standard Gamma(a, 1) variates of spec v1 (the reference draws them from the rank's Boost stream inside
 * inv_gamma_rate_rng, src/BayesRRm.cpp:1729 and :1952; src/distributions_boost.cpp:57-61, 97-103): Marsaglia-Tsang, one Philox
 * block per attempt, keyed by (seed, global marker), counter (attempt, iteration, tag); `which` 0 = nu_var, 1 = lambda_var. */
#define HO_TAG_FHNU 0x46484E55u /* 'FHNU' */
#define HO_TAG_FHLA 0x46484C41u /* 'FHLA' */
HO_API double ho_fh_gamma(uint32_t seed, uint32_t marker, uint32_t iteration, int which, double a) {
    const uint32_t tag = which ? HO_TAG_FHLA : HO_TAG_FHNU;
    const double a1 = (a < 1.0) ? a + 1.0 : a;
    const double d = a1 - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double g = d;
    for (uint32_t attempt = 0; attempt < 4096u; attempt++) {
        uint32_t ctr[4] = {attempt, iteration, tag, 0}, key[2] = {seed, marker}, w[4];
        ho_philox4x32(ctr, key, w);
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
        uint32_t ctr[4] = {0xFFFFFFFFu, iteration, tag, 0}, key[2] = {seed, marker}, w[4];
        ho_philox4x32(ctr, key, w);
        g *= pow(((double)w[0] + 0.5) * (1.0 / 4294967296.0), 1.0 / a);
    }
    return g;
}
/* src/distributions_boost.cpp:97-103: inv_gamma_rate_rng(shape, rate) = 1 / rgamma(shape, 1/rate), g the standard variate */
static double ho_inv_gamma_rate_from(double g, double rate) { return 1.0 / (g * (1.0 / rate)); }

/* ------------------------------------------------------------------------- */
/* Synthetic genotypes; counter-based, integer thresholds     */
/* cell (i,j): word k=i%4 of philox(ctr=(i/4, j, attempt, 'GENO'), key=seed)    */
/*   w <  t[0]            -> missing (01)                                       */
/*   w <  t[1]            -> genotype 2 (00)                                    */
/*   w <  t[2]            -> genotype 1 (10)                                    */
/*   else                 -> genotype 0 (11)                                    */
/* Padding bits of the last byte are 00 (PLINK pads with 00).                   */
/* ------------------------------------------------------------------------- */
#define HO_TAG_GENO 0x47454E4Fu
HO_API void ho_synth_thresholds(double p, double pmiss, uint32_t t[3]) {
    double tm = floor(pmiss * 4294967296.0);
    double rest = 4294967296.0 - tm;
    double t2 = tm + floor(p * p * rest);
    double t1 = t2 + floor(2.0 * p * (1.0 - p) * rest);
    t[0] = (uint32_t)tm; t[1] = (uint32_t)t2; t[2] = (uint32_t)t1;
}
HO_API void ho_synth_bed_marker(uint32_t seed, uint32_t j, uint32_t attempt, const uint32_t t[3],
                                uint32_t N, uint8_t *out) {
    uint32_t nb = (N + 3) / 4;
    for (uint32_t q = 0; q < nb; q++) {
        uint32_t ctr[4] = {q, j, attempt, HO_TAG_GENO}, key[2] = {seed, 0}, w[4];
        ho_philox4x32(ctr, key, w);
        uint8_t byte = 0;
        for (int k = 0; k < 4; k++) {
            uint32_t i = q * 4 + k;
            uint32_t code;
            if (i >= N) code = 0;
            else if (w[k] < t[0]) code = 1;
            else if (w[k] < t[1]) code = 0;
            else if (w[k] < t[2]) code = 2;
            else code = 3;
            byte |= (uint8_t)(code << (2 * k));
        }
        out[q] = byte;
    }
}

/* ------------------------------------------------------------------------- */
/* BayesRRm chain (src/BayesRRm.cpp:1644-2731), T tasks simulated in-process   */
/* ------------------------------------------------------------------------- */
typedef struct {
    /* dimensions */
    int32_t N, Mtot, T, K, G, sync_rate, n_iter, iter0;
    /* genotypes, reference representation, GLOBAL marker indexing 0..Mtot-1.
     * A marker with usebed[m]!=0 is read from bed + m*snpLenByt, else from lists. */
    const uint *I1; const size_t *N1S; const size_t *N1L;
    const uint *I2; const size_t *N2S; const size_t *N2L;
    const uint *IM; const size_t *NMS; const size_t *NML;
    const uint8_t *usebed; const uint8_t *bed; size_t snpLenByt;
    /* model */
    const double *y_raw;     /* N, uncentred phenotype */
    const int32_t *groups;   /* Mtot */
    const double *mS;        /* G*K row-major, column 0 = 0.0 */
    /* tape (inputs) */
    const double *tape_zmu;  /* n_iter*T, standard normals */
    const int32_t *tape_perm;/* n_iter*Mtot, task blocks concatenated, LOCAL indices */
    const double *tape_u;    /* n_iter*Mtot, indexed [it][MrankS[r]+j] */
    const double *tape_z;    /* n_iter*Mtot */
    const double *tape_sigmaG0; /* G: initial sigmaG (src/BayesRRm.cpp:1233) */
    const double *tape_sigmaG;  /* n_iter*G or NULL -> draw with mt_hyper */
    const double *tape_sigmaE;  /* n_iter   or NULL */
    const double *tape_pi;      /* n_iter*G*K or NULL */
    uint32_t hyper_seed;        /* used when the three above are NULL */
    /* outputs (may be NULL) */
    double *out_beta;   /* n_iter*Mtot */
    int32_t *out_comp;  /* n_iter*Mtot */
    double *out_acum;   /* n_iter*Mtot */
    double *out_mu;     /* n_iter*T */
    double *out_eps;    /* n_iter*T*N  (task-wise epsilon at end of iteration) */
    double *out_sigmaG; /* n_iter*G */
    double *out_sigmaE; /* n_iter */
    double *out_pi;     /* n_iter*G*K */
    double *out_bsq;    /* n_iter*G   (summed over tasks) */
    int32_t *out_cass;  /* n_iter*G*K (summed over tasks) */
    double *out_esqn;   /* n_iter */
    double *out_epssum; /* n_iter*T */
    int64_t *out_nsync; /* n_iter: number of epsilon synchronisations */
    double *out_loop_seconds; /* n_iter: wall time of the marker loop only */
    /* fixed effects (--covariates, src/BayesRRm.cpp:2648-2681); F == 0: none */
    const double *X;          /* N x F row-major, used as read (src/data.cpp:1615-1672: no centring or scaling) */
    int32_t F;
    const int32_t *tape_xI;   /* n_iter*F: order of the covariates per iteration, or NULL -> Fisher-Yates of the running order with mt_hyper */
    const double *tape_zcov;  /* n_iter*F: standard normals, or NULL -> mt_hyper */
    double *out_gamma;        /* n_iter*F */
    /* per-group priors: --groupPriorsFile (v0G, s02G per group, :2545-2548) and --dPriorsFile (Dirichlet parameters, :2551-2554) */
    const double *group_priors; /* G*2 or NULL -> v0G = s02G = 0.0001 */
    const double *dirichlet;    /* G*K or NULL -> 1.0 (:1184-1185) */
    /* bayesFHMPI (horseshoe-type local scales; src/BayesRRm.cpp:1125-1163, 1727-1731, 1747-1748, 1869-1872, 1942-1952, 2503-2510,
     * 2557-2565); fh == 0: off */
    int32_t fh;
    double fh_v0L, fh_v0t, fh_v0c, fh_s02c, fh_tau0;   /* src/options.hpp:91-96 */
    const double *fh_state0;    /* 2+G: hypTau, tau, c_slab[G] after :1147-1154, or NULL -> drawn with mt_hyper in that order */
    const double *tape_gnu;     /* n_iter*Mtot standard Gamma(0.5+0.5*v0L) variates of the nu_var draws (:1729) by GLOBAL marker, NULL -> ho_fh_gamma */
    const double *tape_glam;    /* n_iter*Mtot, the lambda_var draws (:1952) */
    const double *tape_fh_hyper;/* n_iter*G*3: hypTau, tau, c_slab[g] as drawn in group g's pass (:2559-2562), NULL -> mt_hyper */
    uint32_t fh_seed;           /* seed of ho_fh_gamma */
    double *out_fh;             /* n_iter*(3+G): hypTau, tau, scaledBSQN, c_slab[G] */
    double *out_lambda;         /* n_iter*Mtot */
    double *out_nu;             /* n_iter*Mtot */
} ho_brr_args;

static double now_s(void) {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
#endif
}

HO_API int ho_brr_chain(const ho_brr_args *a) {
    const int N = a->N, Mtot = a->Mtot, T = a->T, K = a->K, G = a->G;
    const int km1 = K - 1;
    const double dN = (double)N, dNm1 = (double)(N - 1);
    const double v0E = 0.0001, s02E = 0.0001; /* src/BayesRRm.h:30-33 */
    double v0G = 0.0001, s02G = 0.0001;       /* overwritten group by group where a priors file was given (:2545-2548) */

    int *MrankS = (int *)malloc(sizeof(int) * T), *MrankL = (int *)malloc(sizeof(int) * T);
    ho_define_blocks_of_markers(Mtot, MrankS, MrankL, (uint)T);
    int lmax = 0;
    for (int r = 0; r < T; r++) if (MrankL[r] > lmax) lmax = MrankL[r];

    /* src/BayesRRm.cpp:1097-1110 */
    double *cVa = (double *)calloc((size_t)G * K, sizeof(double));
    double *cVaI = (double *)calloc((size_t)G * K, sizeof(double));
    double *estPi = (double *)calloc((size_t)G * K, sizeof(double));
    for (int g = 0; g < G; g++) {
        double s = 0.0;
        for (int k = 1; k < K; k++) { cVa[g * K + k] = a->mS[g * K + k]; cVaI[g * K + k] = 1.0 / cVa[g * K + k]; s += cVa[g * K + k]; }
        estPi[g * K + 0] = 0.5;
        for (int k = 1; k < K; k++) estPi[g * K + k] = 0.5 * cVa[g * K + k] / s;
    }
    int *MtotGrp = (int *)calloc(G, sizeof(int));
    for (int m = 0; m < Mtot; m++) MtotGrp[a->groups[m]] += 1;  /* :1189-1193 */

    double *sigmaG = (double *)malloc(sizeof(double) * G);
    for (int g = 0; g < G; g++) sigmaG[g] = a->tape_sigmaG0[g];
    for (int g = 0; g < G; g++) if (MtotGrp[g] == 0) sigmaG[g] = 0.0; /* :1239-1240 */

    ho_mt mt_hyper; ho_mt_seed(&mt_hyper, a->hyper_seed);
    /* :1125-1163 FH initialisation */
    const int fh = a->fh;
    const double v0L = a->fh_v0L, v0t = a->fh_v0t, v0c = a->fh_v0c, s02c = a->fh_s02c, tau0 = a->fh_tau0;
    double hypTau = 0.0, tau = 0.0, scaledBSQN = 0.0;
    double *c_slab = (double *)calloc((size_t)G, sizeof(double));
    double *lambda_var = (double *)calloc((size_t)(Mtot > 0 ? Mtot : 1), sizeof(double));
    double *nu_var = (double *)calloc((size_t)(Mtot > 0 ? Mtot : 1), sizeof(double));
    double *gnu = (double *)calloc((size_t)(Mtot > 0 ? Mtot : 1), sizeof(double));
    double *glam = (double *)calloc((size_t)(Mtot > 0 ? Mtot : 1), sizeof(double));
    if (fh) {
        if (a->fh_state0) {
            hypTau = a->fh_state0[0]; tau = a->fh_state0[1];
            for (int g = 0; g < G; g++) c_slab[g] = a->fh_state0[2 + g];
        } else {
            hypTau = ho_inv_gamma_rate_from(ho_mt_gamma(&mt_hyper, 0.5), 1.0 / (tau0 * tau0));        /* :1147 */
            tau = ho_inv_gamma_rate_from(ho_mt_gamma(&mt_hyper, 0.5 * v0t), v0t / hypTau);            /* :1150 */
            for (int g = 0; g < G; g++) c_slab[g] = ho_inv_scaled_chisq(&mt_hyper, v0c, s02c);        /* :1153-1154 */
        }
        double cs = 0.0;
        for (int g = 0; g < G; g++) cs += c_slab[g];
        for (int m = 0; m < Mtot; m++) lambda_var[m] = cs / (double)Mtot;                             /* :1161 */
    }
    int *xI = (int *)malloc(sizeof(int) * (size_t)(a->F > 0 ? a->F : 1));
    double *gam = (double *)calloc((size_t)(a->F > 0 ? a->F : 1), sizeof(double));   /* gamma.setZero() :1092 */
    for (int i = 0; i < a->F; i++) xI[i] = i;                                         /* :1113-1117 */

    /* mave/mstd :1502-1508 (global arrays) */
    double *mave = (double *)malloc(sizeof(double) * Mtot), *mstd = (double *)malloc(sizeof(double) * Mtot);
    ho_marker_stats_brr(Mtot, (uint)N, a->N1L, a->N2L, a->NML, mave, mstd);

    /* per-task state :1528-1537 */
    double *y = (double *)malloc(sizeof(double) * N);
    for (int i = 0; i < N; i++) y[i] = a->y_raw[i];
    ho_center_and_scale(y, N);                                  /* :1565-1566 */
    double **eps = (double **)malloc(sizeof(double *) * T), **tmpEps = (double **)malloc(sizeof(double *) * T);
    double **dEpsSum = (double **)malloc(sizeof(double *) * T);
    double **deltaEpsT = (double **)malloc(sizeof(double *) * T);
    double *deltaSum = (double *)malloc(sizeof(double) * N);
    for (int r = 0; r < T; r++) {
        deltaEpsT[r] = (double *)malloc(sizeof(double) * N);
        eps[r] = (double *)malloc(sizeof(double) * N);
        tmpEps[r] = (double *)malloc(sizeof(double) * N);
        dEpsSum[r] = (double *)calloc(N, sizeof(double));      /* :1544 */
        for (int i = 0; i < N; i++) eps[r][i] = y[i];          /* :1575 */
    }
    double sigmaE = 0.0;
    for (int i = 0; i < N; i++) sigmaE += eps[0][i] * eps[0][i];
    sigmaE = sigmaE / dN * 0.5;                                 /* :1576-1579 */

    double *Beta = (double *)calloc(Mtot, sizeof(double));
    int *components = (int *)calloc(Mtot, sizeof(int));
    double *Acum = (double *)calloc(Mtot, sizeof(double));
    uint8_t *adaV = (uint8_t *)malloc(Mtot);
    for (int m = 0; m < Mtot; m++) adaV[m] = (sigmaG[a->groups[m]] == 0.0) ? 0 : 1; /* :1592-1597 */
    double *mu = (double *)calloc(T, sizeof(double));
    int *cass = (int *)malloc(sizeof(int) * (size_t)T * G * K);
    int *sum_cass = (int *)malloc(sizeof(int) * (size_t)G * K);
    double *bsq = (double *)malloc(sizeof(double) * G);
    int *m0 = (int *)malloc(sizeof(int) * G);
    double *tsad = (double *)malloc(sizeof(double) * T); /* task_sum_abs_deltabeta */
    if (K > 64) return -1;

    for (int itx = 0; itx < a->n_iter; itx++) {
        const int32_t *perm = a->tape_perm + (size_t)itx * Mtot;
        const double *tu = a->tape_u + (size_t)itx * Mtot;
        const double *tz = a->tape_z + (size_t)itx * Mtot;

        for (int r = 0; r < T; r++) {
            double *e = eps[r];
            for (int i = 0; i < N; ++i) e[i] += mu[r];          /* :1675 */
            double epssum = 0.0;
            for (int i = 0; i < N; ++i) epssum += e[i];         /* :1677-1678 */
            if (a->out_epssum) a->out_epssum[(size_t)itx * T + r] = epssum;
            /* :1682 norm_rng(mean, var) = mean + sqrt(var) * z */
            mu[r] = epssum / dN + sqrt(sigmaE / dN) * a->tape_zmu[(size_t)itx * T + r];
            for (int i = 0; i < N; ++i) e[i] -= mu[r];          /* :1686 */
            for (int i = 0; i < N; ++i) tmpEps[r][i] = e[i];    /* :1700 */
            tsad[r] = 0.0;
        }
        memset(cass, 0, sizeof(int) * (size_t)T * G * K);       /* :1697 */
        if (fh) {   /* the standard gamma variates behind this iteration's nu_var / lambda_var draws */
            const double sh = 0.5 + 0.5 * v0L;
            for (int m = 0; m < Mtot; m++) {
                gnu[m] = a->tape_gnu ? a->tape_gnu[(size_t)itx * Mtot + m] : ho_fh_gamma(a->fh_seed, (uint32_t)m, (uint32_t)(a->iter0 + itx), 0, sh);
                glam[m] = a->tape_glam ? a->tape_glam[(size_t)itx * Mtot + m] : ho_fh_gamma(a->fh_seed, (uint32_t)m, (uint32_t)(a->iter0 + itx), 1, sh);
            }
        }
        int sinceLastSync = 0;
        int64_t nsync = 0;
        double t_loop = now_s();

        for (int j = 0; j < lmax; j++) {                        /* :1709 */
            sinceLastSync += 1;
            /* tasks are independent inside one step (each MPI rank works on its own
             * copy); the timing build runs them as OpenMP threads = "T tasks x 1 thread" */
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (T > 1)
#endif
            for (int r = 0; r < T; r++) {
                double muk[64], denom[64], logL[64];
                double *deltaEps = deltaEpsT[r];
                double deltaBeta = 0.0;
                if (j < MrankL[r]) {
                    const int marker = perm[MrankS[r] + j];     /* local index :1717 */
                    const int gm = MrankS[r] + marker;          /* global marker */
                    const int grp = a->groups[gm];
                    double beta = Beta[gm];
                    const double sigE_G = sigmaE / sigmaG[grp]; /* :1721-1723 */
                    const double sigG_E = sigmaG[grp] / sigmaE;
                    const double i_2sigE = 1.0 / (2.0 * sigmaE);
                    double lambda_tilde = 0.0;
                    if (fh) {                                   /* :1727-1731 */
                        nu_var[gm] = ho_inv_gamma_rate_from(gnu[gm], v0L / lambda_var[gm] + 1);
                        lambda_tilde = tau * c_slab[grp] / (tau + c_slab[grp] * lambda_var[gm]);
                    }
                    if (adaV[gm]) {
                        for (int i = 1; i <= km1; ++i) {
                            if (fh) denom[i - 1] = dNm1 + sigmaE / lambda_tilde;                 /* :1747-1748 */
                            else denom[i - 1] = dNm1 + sigE_G * cVaI[grp * K + i];               /* :1750 */
                        }
                        double num;
                        if (a->usebed && a->usebed[gm]) {
                            num = ho_lut_dotprod(a->bed + (size_t)gm * a->snpLenByt, eps[r], N, mave[gm], mstd[gm]);
                        } else {
                            num = ho_sparse_dotprod(eps[r], a->I1, a->N1S[gm], a->N1L[gm], a->I2, a->N2S[gm], a->N2L[gm],
                                                    a->IM, a->NMS[gm], a->NML[gm], mave[gm], mstd[gm], N);
                        }
                        num += beta * (double)(N - 1);          /* :1855 */
                        for (int i = 1; i <= km1; i++) muk[i] = num / denom[i - 1]; /* :1859 */
                        muk[0] = 0.0;
                        for (int i = 0; i < K; i++) logL[i] = log(estPi[grp * K + i]); /* :1863-1864 */
                        for (int i = 1; i < 1 + km1; i++) {
                            if (fh) logL[i] = logL[i] - 0.5 * log((lambda_tilde / sigmaE) * dNm1 + 1.0) + muk[i] * num * i_2sigE;  /* :1869-1872 */
                            else logL[i] = logL[i] - 0.5 * log(sigG_E * dNm1 * cVa[grp * K + i] + 1.0) + muk[i] * num * i_2sigE;   /* :1874-1876 */
                        }
                        double prob = tu[MrankS[r] + j];        /* :1880 */
                        double acum;
                        int any = 0;
                        for (int i = 1; i <= km1; i++) if (fabs(logL[i] - logL[0]) > 700.0) any = 1; /* :1884 */
                        if (any) acum = 0.0;
                        else { double s = 0.0; for (int i = 0; i < K; i++) s += exp(logL[i] - logL[0]); acum = 1.0 / s; }
                        Acum[gm] = acum;                        /* :1892 */
                        for (int k = 0; k < K; k++) {           /* :1894-1921 */
                            if (prob <= acum || k == km1) {
                                if (k == 0) Beta[gm] = 0.0;
                                else Beta[gm] = muk[k] + sqrt(sigmaE / denom[k - 1]) * tz[MrankS[r] + j]; /* :1901 */
                                cass[((size_t)r * G + grp) * K + k] += 1;
                                components[gm] = k;
                                break;
                            } else {
                                int any2 = 0;
                                for (int i = k + 1; i < K; i++) if (fabs(logL[i] - logL[k + 1]) > 700.0) any2 = 1; /* :1915 */
                                if (!any2) { double s = 0.0; for (int i = 0; i < K; i++) s += exp(logL[i] - logL[k + 1]); acum += 1.0 / s; }
                            }
                        }
                    } else {
                        Beta[gm] = 0.0; Acum[gm] = 1.0;         /* :1924-1925 */
                    }
                    double betaOld = beta;
                    beta = Beta[gm];
                    deltaBeta = betaOld - beta;                 /* :1933 */
                    if (fh)                                     /* :1942-1952 */
                        lambda_var[gm] = ho_inv_gamma_rate_from(glam[gm], 0.5 * beta * beta / tau + v0L / nu_var[gm]);
                    if (deltaBeta != 0.0) {                     /* :1965 */
                        if (a->usebed && a->usebed[gm])
                            ho_lut_scaadd(deltaEps, a->bed + (size_t)gm * a->snpLenByt, deltaBeta, mave[gm], mstd[gm], N);
                        else
                            ho_sparse_scaadd(deltaEps, deltaBeta, a->I1, a->N1S[gm], a->N1L[gm], a->I2, a->N2S[gm], a->N2L[gm],
                                             a->IM, a->NMS[gm], a->NML[gm], mave[gm], mstd[gm], N);
                        sum_vectors_f64_inplace(dEpsSum[r], deltaEps, N); /* :2022 */
                    }
                }
                tsad[r] += fabs(deltaBeta);                     /* :2036 */
            }
            const int check = (sinceLastSync >= a->sync_rate || j == lmax - 1);
            double cumSum = 0.0;                                /* :2044-2062 */
            for (int r = 0; r < T; r++) cumSum += tsad[r];
            if (check && cumSum != 0.0) {                       /* :2069 */
                if (T > 1) {
                    /* :2456 MPI_Allreduce(SUM): rank order 0..T-1 (MPI leaves it unspecified) */
                    for (int i = 0; i < N; i++) deltaSum[i] = dEpsSum[0][i];
                    for (int r = 1; r < T; r++) sum_vectors_f64_inplace(deltaSum, dEpsSum[r], N);
                    for (int r = 0; r < T; r++) sum_vectors_f64(eps[r], tmpEps[r], deltaSum, N); /* :2460 */
                } else {
                    sum_vectors_f64(eps[0], tmpEps[0], dEpsSum[0], N); /* :2471 */
                }
                for (int r = 0; r < T; r++) {
                    copy_vector_f64(tmpEps[r], eps[r], N);      /* :2478 */
                    set_vector_f64(dEpsSum[r], 0.0, N);         /* :2481 */
                    tsad[r] = 0.0;                              /* :2485 */
                }
                sinceLastSync = 0;                              /* :2487 */
                nsync++;
            }
        }
        t_loop = now_s() - t_loop;
        if (a->out_loop_seconds) a->out_loop_seconds[itx] = t_loop;
        if (a->out_nsync) a->out_nsync[itx] = nsync;

        /* :2496-2521 */
        for (int g = 0; g < G; g++) bsq[g] = 0.0;
        for (int r = 0; r < T; r++) {
            /* each task sums its own block, then Allreduce in rank order */
            double *loc = (double *)calloc(G, sizeof(double));
            for (int i = 0; i < MrankL[r]; i++) { int gm = MrankS[r] + i; loc[a->groups[gm]] += Beta[gm] * Beta[gm]; }
            for (int g = 0; g < G; g++) bsq[g] = (r == 0) ? loc[g] : bsq[g] + loc[g];
            free(loc);
        }
        for (int x = 0; x < G * K; x++) { int s = 0; for (int r = 0; r < T; r++) s += cass[(size_t)r * G * K + x]; sum_cass[x] = s; }
        /* :2503-2510 scaled sum of squares. The reference sums over the rank's own markers and does not reduce it (nor does it
         * broadcast tau / c_slab: its ranks' FH parameters diverge); here, as for the other hyper-parameters, ONE value: the sum
         * over all markers, which is what a one-rank run of the reference computes. */
        scaledBSQN = 0.0;
        if (fh) for (int i = 0; i < Mtot; i++) scaledBSQN += Beta[i] * Beta[i] / lambda_var[i];

        /* :2525-2578 */
        for (int g = 0; g < G; g++) {
            m0[g] = 0;
            if (MtotGrp[g] == 0) continue;
            m0[g] = MtotGrp[g] - sum_cass[g * K + 0];
            int rowsum = 0; for (int k = 0; k < K; k++) rowsum += sum_cass[g * K + k];
            if (m0[g] == 0 || rowsum == 0) {
                for (int m = 0; m < Mtot; m++) if (a->groups[m] == g) adaV[m] = 0;
                sigmaG[g] = 0.0;
                continue;
            }
            if (a->group_priors) { v0G = a->group_priors[g * 2 + 0]; s02G = a->group_priors[g * 2 + 1]; }   /* :2545-2548 */
            if (fh) {                                           /* :2557-2565 */
                if (a->tape_fh_hyper) {
                    const double *h = a->tape_fh_hyper + ((size_t)itx * G + g) * 3;
                    hypTau = h[0]; tau = h[1]; c_slab[g] = h[2];
                } else {
                    hypTau = ho_inv_gamma_rate_from(ho_mt_gamma(&mt_hyper, 0.5 + 0.5 * v0t), 1.0 / (tau0 * tau0) + 1.0 / tau);
                    tau = ho_inv_gamma_rate_from(ho_mt_gamma(&mt_hyper, 0.5 * (m0[g] + v0t)), v0t / hypTau + (0.5 * scaledBSQN));
                    c_slab[g] = ho_inv_scaled_chisq(&mt_hyper, v0c + (double)m0[g], (bsq[g] * (double)m0[g] + v0c * s02c) / (v0c + (double)m0[g]));
                }
                sigmaG[g] = bsq[g];                             /* :2565 */
            }
            if (a->tape_sigmaG) {
                if (!fh) sigmaG[g] = a->tape_sigmaG[(size_t)itx * G + g];
                for (int k = 0; k < K; k++) estPi[g * K + k] = a->tape_pi[((size_t)itx * G + g) * K + k];
            } else {
                if (!fh) sigmaG[g] = ho_inv_scaled_chisq(&mt_hyper, v0G + (double)m0[g],
                                                         (bsq[g] * (double)m0[g] + v0G * s02G) / (v0G + (double)m0[g])); /* :2570 */
                double s = 0.0;                                 /* :2576-2577 dirichlet(cass + dirc), dirc = 1 unless --dPriorsFile */
                for (int k = 0; k < K; k++) {
                    const double dk = a->dirichlet ? a->dirichlet[g * K + k] : 1.0;
                    estPi[g * K + k] = ho_mt_gamma(&mt_hyper, (double)sum_cass[g * K + k] + dk); s += estPi[g * K + k];
                }
                for (int k = 0; k < K; k++) estPi[g * K + k] /= s;
            }
        }
        /* :2648-2681 fixed effects. The reference draws gamma on every rank from the rank's own stream without a broadcast
         * (the ranks' epsilon then diverge); here, as for the other hyper-parameters, ONE value per covariate
         * (rank 0's view: its epsilon, the hyper-parameter stream) is applied to every task. sigmaF = s02F = 1 (:1605, :2680). */
        if (a->F > 0) {
            const int F = a->F;
            const double sigE_sigF = sigmaE / 1.0;
            if (a->tape_xI) for (int i = 0; i < F; i++) xI[i] = a->tape_xI[(size_t)itx * F + i];
            else ho_shuffle(&mt_hyper, xI, F);                   /* :2653 std::shuffle(xI) */
            for (int i = 0; i < F; i++) {
                const int f = xI[i];
                const double gamma_old = gam[f];
                double num_f = 0.0;
                for (int k = 0; k < N; k++) num_f += a->X[(size_t)k * F + f] * (eps[0][k] + gamma_old * a->X[(size_t)k * F + f]);   /* :2666-2668 */
                const double denom_f = dNm1 + sigE_sigF;                                                                     /* :2670 */
                const double z = a->tape_zcov ? a->tape_zcov[(size_t)itx * F + i] : ho_mt_normal(&mt_hyper);
                gam[f] = num_f / denom_f + sqrt(sigmaE / denom_f) * z;                                                       /* :2671 */
                for (int r = 0; r < T; r++)
                    for (int k = 0; k < N; k++) eps[r][k] = eps[r][k] + (gamma_old - gam[f]) * a->X[(size_t)k * F + f];     /* :2673-2676 */
            }
            if (a->out_gamma) memcpy(a->out_gamma + (size_t)itx * F, gam, sizeof(double) * F);
        }
        /* :2685-2690 (rank 0's epsilon; its sigmaE is broadcast :2705) */
        double e_sqn = 0.0;
        for (int i = 0; i < N; ++i) e_sqn += eps[0][i] * eps[0][i];
        if (a->tape_sigmaE) sigmaE = a->tape_sigmaE[itx];
        else sigmaE = ho_inv_scaled_chisq(&mt_hyper, v0E + dN, (e_sqn + v0E * s02E) / (v0E + dN));

        /* record */
        if (a->out_beta) memcpy(a->out_beta + (size_t)itx * Mtot, Beta, sizeof(double) * Mtot);
        if (a->out_comp) memcpy(a->out_comp + (size_t)itx * Mtot, components, sizeof(int) * Mtot);
        if (a->out_acum) memcpy(a->out_acum + (size_t)itx * Mtot, Acum, sizeof(double) * Mtot);
        if (a->out_mu) memcpy(a->out_mu + (size_t)itx * T, mu, sizeof(double) * T);
        if (a->out_eps) for (int r = 0; r < T; r++) memcpy(a->out_eps + ((size_t)itx * T + r) * N, eps[r], sizeof(double) * N);
        if (a->out_sigmaG) memcpy(a->out_sigmaG + (size_t)itx * G, sigmaG, sizeof(double) * G);
        if (a->out_sigmaE) a->out_sigmaE[itx] = sigmaE;
        if (a->out_pi) memcpy(a->out_pi + (size_t)itx * G * K, estPi, sizeof(double) * G * K);
        if (a->out_bsq) memcpy(a->out_bsq + (size_t)itx * G, bsq, sizeof(double) * G);
        if (a->out_cass) memcpy(a->out_cass + (size_t)itx * G * K, sum_cass, sizeof(int) * G * K);
        if (a->out_esqn) a->out_esqn[itx] = e_sqn;
        if (fh && a->out_fh) {
            double *o = a->out_fh + (size_t)itx * (3 + G);
            o[0] = hypTau; o[1] = tau; o[2] = scaledBSQN;
            for (int g = 0; g < G; g++) o[3 + g] = c_slab[g];
        }
        if (fh && a->out_lambda) memcpy(a->out_lambda + (size_t)itx * Mtot, lambda_var, sizeof(double) * Mtot);
        if (fh && a->out_nu) memcpy(a->out_nu + (size_t)itx * Mtot, nu_var, sizeof(double) * Mtot);
    }

    for (int r = 0; r < T; r++) { free(eps[r]); free(tmpEps[r]); free(dEpsSum[r]); free(deltaEpsT[r]); }
    free(eps); free(tmpEps); free(dEpsSum); free(deltaEpsT); free(deltaSum); free(y);
    free(Beta); free(components); free(Acum); free(adaV); free(mu); free(cass); free(sum_cass);
    free(bsq); free(m0); free(tsad);
    free(mave); free(mstd); free(sigmaG); free(MtotGrp); free(cVa); free(cVaI); free(estPi);
    free(MrankS); free(MrankL); free(xI); free(gam);
    free(c_slab); free(lambda_var); free(nu_var); free(gnu); free(glam);
    return 0;
}

/* Positional tape of spec v1 for one iteration: z_mu and the permutation come
 * from the task stream mt19937(seed + 1000*task) (src/BayesRRm.cpp:1228 seeding
 * rule; draw order :1682 then :1692), u/z from Philox.  `streams` holds T
 * persistent ho_mt states; perm_state holds the current markerI of every task
 * (identity at start; std::shuffle permutes the previous order in place). */
HO_API void ho_tape_iteration(ho_mt *streams, uint32_t seed, int T, int Mtot, uint32_t iteration,
                              int32_t *perm_state, double *zmu, int32_t *perm_out, double *u, double *z,
                              int shuffle) {
    int *MrankS = (int *)malloc(sizeof(int) * T), *MrankL = (int *)malloc(sizeof(int) * T);
    ho_define_blocks_of_markers(Mtot, MrankS, MrankL, (uint)T);
    for (int r = 0; r < T; r++) {
        zmu[r] = ho_mt_normal(&streams[r]);
        if (shuffle) ho_shuffle(&streams[r], perm_state + MrankS[r], MrankL[r]);
        for (int j = 0; j < MrankL[r]; j++) {
            perm_out[MrankS[r] + j] = perm_state[MrankS[r] + j];
            ho_marker_draws(seed, (uint32_t)r, iteration, (uint32_t)j, &u[MrankS[r] + j], &z[MrankS[r] + j]);
        }
    }
    free(MrankS); free(MrankL);
}

HO_API int ho_sizeof_mt(void) { return (int)sizeof(ho_mt); }
HO_API void ho_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
HO_API int ho_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
