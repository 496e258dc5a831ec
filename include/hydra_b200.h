/*
 * hydra_b200.h -- C ABI of the B200-native (sm_100a) per-marker Gibbs hot path
 * of medical-genomics-group/hydra (BayesRRm / BayesW).
 *
 * hydra has no plugin/FFI interface: its seam is the member-function set of
 * class BayesRRm (reference src/BayesRRm.h:116-150) plus the state of the marker
 * loop in BayesRRm::runMpiGibbs (src/BayesRRm.cpp:933-2939).  Every entry point
 * below names the reference code it replaces.  Plain pointers and sizes only;
 * all host buffers are owned by the caller, all device memory by hb_ctx.
 * The ABI is not re-entrant per ctx (one host thread per GPU, like one MPI rank).
 *
 * Every function returns HB_OK (0) or a negative error code; hb_last_error()
 * returns the message of the last failure on the calling thread (the host maps
 * it to the reference's "FATAL  : ..." line + non-zero exit, cf.
 * src/mpi_utils.hpp:19-36 check_mpi/check_malloc -> MPI_Abort).
 * There is NO CPU fallback: without a CUDA device every call fails loudly.
 */
#ifndef HYDRA_B200_H
#define HYDRA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define HB_OK 0
#define HB_ERR_CUDA -1
#define HB_ERR_ARG -2
#define HB_ERR_STATE -3
#define HB_ERR_NCCL -4
#define HB_ERR_NOMEM -5

#define HB_ABI_VERSION 2

typedef struct hb_ctx hb_ctx;

/* Genotype representation choice, src/data.cpp:886-1069 (mixed), options.hpp:86 */
#define HB_REPR_SPARSE 0 /* all markers as index lists      (--sparse-dir/--sparse-basename) */
#define HB_REPR_BED 1    /* all markers as 2-bit BED bytes  (dotp_lut path)                   */
#define HB_REPR_MIXED 2  /* USEBED iff (n1+n2+nm)/N > threshold_fnz (src/data.cpp:931-932); such a marker is kept as
                          * 2-bit codes in HBM only where that halves its 16-bit index lists (> 25 % non-zeros):
                          * the kernel's sums are exact, results do not depend on the stored form */

typedef struct hb_config {
    int32_t device;           /* CUDA device ordinal */
    uint32_t n_ind_raw;       /* individuals in the genotype files (--number-individuals) */
    uint32_t n_na;            /* individuals with NA phenotype (data.numNAs) */
    const uint32_t *na_inds;  /* their raw indices, ascending (data.NAsInds, src/data.cpp:1827) */
    uint32_t m_total;         /* Mtot (--number-markers) */
    uint32_t n_tasks_total;   /* hydra "tasks" (= MPI ranks of the reference run being reproduced) */
    uint32_t task_first;      /* first task hosted by this device */
    uint32_t n_tasks_local;   /* tasks hosted by this device (contiguous) */
    const int32_t *block_starts; /* optional --marker-blocks-file (src/data.cpp:1391-1440): n_tasks_total */
    const int32_t *block_lens;   /*   starts (0-based) / lengths; NULL = even split, BayesRRm.cpp:396-413 */
    uint32_t sync_rate;       /* --sync-rate (0 behaves as 1, src/BayesRRm.cpp:2044) */
    uint32_t n_groups;        /* numGroups */
    uint32_t n_mix;           /* K = mixtures + 1 (zero component first) */
    int32_t repr_mode;        /* HB_REPR_* */
    double threshold_fnz;     /* --threshold-fnz (default 0.06) */
    uint32_t n_slices;        /* 0 = auto; individuals are split in n_slices shared-memory slices */
    uint32_t max_ctas;        /* 0 = one CTA per SM */
    uint32_t model;           /* 0 = BayesRRm, 1 = BayesW */
    uint32_t reserved[7];
} hb_config;

/* ---- life cycle ------------------------------------------------------- */
int hb_abi_version(void);
int hb_sizeof_config(void);   /* sizeof(hb_config): lets a foreign-language binding check its struct layout */
int hb_sizeof_iter_out(void); /* sizeof(hb_brr_iter_out) */
int hb_sizeof_brr_tape(void);
const char *hb_last_error(void);
int hb_create(const hb_config *cfg, hb_ctx **out);
void hb_destroy(hb_ctx *ctx);
/* layout facts: N after NA removal, local marker range, slices, slice length, replica groups */
int hb_get_layout(hb_ctx *ctx, uint32_t *n_ind, uint32_t *m_start, uint32_t *m_local, uint32_t *n_slices,
                  uint32_t *slice_len, uint32_t *n_groups_of_ctas, uint32_t *lmax);
/* task blocks, src/BayesRRm.cpp:781-827 mpi_assign_blocks_to_tasks: arrays of n_tasks_total */
int hb_get_task_blocks(hb_ctx *ctx, int32_t *starts, int32_t *lens);

/* ---- genotype staging (replaces src/data.cpp:671-1313) ----------------- */
/* Raw PLINK BED columns (3-byte header already skipped, snpLenByt = ceil(n_ind_raw/4) bytes
 * per marker, src/data.cpp:685,700) for local markers [m_first, m_first+n). Decoding rule and
 * index order are those of Data::sparse_data_fill_indices (src/data.cpp:1224-1290); NA-phenotype
 * individuals are compacted away as in sparse_data_correct_for_missing_phenotype (:1112-1158). */
int hb_stage_bed(hb_ctx *ctx, uint32_t m_first, uint32_t n, const uint8_t *bed_cols);
/* Reference sparse triple (raw, i.e. NOT NA-corrected indices), src/data.cpp:742-823:
 * starts N?S are relative to the passed I? arrays. */
int hb_stage_sparse(hb_ctx *ctx, uint32_t m_first, uint32_t n,
                    const uint32_t *I1, const uint64_t *N1S, const uint64_t *N1L,
                    const uint32_t *I2, const uint64_t *N2S, const uint64_t *N2L,
                    const uint32_t *IM, const uint64_t *NMS, const uint64_t *NML);
/* Bench/test scaffold: counter-based synthetic genotypes generated on the device
 * (SURVEY.md 8(d)); thresholds[n][3], attempts[n] (NULL = 0); j_global0 = global id of m_first. */
int hb_stage_synth(hb_ctx *ctx, uint32_t m_first, uint32_t n, uint32_t seed,
                   const uint32_t *thresholds, const uint32_t *attempts);
/* After all local markers are staged: mave/mstd (src/BayesRRm.cpp:1502-1508; BayesW: :1211-1232) */
int hb_stage_finalize(hb_ctx *ctx);

int hb_marker_counts(hb_ctx *ctx, uint32_t *n1, uint32_t *n2, uint32_t *nm); /* m_local each */
int hb_marker_stats(hb_ctx *ctx, double *mave, double *mstd);                /* m_local each */
int hb_marker_is_bed(hb_ctx *ctx, uint8_t *flags);                           /* USEBED, m_local */
uint64_t hb_genotype_bytes(hb_ctx *ctx);                                     /* HBM bytes of marker records */
/* Round trip back to the reference representation (bit-exact parity checks and --bed-to-sparse):
 * lists for local markers [m_first, m_first+n); starts relative to the output arrays. */
int hb_export_sparse(hb_ctx *ctx, uint32_t m_first, uint32_t n,
                     uint32_t *I1, uint64_t *N1S, uint64_t *N1L,
                     uint32_t *I2, uint64_t *N2S, uint64_t *N2L,
                     uint32_t *IM, uint64_t *NMS, uint64_t *NML);
/* NA-compacted BED bytes of one marker (ceil(N/4) bytes, pad bits 00), src/data.cpp:826-865 */
int hb_export_bed(hb_ctx *ctx, uint32_t m, uint8_t *out);

/* ---- unit-level kernels (parity tests; reference member functions) ------ */
int hb_set_epsilon(hb_ctx *ctx, const double *eps); /* N doubles */
int hb_get_epsilon(hb_ctx *ctx, double *eps);
/* num[i] = mstd*(x_j . eps), centred/scaled column: BayesRRm::sparse_dotprod (src/BayesRRm.cpp:316-342)
 * for list markers, the dotp_lut loop (:1757-1809) for BED markers */
int hb_dot_markers(hb_ctx *ctx, const uint32_t *markers, uint32_t n, double *num);
/* eps += sum_i deltaEps(marker_i, dbeta_i): sparse_scaadd (:250-281) / LUT deltaEps (:1976-2010)
 * followed by sum_vectors_f64 (:2022) and the epsilon update (:2460-2471) */
int hb_scaadd_markers(hb_ctx *ctx, const uint32_t *markers, const double *dbeta, uint32_t n);

/* ---- BayesRRm chain (src/BayesRRm.cpp:1565-2731) ------------------------- */
/* y: N raw phenotypes (centred and scaled inside, :371-388); groups: m_total ints; mS: n_groups*n_mix
 * row-major with column 0 = 0; sigmaG0: initial sigmaG per group (:1233), NULL = draw by RNG spec v1 */
int hb_brr_init(hb_ctx *ctx, const double *y, const int32_t *groups, const double *mS,
                const double *sigmaG0, uint32_t seed);

typedef struct hb_brr_tape {
    /* positional draw tape for ONE iteration, local tasks only (DESIGN.md "Draw tape") */
    const double *zmu;    /* n_tasks_local standard normals for mu (:1682) */
    const int32_t *perm;  /* m_local: marker order of each local task block, task-local indices (:1692) */
    const double *u;      /* m_local: U(0,1) of the marker processed at step j of each task (:1880) */
    const double *z;      /* m_local: standard normal for beta (:1901), used iff component > 0 */
    const double *sigmaG; /* n_groups values AFTER this iteration (:2570), NULL = draw */
    const double *pi;     /* n_groups*n_mix (:2577), NULL = draw */
    const double *sigmaE; /* 1 value (:2690), NULL = draw */
    const int32_t *xI;    /* n_covariates: order of the fixed effects in this iteration (std::shuffle(xI), :2653), NULL = draw */
    const double *zcov;   /* n_covariates standard normals for gamma (:2671), NULL = draw */
    /* bayesFH (hb_brr_set_fh); all three ignored otherwise */
    const double *gnu;    /* m_local standard Gamma(0.5 + 0.5*v0L, 1) variates behind the nu_var draws (:1729), by local marker */
    const double *glam;   /* m_local, the lambda_var draws (:1952) */
    const double *fh_hyper; /* n_groups*3: hypTau, tau, c_slab[g] as drawn in group g's pass (:2559-2562), NULL = draw */
} hb_brr_tape;

typedef struct hb_brr_iter_out {
    double sigmaE;        /* after the iteration */
    double e_sqn;         /* sum eps^2 of task 0 (:2685-2686) */
    double epssum;        /* sum of (eps + mu) at iteration start (:1677-1678) */
    double loop_ms;       /* device time of the marker loop (CUDA events) */
    double iter_ms;       /* device time of the whole iteration */
    uint64_t n_sync;      /* epsilon synchronisations performed (:2069) */
    uint64_t n_windows;   /* sync windows processed */
    uint64_t n_launches;  /* kernels launched by this call */
    uint64_t nnz_processed;   /* stored non-zeros visited by the dot products */
    uint64_t nnz_updated;     /* stored non-zeros visited by epsilon updates */
    uint64_t bed_markers;     /* markers processed through the BED path */
    uint64_t markers_changed; /* markers with deltaBeta != 0 */
    uint64_t phase_cycles[8]; /* CTA 0 SM cycles: table, dot, publish+draw, grid barrier, update set-up, slice sum, update loads, update apply */
    uint64_t windows_ahead;   /* windows of sync_rate steps taken where the reference synchronises after every step (no change expected) */
    uint64_t draws_repeated;  /* marker draws of such windows that were discarded and repeated after an earlier step changed a marker */
} hb_brr_iter_out;

/* One Gibbs iteration: mu, marker loop (sync windows), group statistics, hyper-parameters.
 * tape == NULL: RNG spec v1 (host mt19937 streams + device Philox). Per-group / per-task results
 * are fetched with hb_brr_get_hyper / hb_brr_get_state. */
int hb_brr_iteration(hb_ctx *ctx, const hb_brr_tape *tape, hb_brr_iter_out *out);
int hb_brr_get_hyper(hb_ctx *ctx, double *sigmaG, double *pi, double *sigmaE, double *mu_tasks_local,
                     double *bsq, int32_t *cass, int32_t *m0);
/* Beta, components, Acum of the local markers (the slices written to .bet/.cpn/.acu, :2779-2785) */
int hb_brr_get_state(hb_ctx *ctx, double *beta, int32_t *components, double *acum);
/* The same for a writer that thins every iteration: returns at once, the copies (into the caller's buffers, page-locked for a
 * true DMA) run on a second stream while the next hb_brr_iteration computes; the buffers are the library's until
 * hb_brr_state_wait returns. */
int hb_brr_get_state_async(hb_ctx *ctx, double *beta, int32_t *components, double *acum);
int hb_brr_state_wait(hb_ctx *ctx);
int hb_brr_set_state(hb_ctx *ctx, const double *beta, const int32_t *components); /* --restart */
/* The reference's restart from its OUTPUT files (src/BayesRRm.cpp:842-928, data.cpp read_mcmc_output_*): what .csv, .xbet, .xcpn,
 * .mus.<task>, .eps.<task>, .mrk.<task> of the last save point hold is put back (arrays of the local markers / tasks, eps of local
 * task 0, perm_local may be NULL); the next hb_brr_iteration is iteration `iterations_done`. The random streams are not in those
 * files: they are re-seeded from (seed, iterations_done) -- a continuation with fresh draws; hb_brr_load_state continues bit for bit. */
int hb_brr_restore_outputs(hb_ctx *ctx, uint32_t iterations_done, const double *sigmaG, const double *pi, double sigmaE,
                           const double *mu_tasks_local, const double *beta, const int32_t *components, const double *eps_task0,
                           const int32_t *perm_local);
/* --restart (src/BayesRRm.cpp:842-928: the reference reads .csv .bet .cpn .eps .mrk .mus .rng back). The complete chain
 * state of this GPU as one opaque blob: hyper-parameters, per-task mu, residual, effects / components / Acum and the host
 * random streams. `buf == NULL`: only *need is set. After hb_brr_init (or hb_bw_init: the pair serves both models) with the
 * same inputs and layout, hb_brr_load_state continues the chain bit-identically to the uninterrupted run. */
int hb_brr_save_state(hb_ctx *ctx, void *buf, size_t cap, size_t *need);
int hb_brr_load_state(hb_ctx *ctx, const void *buf, size_t n);
/* epsilon of local task t as the reference dumps it to .eps.<rank> (:2827) */
int hb_brr_get_task_epsilon(hb_ctx *ctx, uint32_t task_local, double *eps);
/* random stream of local task t as text (.rng.<rank>, :2805, src/distributions_boost.cpp:38-44: the reference dumps its
 * boost::mt19937 with operator<<; this is the std::mt19937 of RNG spec v1 in libstdc++'s text form). buf == NULL: only *need. */
int hb_brr_get_task_rng(hb_ctx *ctx, uint32_t task_local, char *buf, size_t cap, size_t *need);
/* current marker order of local task t (.mrk.<rank>, :2828) */
int hb_brr_get_task_perm(hb_ctx *ctx, uint32_t task_local, int32_t *perm);

/* Fixed effects, hydra's --covariates (src/BayesRRm.cpp:2648-2681, reader src/data.cpp:1615-1672): X is n_ind x n_cov,
 * row-major, used as read (the reference neither centres nor scales it; denom = (N-1) + sigmaE/sigmaF assumes standardised
 * columns, sigmaF = s02F = 1, src/BayesRRm.h:34). Call after hb_brr_init; every hb_brr_iteration then updates gamma and
 * epsilon between the group hyper-parameters and sigmaE. One gamma per covariate for all tasks, drawn from the
 * hyper-parameter stream (the reference draws on every rank from the rank's own stream and does not broadcast,
 * its ranks' residuals diverge; flagged, not reproduced). n_cov == 0 switches the block off. */
int hb_brr_set_covariates(hb_ctx *ctx, const double *X, uint32_t n_cov);
/* gamma[n_cov] and the current order xI[n_cov] (either may be NULL); the .gam / .xiv files of :2811-2831 */
int hb_brr_get_gamma(hb_ctx *ctx, double *gamma, int32_t *xI);

/* Per-group priors, hydra's --groupPriorsFile (v0G, s02G of the sigmaG draw per group, src/BayesRRm.cpp:2545-2548, reader
 * src/data.cpp:2034-2061) and --dPriorsFile (Dirichlet parameters of the pi draw per group, :2551-2554, 2576-2577, reader
 * src/data.cpp:2069-2096): v0G_s02G is n_groups x 2, dirichlet n_groups x n_mix, row-major; NULL = the built-in constants
 * (0.0001, 0.0001 and 1.0). Call after hb_brr_init. */
int hb_brr_set_group_priors(hb_ctx *ctx, const double *v0G_s02G, const double *dirichlet);

/* bayesFHMPI (hydra --bayesType bayesFHMPI; src/BayesRRm.cpp:1125-1163 initialisation, :1727-1731 nu_var and the marker's own
 * prior variance, :1747-1748 / :1869-1872 its use in the mixture draw, :1942-1952 lambda_var, :2503-2510 scaled sum of squares,
 * :2557-2565 hypTau / tau / c_slab; defaults src/options.hpp:91-96). Call after hb_brr_init (cfg == NULL switches it off);
 * state0 = { hypTau, tau, c_slab[n_groups] } or NULL = drawn from the hyper-parameter stream in the reference's order. The two
 * per-marker draws leave the marker loop (they depend on the marker's own previous state only): one kernel before it, one after.
 * The reference neither reduces the scaled sum of squares over its ranks nor broadcasts tau / c_slab: its ranks' FH parameters
 * diverge (flagged, not reproduced). Here one set of parameters, as in a one-rank run of the reference: the sum runs over all
 * markers of all tasks and GPUs, tau / hypTau / c_slab are drawn on every GPU from the common hyper-parameter stream. */
typedef struct hb_fh_config {
    double v0L, v0t, v0c, s02c, tau0;
} hb_fh_config;
int hb_brr_set_fh(hb_ctx *ctx, const hb_fh_config *cfg, const double *state0);
/* scalars3 = { hypTau, tau, scaledBSQN of the last iteration }, c_slab[n_groups], lambda_var / nu_var of the local markers
 * (any may be NULL) */
int hb_brr_get_fh(hb_ctx *ctx, double *scalars3, double *c_slab, double *lambda_var, double *nu_var);

/* ---- BayesW chain (Weibull survival; src/BayesW.cpp:905-1907) ------------------------------------------
 * Several GPUs (hb_comm_init before hb_bw_init, one process per GPU): markers and tasks are split as for BayesRRm, epsilon is
 * replicated; per synchronisation window every GPU's epsilon change is summed with ncclAllReduce over NVLink / NVSwitch (the
 * reference's MPI_Allreduce of deltaEps, :1799-1835) and added on every GPU, so the replicas stay bit-identical; beta_squaredNorm
 * and cass are all-reduced per iteration (:1866-1867); mu, alpha and the fixed effects are drawn on every GPU from the same
 * streams on the same residual. */
/* hb_config.model = 1. y: N log-times, fail: N failure indicators (0/1; src/data.cpp:1779), mS as for BayesRRm,
 * quad_points in {3,5,7,9,11,13,15,17,25} (--quad_points, :706-709). Initial values as BayesW::init (:728-853). */
int hb_bw_init(hb_ctx *ctx, const double *y, const double *fail, const int32_t *groups, const double *mS,
               uint32_t quad_points, uint32_t seed);

typedef struct hb_bw_tape {
    const int32_t *perm;  /* m_local: marker order of each local task block (:1457-1459) */
    const double *p;      /* m_local: U(0,1) of the marker processed at step j of each task (:1528) */
    const double *sigmaG; /* n_groups values AFTER this iteration (:1893), NULL = draw */
    const double *pi;     /* n_groups*n_mix (:1899-1903), NULL = draw */
    const int32_t *xI;    /* n_cov: order of the fixed effects in this iteration (:1369), NULL = shuffled from the hyper-parameter stream */
} hb_bw_tape;

typedef struct hb_bw_iter_out {
    double mu, alpha;     /* after the iteration's ARMS draws (:1336-1363, :1421-1447) */
    double loop_ms, iter_ms;
    uint64_t n_sync, n_windows, n_launches, markers_changed;
    uint64_t density_evals; /* N-sum evaluations of the mu / alpha log-densities (K11) */
} hb_bw_iter_out;

/* One Gibbs iteration: ARMS for mu and alpha (N-sums of exp on the device), marker loop, sigmaG / pi_L draws.
 * ARMS uniforms: RNG spec v1 (Philox, 31-bit integers in the form of src/BayesW_arms.cpp:913-918). */
int hb_bw_iteration(hb_ctx *ctx, const hb_bw_tape *tape, hb_bw_iter_out *out);
/* Fixed effects of BayesW (--covariates with bayesWMPI; src/BayesW.cpp:1366-1413, gamma_dens :119-129): X is N x n_cov, row-major,
 * as read from the covariate file. Call after hb_bw_init; every hb_bw_iteration then draws each gamma by ARMS on gamma +- 0.075
 * (between mu and alpha, in the iteration's order; the density's N-sum runs on the device, uniforms: Philox stream 'ARMG') and
 * moves the residual. n_cov == 0 switches the block off. hb_bw_get_gamma: gamma[n_cov] and the last order xI[n_cov]. */
int hb_bw_set_covariates(hb_ctx *ctx, const double *X, uint32_t n_cov);
int hb_bw_get_gamma(hb_ctx *ctx, double *gamma, int32_t *xI);
int hb_bw_get_hyper(hb_ctx *ctx, double *sigmaG, double *pi, double *mu, double *alpha, double *bsq, int32_t *cass, int32_t *m0);
/* marker statistics of BayesW: mstd is the SD (not its inverse), sum_failure (:1211-1232); m_local each */
int hb_bw_marker_stats(hb_ctx *ctx, double *sd, double *sum_failure);
/* unit-level kernels for parity tests */
/* (vi_sum, vi_1, vi_2, vi_0) of markers against the current epsilon with the marker's own effect beta_old removed
 * (partial_sum :49-65 and :1499-1525); out: n*4 */
int hb_bw_vi_sums(hb_ctx *ctx, const uint32_t *markers, const double *beta_old, uint32_t n, double alpha, double *out);
/* sum_i exp(a*eps_i + b): the N-sum inside mu_dens / alpha_dens / gamma_dens (:77-142) */
int hb_bw_sum_exp(hb_ctx *ctx, double a, double b, double *out);
/* pars[10] = alpha, sigmaG, sum_failure, vi_sum, vi_0, vi_1, vi_2, mean, sd, mean/sd */
int hb_bw_marginal_likelihoods(hb_ctx *ctx, uint32_t quad_points, const double *pars, const double *prior, const double *cVa,
                               uint32_t km1, double *post);
/* one ARMS draw of beta on the device (:1562-1582); out[4] = beta, error code, density evaluations, uniforms used */
int hb_bw_arms_beta(hb_ctx *ctx, const double *pars, double C_k, double sum_sigmaG, double beta_old, uint32_t seed, uint32_t task,
                    uint32_t iteration, uint32_t j, double *out);

/* ---- multi-GPU (replaces MPI_Allreduce, src/BayesRRm.cpp:2051,2456,2517-2518) --- */
#define HB_NCCL_ID_BYTES 128
int hb_comm_get_unique_id(uint8_t id[HB_NCCL_ID_BYTES]);
int hb_comm_init(hb_ctx *ctx, const uint8_t id[HB_NCCL_ID_BYTES], int rank, int nranks);
/* Fails (HB_ERR_ARG, message names `what`) unless every GPU of the run passes the same n <= 64 values: the seed (the
 * reference broadcasts rank 0's hyper-parameter draws, src/BayesRRm.cpp:2585, 2705, 2731; here every GPU draws them from
 * the same stream, so the seed must be common) and the restart point of --restart (:842-928). No-op on one GPU. */
int hb_comm_check_equal(hb_ctx *ctx, const uint64_t *vals, uint32_t n, const char *what);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* HYDRA_B200_H */
