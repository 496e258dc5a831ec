// hydra_b200.cu -- C ABI (include/hydra_b200.h) of the B200-native BayesRRm hot path.
// Host side: context, genotype staging, per-iteration driver of the window kernel.
// There is no CPU fallback in this file: every entry point needs a CUDA device.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <memory>
#include <new>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "brr_kernel.cuh"
#include "bw_kernels.cuh"
#include "common.cuh"
#include "convert.cuh"
#include "cov_kernel.cuh"
#include "fh_kernels.cuh"
#include "host_rng.hpp"

#include <nccl.h>

#define HB_NCCL(x)                                                                                  \
    do {                                                                                            \
        ncclResult_t r_ = (x);                                                                      \
        if (r_ != ncclSuccess) {                                                                    \
            hb::set_error("%s:%d NCCL error in %s: %s", __FILE__, __LINE__, #x, ncclGetErrorString(r_)); \
            return HB_ERR_NCCL;                                                                     \
        }                                                                                           \
    } while (0)


namespace hb {

static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    int alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
            p = nullptr;
            return HB_ERR_NOMEM;
        }
        n = count;
        return HB_OK;
    }
    int ensure(size_t count) { return (count <= n && p) ? HB_OK : alloc(count); }
    int zero(cudaStream_t s) {
        HB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
        return HB_OK;
    }
};

}  // namespace hb

using namespace hb;

struct hb_ctx {
    hb_config cfg{};
    int dev = 0, n_sms = 0;
    uint32_t Nraw = 0, N = 0, S = 0, L = 0, R = 0;
    uint32_t Mtot = 0, M = 0, m_start = 0;          // local markers [m_start, m_start+M)
    uint32_t Ttot = 0, T = 0, t_first = 0, lmax = 0; // tasks
    uint32_t K = 0, G = 0, SR = 1;
    std::vector<int32_t> blkS, blkL;                // all tasks (global marker ids)
    std::vector<uint32_t> na;
    size_t smem_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};

    // genotype records
    DevBuf<uint32_t> d_rmap;
    DevBuf<uint64_t> d_rec;
    DevBuf<double> d_mave, d_mstd;
    DevBuf<int32_t> d_grp;
    std::vector<void *> arenas;
    std::vector<uint32_t> n1, n2, nm;
    std::vector<uint8_t> is_bed, stored_bed, staged;   // is_bed: the reference's USEBED rule; stored_bed: the form of the record in HBM
    std::vector<uint64_t> rec_h;
    std::vector<double> mave_h, mstd_h;
    uint64_t geno_bytes = 0;
    bool finalized = false;
    // staging scratch
    DevBuf<uint8_t> d_raw;
    DevBuf<uint32_t> d_cnt3, d_start, d_meta;

    // epsilon
    DevBuf<double> d_E[2];
    int cur = 0;
    double shift = 0.0;  // true eps(task 0) = E + shift
    DevBuf<double> d_small;  // [0]=off, [1..S]=slice_sum, [1+S..2S]=slice_sq, then bsq[G]
    std::vector<double> slice_sum_h;
    bool eps_set = false;
    // fixed-point grid of the marker kernel: the largest sum of |eps| over one slice and the largest |eps| (of the stored
    // residual, before the scalar `shift` is folded in); refreshed by every launch, set by hb_set_epsilon / load_state
    double eps_slice_abs = 0.0, eps_abs_max = 0.0;
    int last_sh = 0;

    // marker state
    DevBuf<double> d_beta, d_acum;
    // hb_brr_get_state_async: snapshot + copy stream, so that the read-back of iteration i overlaps the marker loop of i+1
    DevBuf<double> d_snap_beta, d_snap_acum;
    DevBuf<int32_t> d_snap_comp;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_snap = nullptr, ev_copy = nullptr;
    bool copy_pending = false;
    DevBuf<int32_t> d_comp, d_cass;
    // per-iteration inputs
    DevBuf<int32_t> d_order, d_perm, d_task_len, d_task_off;
    DevBuf<uint32_t> d_wts;   // per marker: stored non-zeros (BED: N/2), the weight used to balance the CTA groups
    bool balance = false;
    DevBuf<WinMeta> d_wmeta;   // window-ordered marker data of the iteration
    DevBuf<uint4> d_dirw;     // window-ordered slice directory entries [S][lmax*T]
    DevBuf<double> d_u, d_z, d_ut, d_zt;
    DevBuf<double> d_hyp;
    DevBuf<uint8_t> d_active;
    // scratch
    DevBuf<uint4> d_slots;
    DevBuf<ChgEnt> d_chg_list;
    DevBuf<uint4> d_chg_dir;
    DevBuf<double> d_dB, d_dMave, d_num, d_bsq_part;
    DevBuf<uint64_t> d_dRec;
    DevBuf<uint32_t> d_chg_cnt, d_chg_off, d_bar, d_markers;
    DevBuf<unsigned long long> d_stats, d_ctacyc;
    bool debug_cycles = false;
    uint32_t Wmax = 0;

    // BayesW
    bool bw_ready = false;
    int bw_rule = -1;
    double bw_alpha = 0.0, bw_mu = 0.0, bw_d = 0.0, bw_sumSigmaG = 0.0;
    std::vector<double> bw_fail_h, bw_sff;   // failure indicators (host copy) and sum_failure_fix of the fixed effects (src/BayesW.cpp:1235-1237)
    uint64_t bw_evals = 0;
    DevBuf<double> d_sd, d_sumfail, d_fail, d_bwsc, d_bw_vi, d_bw_delta;
    DevBuf<uint32_t> d_bw_ctl;   // window sequence of the BayesW marker loop, two blocks of 8 (bw_kernels.cuh)
    std::vector<double> sd_h, sumfail_h;

    // chain (host)
    bool brr_ready = false;
    uint32_t seed = 0, iteration = 0;
    std::vector<double> cVa, cVaI, pi, sigmaG, mu, bsq;
    std::vector<int32_t> cass, m0, MtotGrp, groups_local;
    std::vector<uint8_t> active;
    double sigmaE = 0.0;
    std::vector<HostRng> task_rng;
    HostRng hyper_rng;
    // fixed effects (--covariates)
    uint32_t F = 0;
    DevBuf<double> d_X, d_gamma, d_zcov, d_covpart;
    DevBuf<int32_t> d_xI;
    DevBuf<uint32_t> d_covbar;
    std::vector<double> gamma;
    std::vector<int32_t> xI;
    // per-group priors (--groupPriorsFile: v0G, s02G; --dPriorsFile: Dirichlet parameters); empty = the built-in constants
    std::vector<double> grp_priors, dirichlet;
    // bayesFHMPI (src/BayesRRm.cpp:1125-1163 ...): global scale tau and its hyper-parameter, slab variance per group, local scales
    bool fh_on = false;
    hb_fh_config fh_cfg{};
    double fh_hypTau = 0.0, fh_tau = 0.0, fh_sbsqn = 0.0;
    std::vector<double> fh_c;
    DevBuf<double> d_fh_lambda, d_fh_nu, d_fh_par, d_fh_c, d_fh_part, d_fh_g;
    std::vector<int32_t> perm;  // M: task-local order per local task block
    std::vector<int32_t> perm_next;   // the next iteration's order, shuffled on a worker thread during the marker loop
    void *pinned_perm[2] = {nullptr, nullptr};  // both buffers are page-locked (cudaHostRegister): the per-iteration H2D copy is a DMA
    std::vector<double> zmu_next;
    std::thread prefetch;
    bool have_next = false;
    uint32_t n_ahead = 0;       // steps of a window run ahead
    bool order_ready = false;   // the next iteration's window order (k_window_order) is already on the device
    double *pin = nullptr;      // pinned host scratch
    size_t pin_n = 0;

    // multi-GPU (one process per GPU)
    ncclComm_t nccl = nullptr;
    int rank = 0, nranks = 1;
    unsigned char *inbox = nullptr;             // this GPU's inbox, written by the peers over NVLink
    unsigned long long *flags = nullptr;        // [nranks] peers' sequence numbers (same allocation as the inbox)
    unsigned char *peer_inbox[kMaxRanks] = {};  // IPC mappings of the peers' inboxes
    size_t inbox_stride = 0, comm_bytes = 0;
    unsigned long long seq_base = 0;
    DevBuf<uint32_t> d_rec_bytes, d_err;
    DevBuf<double> d_red;                       // per-iteration statistics that are summed over the GPUs
    std::vector<uint32_t> rec_bytes_h;

    ~hb_ctx() {
        if (prefetch.joinable()) prefetch.join();
        for (void *q : pinned_perm) if (q) cudaHostUnregister(q);
        for (int h = 0; h < nranks; h++)
            if (h != rank && peer_inbox[h]) cudaIpcCloseMemHandle(peer_inbox[h]);
        if (inbox) cudaFree(inbox);
        if (nccl) ncclCommDestroy(nccl);
        for (void *a : arenas) cudaFree(a);
        if (pin) cudaFreeHost(pin);
        for (auto &e : ev)
            if (e) cudaEventDestroy(e);
        if (ev_snap) cudaEventDestroy(ev_snap);
        if (ev_copy) cudaEventDestroy(ev_copy);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace hb {

// contiguous blocks, first Mtot % n tasks get one extra marker (src/BayesRRm.cpp:396-413)
static void define_blocks(uint32_t Mtot, uint32_t nblocks, std::vector<int32_t> &S, std::vector<int32_t> &Lb) {
    S.resize(nblocks);
    Lb.resize(nblocks);
    const uint32_t modu = Mtot % nblocks;
    uint32_t start = 0;
    for (uint32_t i = 0; i < nblocks; ++i) {
        Lb[i] = (int32_t)(Mtot / nblocks) + ((modu != 0 && i < modu) ? 1 : 0);
        S[i] = (int32_t)start;
        start += (uint32_t)Lb[i];
    }
}

static size_t smem_for(uint32_t L) { return (((size_t)L + 2) * 8 + 15) / 16 * 16 + 2 * sizeof(ItemTab) + sizeof(ChgTab); }

static int ensure_scratch(hb_ctx *c, uint32_t W) {
    if (W <= c->Wmax) return HB_OK;
    W = (W + 255u) & ~255u;
    HB_TRY(c->d_slots.alloc((size_t)W * c->S));
    HB_TRY(c->d_chg_list.alloc((size_t)3 * W));
    HB_TRY(c->d_chg_off.alloc((size_t)3 * W));
    HB_TRY(c->d_chg_dir.alloc((size_t)3 * W * c->S));
    HB_TRY(c->d_chg_cnt.alloc(16));
    HB_TRY(c->d_dB.alloc((size_t)2 * W));
    HB_TRY(c->d_dMave.alloc((size_t)2 * W));
    HB_TRY(c->d_dRec.alloc((size_t)2 * W));
    HB_TRY(c->d_chg_list.zero(c->stream));
    HB_TRY(c->d_dB.zero(c->stream));
    c->Wmax = W;
    return HB_OK;
}

static int ensure_pin(hb_ctx *c, size_t n) {
    if (n <= c->pin_n) return HB_OK;
    if (c->pin) cudaFreeHost(c->pin);
    c->pin = nullptr;
    HB_CUDA(cudaMallocHost((void **)&c->pin, n * sizeof(double)));
    c->pin_n = n;
    return HB_OK;
}

static void fill_params(hb_ctx *c, BrrParams &P) {
    memset(&P, 0, sizeof(P));
    P.N = c->N; P.S = c->S; P.L = c->L; P.R = c->R; P.M = c->M;
    P.rec = c->d_rec.p; P.mave = c->d_mave.p; P.mstd = c->d_mstd.p; P.grp = c->d_grp.p;
    P.E_in = c->d_E[c->cur].p; P.E_out = c->d_E[c->cur ^ 1].p;
    P.shift_in = 0.0;
    P.off_out = c->d_small.p;
    P.slice_sum_out = c->d_small.p + 1;
    P.slice_sq_out = c->d_small.p + 1 + c->S;
    P.slice_abs_out = c->d_small.p + 1 + 2 * c->S + c->G;
    P.slice_max_out = c->d_small.p + 1 + 3 * c->S + c->G;
    P.beta = c->d_beta.p; P.comp = c->d_comp.p; P.acum = c->d_acum.p; P.cass = c->d_cass.p;
    P.order = c->d_order.p; P.u = c->d_u.p; P.z = c->d_z.p;
    P.T = 1; P.SR = 1; P.lmax = 0; P.K = c->K; P.G = c->G; P.n_ahead = 0; P.fh = nullptr;
    const size_t gk = (size_t)c->G * c->K;
    P.logPi = c->d_hyp.p; P.chalf = c->d_hyp.p + gk; P.denom = c->d_hyp.p + 2 * gk; P.sdk = c->d_hyp.p + 3 * gk;
    P.grp_active = c->d_active.p;
    P.dNm1 = (double)(c->N - 1);
    P.slots = c->d_slots.p; P.chg_cnt = reinterpret_cast<unsigned long long *>(c->d_chg_cnt.p); P.chg_list = c->d_chg_list.p; P.chg_off = c->d_chg_off.p; P.chg_dir = c->d_chg_dir.p; P.dB = c->d_dB.p; P.dMave = c->d_dMave.p; P.dRec = c->d_dRec.p; P.Wmax = c->Wmax;
    P.bar = c->d_bar.p; P.stats = c->d_stats.p;
    P.mode = MODE_CHAIN;
    P.num_out = c->d_num.p;
    P.cta_cycles = c->debug_cycles ? c->d_ctacyc.p : nullptr;
    P.flags = (getenv("HB_NO_PREFETCH") ? 1u : 0u) | (getenv("HB_DBG_FLAGS") ? (uint32_t)atoi(getenv("HB_DBG_FLAGS")) : 0u);   // developer knobs
    P.pc.nranks = (uint32_t)c->nranks; P.pc.rank = (uint32_t)c->rank;
    P.pc.T_total = c->Ttot; P.pc.t_first = c->t_first;
    P.pc.inbox_local = c->inbox;
    for (int h = 0; h < c->nranks; h++) P.pc.inbox_peer[h] = c->peer_inbox[h];
    P.pc.inbox_stride = c->inbox_stride; P.pc.seq_base = c->seq_base;
    P.pc.rec_bytes = c->d_rec_bytes.p; P.pc.err = c->d_err.p;
    {
        const char *t = getenv("HB_PEER_TIMEOUT_S");
        P.pc.timeout_cycles = (long long)((t ? atof(t) : 10.0) * 1.9e9);
    }
}

// All GPUs of a run must agree on a few scalars (the seed: hyper-parameters are drawn on every GPU from the same stream
// instead of the reference's MPI_Bcast from rank 0, src/BayesRRm.cpp:2585, 2705, 2731; the restart point). min == max over the ranks.
static int check_equal_over_ranks(hb_ctx *c, const uint64_t *vals, uint32_t n, const char *what) {
    if (c->nranks <= 1 || !c->nccl) return HB_OK;
    DevBuf<int64_t> d;
    HB_TRY(d.alloc(2 * (size_t)n));
    std::vector<int64_t> h(2 * (size_t)n);
    for (uint32_t i = 0; i < n; i++) { h[i] = (int64_t)vals[i]; h[n + i] = -(int64_t)vals[i]; }
    HB_CUDA(cudaMemcpyAsync(d.p, h.data(), sizeof(int64_t) * 2 * n, cudaMemcpyHostToDevice, c->stream));
    HB_NCCL(ncclAllReduce(d.p, d.p, 2 * (size_t)n, ncclInt64, ncclMax, c->nccl, c->stream));
    HB_CUDA(cudaMemcpyAsync(h.data(), d.p, sizeof(int64_t) * 2 * n, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    for (uint32_t i = 0; i < n; i++)
        HB_CHECK(h[i] == -h[n + i], HB_ERR_ARG, "%s differs between the GPUs of this run (value %u: min %lld, max %lld, here %llu)", what, i,
                 (long long)-h[n + i], (long long)h[i], (unsigned long long)vals[i]);
    return HB_OK;
}

// Fixed-point grid 2^-sh of one launch (brr_kernel.cuh): every partial sum of a dot product over one slice,
// S1 + 2 S2 <= 2 sum|eps|, must fit 63 bits, and every value 55 bits (the top byte of a slot is the update's tag). Both
// with a margin for what the residual can do during the launch (x2 on the sums, whose bound is already the worst case of
// an all-same-sign subset, x4 on single values); the kernel raises error 5 if a value leaves the range nevertheless.
static void pick_grid(hb_ctx *c, double shift_in, BrrParams &P) {
    const double a = c->eps_slice_abs + fabs(shift_in) * (double)c->L + 1.0;
    const double m = c->eps_abs_max + fabs(shift_in) + 1e-300;
    int sh = (int)floor(std::min(log2(ldexp(1.0, 61) / (2.0 * a)), log2(ldexp(1.0, 53) / m)));
    sh = std::max(8, std::min(sh, 60));
    if (const char *f = getenv("HB_FIXED_SHIFT")) sh = atoi(f);   // developer knob
    c->last_sh = sh;
    P.q_scale = ldexp(1.0, sh);
    P.q_inv = ldexp(1.0, -sh);
}

static int launch_window_kernel(hb_ctx *c, BrrParams &P) {
    pick_grid(c, P.shift_in, P);
    HB_CUDA(cudaMemsetAsync(c->d_bar.p, 0, sizeof(uint32_t), c->stream));
    HB_CUDA(cudaMemsetAsync(c->d_chg_cnt.p, 0, 16 * sizeof(uint32_t), c->stream));
    // the slots carry window tags starting at 1: clear what this launch can touch
    const size_t wuse = std::min<size_t>(c->Wmax, (P.mode == MODE_CHAIN) ? (size_t)std::max(std::max(1u, P.SR), P.n_ahead) * P.T : (size_t)P.lmax * P.T);
    HB_CUDA(cudaMemsetAsync(c->d_slots.p, 0, wuse * c->S * sizeof(uint4), c->stream));
    void *args[] = {(void *)&P};
    dim3 grid(c->S * c->R), block(kThreads);
    HB_CUDA(cudaLaunchCooperativeKernel((const void *)k_brr_iteration, grid, block, args, c->smem_bytes, c->stream));
    return HB_OK;
}

static int check_kernel_error(hb_ctx *c, const char *who) {
    uint32_t err = 0;
    HB_CUDA(cudaMemcpy(&err, c->d_err.p, sizeof(err), cudaMemcpyDeviceToHost));
    if (err != 0) {
        uint32_t zero = 0;
        cudaMemcpy(c->d_err.p, &zero, sizeof(zero), cudaMemcpyHostToDevice);
    }
    HB_CHECK(err != 5, HB_ERR_STATE, "%s: a residual left the fixed-point range of the marker kernel (grid 2^-%d): epsilon grew more than "
             "4x within one launch", who, c->last_sh);
    HB_CHECK(err == 0, HB_ERR_NCCL, "%s: exchange between the GPUs failed (code %u: 1 = inbox too small for the window's changed markers "
             "(HB_INBOX_MB) or more than %u of them on one GPU, 2 = a peer did not answer in time (HB_PEER_TIMEOUT_S), 3 = a peer reported a failure)",
             who, err, kMaxMerged);
    return HB_OK;
}

}  // namespace hb

extern "C" {

int hb_abi_version(void) { return HB_ABI_VERSION; }
int hb_sizeof_config(void) { return (int)sizeof(hb_config); }
int hb_sizeof_iter_out(void) { return (int)sizeof(hb_brr_iter_out); }
int hb_sizeof_brr_tape(void) { return (int)sizeof(hb_brr_tape); }
const char *hb_last_error(void) { return g_err; }

int hb_create(const hb_config *cfg, hb_ctx **out) {
    HB_CHECK(cfg && out, HB_ERR_ARG, "hb_create: null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    HB_CHECK(e == cudaSuccess && ndev > 0, HB_ERR_CUDA,
             "hb_create: no CUDA device (%s); hydra_b200 has no CPU fallback", cudaGetErrorString(e));
    HB_CHECK(cfg->device >= 0 && cfg->device < ndev, HB_ERR_ARG, "hb_create: device %d out of range (%d devices)", cfg->device, ndev);
    HB_CHECK(cfg->n_ind_raw > cfg->n_na + 1, HB_ERR_ARG, "hb_create: need at least 2 individuals with a phenotype");
    HB_CHECK(cfg->m_total > 0, HB_ERR_ARG, "hb_create: m_total == 0");
    HB_CHECK(cfg->n_tasks_total > 0 && cfg->n_tasks_local > 0 && cfg->task_first + cfg->n_tasks_local <= cfg->n_tasks_total,
             HB_ERR_ARG, "hb_create: bad task layout (total %u first %u local %u)", cfg->n_tasks_total, cfg->task_first, cfg->n_tasks_local);
    HB_CHECK(cfg->n_tasks_total <= cfg->m_total, HB_ERR_ARG, "hb_create: more tasks than markers");
    HB_CHECK(cfg->n_groups >= 1 && cfg->n_mix >= 2 && cfg->n_mix <= (uint32_t)kMaxMix, HB_ERR_ARG,
             "hb_create: n_groups >= 1 and 2 <= n_mix <= %d required", kMaxMix);
    HB_CHECK(cfg->repr_mode >= 0 && cfg->repr_mode <= 2, HB_ERR_ARG, "hb_create: bad repr_mode");
    HB_CHECK(cfg->n_na == 0 || cfg->na_inds, HB_ERR_ARG, "hb_create: n_na > 0 but na_inds == NULL");
    HB_CHECK(cfg->model <= 1, HB_ERR_ARG, "hb_create: model %u unknown (0 = BayesRRm, 1 = BayesW)", cfg->model);

    std::unique_ptr<hb_ctx> c(new (std::nothrow) hb_ctx);
    HB_CHECK(c, HB_ERR_NOMEM, "hb_create: out of host memory");
    c->cfg = *cfg;
    c->cfg.na_inds = nullptr; c->cfg.block_starts = nullptr; c->cfg.block_lens = nullptr;
    c->dev = cfg->device;
    HB_CUDA(cudaSetDevice(c->dev));
    cudaDeviceProp prop;
    HB_CUDA(cudaGetDeviceProperties(&prop, c->dev));
    HB_CHECK(prop.major >= 10, HB_ERR_CUDA, "hb_create: device %d is sm_%d%d; this library is built for sm_100a only", c->dev, prop.major, prop.minor);
    HB_CHECK(prop.cooperativeLaunch, HB_ERR_CUDA, "hb_create: device lacks cooperative launch");
    { const uint32_t norot = getenv("HB_NO_BED_ROT") ? 1u : 0u; HB_CUDA(cudaMemcpyToSymbol(g_bed_norot, &norot, sizeof(norot))); }   // developer knob
    c->n_sms = prop.multiProcessorCount;
    c->Nraw = cfg->n_ind_raw;
    c->N = cfg->n_ind_raw - cfg->n_na;
    c->Mtot = cfg->m_total; c->Ttot = cfg->n_tasks_total; c->T = cfg->n_tasks_local; c->t_first = cfg->task_first;
    c->K = cfg->n_mix; c->G = cfg->n_groups; c->SR = cfg->sync_rate == 0 ? 1u : cfg->sync_rate;

    // NA individuals (ascending, unique)
    c->na.assign(cfg->na_inds, cfg->na_inds + cfg->n_na);
    for (uint32_t i = 0; i < cfg->n_na; i++) {
        HB_CHECK(c->na[i] < c->Nraw && (i == 0 || c->na[i] > c->na[i - 1]), HB_ERR_ARG, "hb_create: na_inds must be ascending, unique, < n_ind_raw");
    }
    // task blocks
    if (cfg->block_starts && cfg->block_lens) {
        c->blkS.assign(cfg->block_starts, cfg->block_starts + c->Ttot);
        c->blkL.assign(cfg->block_lens, cfg->block_lens + c->Ttot);
        int64_t pos = 0;
        for (uint32_t t = 0; t < c->Ttot; t++) {
            HB_CHECK(c->blkS[t] == pos && c->blkL[t] >= 0, HB_ERR_ARG, "hb_create: marker blocks must be contiguous and ordered (task %u)", t);
            pos += c->blkL[t];
        }
        HB_CHECK(pos == (int64_t)c->Mtot, HB_ERR_ARG, "hb_create: marker blocks do not cover m_total");
    } else {
        define_blocks(c->Mtot, c->Ttot, c->blkS, c->blkL);
    }
    c->lmax = 0;
    for (uint32_t t = 0; t < c->Ttot; t++) c->lmax = std::max(c->lmax, (uint32_t)c->blkL[t]);  // global lmax (lock-step)
    c->m_start = (uint32_t)c->blkS[c->t_first];
    c->M = 0;
    for (uint32_t t = 0; t < c->T; t++) c->M += (uint32_t)c->blkL[c->t_first + t];
    HB_CHECK(c->M > 0, HB_ERR_ARG, "hb_create: no local markers");

    // slices: smallest S whose slice fits in shared memory next to the item table
    int max_smem = 0;
    HB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->dev));
    cudaFuncAttributes fa;
    HB_CUDA(cudaFuncGetAttributes(&fa, (const void *)k_brr_iteration));
    const size_t avail = (size_t)max_smem - fa.sharedSizeBytes;
    uint32_t max_ctas = cfg->max_ctas ? std::min<uint32_t>(cfg->max_ctas, c->n_sms) : (uint32_t)c->n_sms;
    auto slice_len = [&](uint32_t S) { return ((c->N + S - 1) / S + 63u) & ~63u; };
    auto fits = [&](uint32_t S) { const uint32_t L = slice_len(S); return L <= 65472u && smem_for(L) <= avail; };
    if (cfg->n_slices) {
        HB_CHECK(fits(cfg->n_slices), HB_ERR_ARG, "hb_create: n_slices=%u gives slices of %u individuals, too large for shared memory",
                 cfg->n_slices, slice_len(cfg->n_slices));
        c->S = cfg->n_slices;
    } else {
        uint32_t smin = 1;
        while (smin <= max_ctas && !fits(smin)) smin++;
        HB_CHECK(smin <= max_ctas, HB_ERR_ARG, "hb_create: %u individuals do not fit in the shared memory of %u CTAs", c->N, max_ctas);
        // among the slice counts near the minimum, take the one that leaves the fewest SMs idle
        uint32_t best = smin;
        for (uint32_t S = smin; S <= std::min(max_ctas, smin + smin / 4); S++)
            if (S * (max_ctas / S) > best * (max_ctas / best)) best = S;
        c->S = best;
        // BayesW keeps the residual in global memory (one CTA per slice in k_bw_update): twice the slices measured +6 % (150 -> 142 us per window)
        if (cfg->model == 1 && 2 * best <= max_ctas) c->S = 2 * best;
    }
    c->L = slice_len(c->S);
    HB_CHECK(c->S <= max_ctas, HB_ERR_ARG, "hb_create: %u slices > %u CTAs", c->S, max_ctas);
    c->R = max_ctas / c->S;
    c->smem_bytes = smem_for(c->L);
    HB_CUDA(cudaFuncSetAttribute((const void *)k_brr_iteration, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
    int occ = 0;
    HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)k_brr_iteration, kThreads, c->smem_bytes));
    HB_CHECK(occ >= 1, HB_ERR_CUDA, "hb_create: sampler kernel does not fit on an SM (smem %zu)", c->smem_bytes);
    // with small slices several CTAs may share an SM; co-residency only needs S*R <= occ*SMs (true: S*R <= SMs)

    HB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto &ev : c->ev) HB_CUDA(cudaEventCreate(&ev));

    // compacted individual -> raw individual
    if (c->na.size()) {
        std::vector<uint32_t> rmap(c->N);
        uint32_t k = 0, w = 0;
        for (uint32_t i = 0; i < c->Nraw; i++) {
            if (k < c->na.size() && c->na[k] == i) { k++; continue; }
            rmap[w++] = i;
        }
        HB_TRY(c->d_rmap.alloc(c->N));
        HB_CUDA(cudaMemcpy(c->d_rmap.p, rmap.data(), sizeof(uint32_t) * c->N, cudaMemcpyHostToDevice));
    }
    const size_t M = c->M;
    HB_TRY(c->d_rec.alloc(M)); HB_TRY(c->d_mave.alloc(M)); HB_TRY(c->d_mstd.alloc(M)); HB_TRY(c->d_grp.alloc(M));
    HB_CUDA(cudaMemset(c->d_grp.p, 0, sizeof(int32_t) * M));
    c->n1.assign(M, 0); c->n2.assign(M, 0); c->nm.assign(M, 0);
    c->is_bed.assign(M, 0); c->stored_bed.assign(M, 0); c->staged.assign(M, 0); c->rec_h.assign(M, 0); c->rec_bytes_h.assign(M, 0);
    HB_TRY(c->d_rec_bytes.alloc(M)); HB_TRY(c->d_err.alloc(1));
    HB_CUDA(cudaMemset(c->d_err.p, 0, sizeof(uint32_t)));
    HB_TRY(c->d_red.alloc(2 * ((size_t)cfg->n_groups * (1 + cfg->n_mix) + 8)));
    HB_TRY(c->d_E[0].alloc((size_t)c->S * c->L)); HB_TRY(c->d_E[1].alloc((size_t)c->S * c->L));
    HB_CUDA(cudaMemset(c->d_E[0].p, 0, sizeof(double) * c->S * c->L));
    HB_CUDA(cudaMemset(c->d_E[1].p, 0, sizeof(double) * c->S * c->L));
    HB_TRY(c->d_small.alloc(1 + 4 * (size_t)c->S + c->G));
    HB_CUDA(cudaMemset(c->d_small.p, 0, sizeof(double) * c->d_small.n));
    c->slice_sum_h.assign(c->S, 0.0);
    HB_TRY(c->d_beta.alloc(M)); HB_TRY(c->d_acum.alloc(M)); HB_TRY(c->d_comp.alloc(M));
    HB_CUDA(cudaMemset(c->d_beta.p, 0, sizeof(double) * M));
    HB_CUDA(cudaMemset(c->d_acum.p, 0, sizeof(double) * M));
    HB_CUDA(cudaMemset(c->d_comp.p, 0, sizeof(int32_t) * M));
    const size_t gk = (size_t)c->G * c->K;
    HB_TRY(c->d_cass.alloc(gk)); HB_TRY(c->d_hyp.alloc(4 * gk)); HB_TRY(c->d_active.alloc(c->G));
    HB_CUDA(cudaMemset(c->d_cass.p, 0, sizeof(int32_t) * gk));
    HB_CUDA(cudaMemset(c->d_hyp.p, 0, sizeof(double) * 4 * gk));
    HB_CUDA(cudaMemset(c->d_active.p, 0, c->G));
    HB_TRY(c->d_bar.alloc(1)); HB_TRY(c->d_stats.alloc(32));
    c->debug_cycles = getenv("HB_DEBUG_CYCLES") != nullptr;
    HB_TRY(c->d_ctacyc.alloc((size_t)c->S * c->R * 16));
    HB_CUDA(cudaMemset(c->d_stats.p, 0, 32 * sizeof(unsigned long long)));
    HB_TRY(c->d_order.alloc(1)); HB_TRY(c->d_u.alloc(1)); HB_TRY(c->d_z.alloc(1)); HB_TRY(c->d_num.alloc(1));
    HB_TRY(ensure_scratch(c.get(), 256));
    HB_TRY(ensure_pin(c.get(), 4096));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    *out = c.release();
    return HB_OK;
}

void hb_destroy(hb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->dev);
    cudaDeviceSynchronize();
    delete ctx;
}

int hb_get_layout(hb_ctx *c, uint32_t *n_ind, uint32_t *m_start, uint32_t *m_local, uint32_t *n_slices,
                  uint32_t *slice_len, uint32_t *n_groups_of_ctas, uint32_t *lmax) {
    HB_CHECK(c, HB_ERR_ARG, "null ctx");
    if (n_ind) *n_ind = c->N;
    if (m_start) *m_start = c->m_start;
    if (m_local) *m_local = c->M;
    if (n_slices) *n_slices = c->S;
    if (slice_len) *slice_len = c->L;
    if (n_groups_of_ctas) *n_groups_of_ctas = c->R;
    if (lmax) *lmax = c->lmax;
    return HB_OK;
}

int hb_get_task_blocks(hb_ctx *c, int32_t *starts, int32_t *lens) {
    HB_CHECK(c && starts && lens, HB_ERR_ARG, "null argument");
    std::copy(c->blkS.begin(), c->blkS.end(), starts);
    std::copy(c->blkL.begin(), c->blkL.end(), lens);
    return HB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------
// staging
// ------------------------------------------------------------------------------------
namespace hb {

static inline size_t raw_stride(uint32_t Nraw) { return ((size_t)(Nraw + 3) / 4 + 15) & ~(size_t)15; }

// d_raw holds n raw BED columns (stride raw_stride): build the records of local markers [m_first, m_first+n)
static int records_from_raw(hb_ctx *c, uint32_t m_first, uint32_t n) {
    const size_t stride = raw_stride(c->Nraw);
    const uint32_t S = c->S, L = c->L;
    HB_TRY(c->d_cnt3.ensure((size_t)n * S * 3));
    HB_TRY(c->d_start.ensure((size_t)n * S));
    HB_TRY(c->d_meta.ensure((size_t)n * 4));
    k_count<<<n, 256, 0, c->stream>>>(c->d_raw.p, stride, c->d_rmap.p, c->N, S, L, c->d_cnt3.p, c->d_start.p, c->d_meta.p);
    HB_CUDA(cudaGetLastError());
    std::vector<uint32_t> meta((size_t)n * 4);
    HB_CUDA(cudaMemcpyAsync(meta.data(), c->d_meta.p, sizeof(uint32_t) * 4 * n, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    size_t total = 0;
    std::vector<size_t> off(n);
    const size_t bed_bytes = (size_t)S * L / 4;
    const double bed_ratio = getenv("HB_BED_RATIO") ? atof(getenv("HB_BED_RATIO")) : 2.0;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t m = m_first + i;
        c->n1[m] = meta[i * 4 + 0]; c->n2[m] = meta[i * 4 + 1]; c->nm[m] = meta[i * 4 + 2];
        bool bed = false;
        if (c->cfg.repr_mode == HB_REPR_BED) bed = true;
        else if (c->cfg.repr_mode == HB_REPR_MIXED)  // src/data.cpp:931-932
            bed = ((double)(c->n1[m] + c->n2[m] + c->nm[m]) / (double)c->N) > c->cfg.threshold_fnz;
        c->is_bed[m] = bed ? 1 : 0;
        // Mixed representation: a marker the reference keeps as BED bytes is stored as 2-bit codes only where that saves
        // at least half of the bytes of the index form (16-bit slice-local indices: 2 B per non-zero against N/4 B, i.e.
        // above 25 % non-zeros; HB_BED_RATIO overrides the factor 2): the index form is ~4x faster per non-zero in the dot
        // phase. The kernel's sums are exact integers, so the results do not depend on the form of the record.
        const size_t sparse_bytes = dir_bytes(S) + 8 * (size_t)meta[i * 4 + 3];
        if (c->cfg.repr_mode == HB_REPR_MIXED && bed && (double)sparse_bytes <= bed_ratio * (double)bed_bytes) bed = false;
        c->stored_bed[m] = bed ? 1 : 0;
        const size_t bytes = bed ? bed_bytes : sparse_bytes;
        c->rec_bytes_h[m] = (uint32_t)((bytes + 15) & ~(size_t)15);
        off[i] = total;
        total += (bytes + 15) & ~(size_t)15;
    }
    void *arena = nullptr;
    cudaError_t e = cudaMalloc(&arena, std::max<size_t>(total, 16));
    HB_CHECK(e == cudaSuccess, HB_ERR_NOMEM, "genotype arena of %zu bytes: %s", total, cudaGetErrorString(e));
    c->arenas.push_back(arena);
    c->geno_bytes += total;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t m = m_first + i;
        c->rec_h[m] = (uint64_t)((uintptr_t)arena + off[i]) | (c->stored_bed[m] ? 1ull : 0ull);
        c->staged[m] = 1;
    }
    HB_CUDA(cudaMemcpyAsync(c->d_rec.p + m_first, c->rec_h.data() + m_first, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, c->stream));
    k_fill_sparse<<<n, 256, 0, c->stream>>>(c->d_raw.p, stride, c->d_rmap.p, c->N, S, L, c->d_cnt3.p, c->d_start.p, c->d_meta.p,
                                            c->d_rec.p + m_first);
    HB_CUDA(cudaGetLastError());
    if (!getenv("HB_NO_BANK_ORDER")) {
        const size_t bsm = (size_t)8 * 2 * kBankMax * sizeof(uint16_t);
        static bool attr_set = false;
        if (!attr_set) {
            HB_CUDA(cudaFuncSetAttribute((const void *)k_bank_order, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
            attr_set = true;
        }
        k_bank_order<<<n, 256, bsm, c->stream>>>(c->d_rec.p + m_first, S, L);
        HB_CUDA(cudaGetLastError());
    }
    dim3 g((S * L / 16 + 255) / 256, n);
    k_fill_bed<<<g, 256, 0, c->stream>>>(c->d_raw.p, stride, c->d_rmap.p, c->N, S, L, c->d_rec.p + m_first);
    HB_CUDA(cudaGetLastError());
    HB_CUDA(cudaStreamSynchronize(c->stream));
    c->finalized = false;
    return HB_OK;
}

static uint32_t stage_chunk(hb_ctx *c) {
    const size_t stride = raw_stride(c->Nraw);
    size_t ch = ((size_t)768 << 20) / stride;
    ch = std::max<size_t>(1, std::min<size_t>(ch, 32768));
    return (uint32_t)ch;
}

static int check_range(hb_ctx *c, uint32_t m_first, uint32_t n) {
    HB_CHECK(c, HB_ERR_ARG, "null ctx");
    HB_CHECK((uint64_t)m_first + n <= c->M, HB_ERR_ARG, "marker range [%u,+%u) outside the %u local markers", m_first, n, c->M);
    return HB_OK;
}

}  // namespace hb

extern "C" {

int hb_stage_bed(hb_ctx *c, uint32_t m_first, uint32_t n, const uint8_t *bed_cols) {
    HB_TRY(check_range(c, m_first, n));
    HB_CHECK(bed_cols || n == 0, HB_ERR_ARG, "hb_stage_bed: null data");
    HB_CUDA(cudaSetDevice(c->dev));
    const size_t nb = (size_t)(c->Nraw + 3) / 4, stride = raw_stride(c->Nraw);
    const uint32_t ch = stage_chunk(c);
    for (uint32_t o = 0; o < n; o += ch) {
        const uint32_t k = std::min(ch, n - o);
        HB_TRY(c->d_raw.ensure((size_t)k * stride));
        HB_CUDA(cudaMemcpy2DAsync(c->d_raw.p, stride, bed_cols + (size_t)o * nb, nb, nb, k, cudaMemcpyHostToDevice, c->stream));
        HB_TRY(records_from_raw(c, m_first + o, k));
    }
    return HB_OK;
}

int hb_stage_sparse(hb_ctx *c, uint32_t m_first, uint32_t n,
                    const uint32_t *I1, const uint64_t *N1S, const uint64_t *N1L,
                    const uint32_t *I2, const uint64_t *N2S, const uint64_t *N2L,
                    const uint32_t *IM, const uint64_t *NMS, const uint64_t *NML) {
    HB_TRY(check_range(c, m_first, n));
    HB_CHECK(N1S && N1L && N2S && N2L && NMS && NML, HB_ERR_ARG, "hb_stage_sparse: null start/length arrays");
    HB_CUDA(cudaSetDevice(c->dev));
    const size_t stride = raw_stride(c->Nraw);
    const uint32_t ch = stage_chunk(c);
    const uint32_t *I[3] = {I1, I2, IM};
    const uint64_t *NS[3] = {N1S, N2S, NMS}, *NL[3] = {N1L, N2L, NML};
    const uint32_t mask[3] = {1u, 3u, 2u};  // XOR masks of src/data.cpp:839-864
    DevBuf<uint32_t> d_I;
    DevBuf<uint64_t> d_S, d_L;
    for (uint32_t o = 0; o < n; o += ch) {
        const uint32_t k = std::min(ch, n - o);
        HB_TRY(c->d_raw.ensure((size_t)k * stride));
        HB_CUDA(cudaMemsetAsync(c->d_raw.p, 0xFF, (size_t)k * stride, c->stream));
        HB_TRY(d_S.ensure(k)); HB_TRY(d_L.ensure(k));
        for (int w = 0; w < 3; w++) {
            // the chunk's entries span [lo, hi) of the list array
            uint64_t lo = ~0ull, hi = 0, maxl = 0;
            for (uint32_t i = 0; i < k; i++) {
                const uint64_t s = NS[w][o + i], l = NL[w][o + i];
                if (l == 0) continue;
                lo = std::min(lo, s); hi = std::max(hi, s + l); maxl = std::max(maxl, l);
            }
            if (hi == 0) continue;
            HB_CHECK(I[w], HB_ERR_ARG, "hb_stage_sparse: null index array with non-empty lists");
            std::vector<uint64_t> rs(k);
            for (uint32_t i = 0; i < k; i++) rs[i] = NL[w][o + i] ? NS[w][o + i] - lo : 0;
            HB_TRY(d_I.ensure(hi - lo));
            HB_CUDA(cudaMemcpyAsync(d_I.p, I[w] + lo, sizeof(uint32_t) * (hi - lo), cudaMemcpyHostToDevice, c->stream));
            HB_CUDA(cudaMemcpyAsync(d_S.p, rs.data(), sizeof(uint64_t) * k, cudaMemcpyHostToDevice, c->stream));
            HB_CUDA(cudaMemcpyAsync(d_L.p, NL[w] + o, sizeof(uint64_t) * k, cudaMemcpyHostToDevice, c->stream));
            // index validity is checked on the host (cheap, and a bad index would corrupt device memory)
            for (uint64_t x = lo; x < hi; x++)
                HB_CHECK(I[w][x] < c->Nraw, HB_ERR_ARG, "hb_stage_sparse: index %u >= number of individuals %u", I[w][x], c->Nraw);
            dim3 g((unsigned)std::min<uint64_t>((maxl + 255) / 256, 64), k);
            k_lists_to_bed<<<g, 256, 0, c->stream>>>(c->d_raw.p, stride, d_I.p, d_S.p, d_L.p, mask[w]);
            HB_CUDA(cudaGetLastError());
            HB_CUDA(cudaStreamSynchronize(c->stream));  // rs / host arrays reused
        }
        HB_TRY(records_from_raw(c, m_first + o, k));
    }
    return HB_OK;
}

int hb_stage_synth(hb_ctx *c, uint32_t m_first, uint32_t n, uint32_t seed, const uint32_t *thresholds,
                   const uint32_t *attempts) {
    HB_TRY(check_range(c, m_first, n));
    HB_CHECK(thresholds, HB_ERR_ARG, "hb_stage_synth: null thresholds");
    HB_CUDA(cudaSetDevice(c->dev));
    const size_t stride = raw_stride(c->Nraw);
    const uint32_t nb = (c->Nraw + 3) / 4;
    const uint32_t ch = std::min<uint32_t>(stage_chunk(c), 65535u);
    DevBuf<uint32_t> d_thr, d_att;
    for (uint32_t o = 0; o < n; o += ch) {
        const uint32_t k = std::min(ch, n - o);
        HB_TRY(c->d_raw.ensure((size_t)k * stride));
        HB_TRY(d_thr.ensure((size_t)k * 3));
        HB_CUDA(cudaMemcpyAsync(d_thr.p, thresholds + (size_t)o * 3, sizeof(uint32_t) * 3 * k, cudaMemcpyHostToDevice, c->stream));
        if (attempts) {
            HB_TRY(d_att.ensure(k));
            HB_CUDA(cudaMemcpyAsync(d_att.p, attempts + o, sizeof(uint32_t) * k, cudaMemcpyHostToDevice, c->stream));
        }
        dim3 g(std::min<uint32_t>((nb + 255) / 256, 64), k);
        k_synth_bed<<<g, 256, 0, c->stream>>>(c->d_raw.p, stride, nb, c->Nraw, seed, c->m_start + m_first + o, d_thr.p,
                                              attempts ? d_att.p : nullptr);
        HB_CUDA(cudaGetLastError());
        HB_TRY(records_from_raw(c, m_first + o, k));
    }
    return HB_OK;
}

int hb_stage_finalize(hb_ctx *c) {
    HB_CHECK(c, HB_ERR_ARG, "null ctx");
    HB_CUDA(cudaSetDevice(c->dev));
    for (uint32_t m = 0; m < c->M; m++) HB_CHECK(c->staged[m], HB_ERR_STATE, "hb_stage_finalize: local marker %u was never staged", m);
    c->mave_h.resize(c->M); c->mstd_h.resize(c->M);
    const double dN = (double)c->N;
    for (uint32_t i = 0; i < c->M; i++) {  // src/BayesRRm.cpp:1502-1508
        const double a = (double)c->n1[i], b = (double)c->n2[i], nm = (double)c->nm[i];
        const double mave = (a + 2.0 * b) / (dN - nm);
        const double tmp1 = a * (1.0 - mave) * (1.0 - mave);
        const double tmp2 = b * (2.0 - mave) * (2.0 - mave);
        const double tmp0 = (double)(c->N - c->n1[i] - c->n2[i] - c->nm[i]) * (0.0 - mave) * (0.0 - mave);
        c->mave_h[i] = mave;
        c->mstd_h[i] = sqrt((double)(c->N - 1) / (tmp0 + tmp1 + tmp2));
        if (c->cfg.model == 1) {  // BayesW: mstd is the SD; the update multiplies by 1/mstd (src/BayesW.cpp:1218, 1506, 1619)
            c->sd_h.resize(c->M);
            c->sd_h[i] = sqrt((tmp0 + tmp1 + tmp2) / (double)(c->N - 1));
            c->mstd_h[i] = 1 / c->sd_h[i];
        }
    }
    if (c->cfg.model == 1) {
        HB_TRY(c->d_sd.alloc(c->M));
        HB_CUDA(cudaMemcpy(c->d_sd.p, c->sd_h.data(), sizeof(double) * c->M, cudaMemcpyHostToDevice));
    }
    HB_CUDA(cudaMemcpy(c->d_mave.p, c->mave_h.data(), sizeof(double) * c->M, cudaMemcpyHostToDevice));
    HB_CUDA(cudaMemcpy(c->d_mstd.p, c->mstd_h.data(), sizeof(double) * c->M, cudaMemcpyHostToDevice));
    HB_CUDA(cudaMemcpy(c->d_rec_bytes.p, c->rec_bytes_h.data(), sizeof(uint32_t) * c->M, cudaMemcpyHostToDevice));
    c->d_raw.release(); c->d_cnt3.release(); c->d_start.release(); c->d_meta.release();
    c->finalized = true;
    return HB_OK;
}

int hb_marker_counts(hb_ctx *c, uint32_t *n1, uint32_t *n2, uint32_t *nm) {
    HB_CHECK(c, HB_ERR_ARG, "null ctx");
    if (n1) std::copy(c->n1.begin(), c->n1.end(), n1);
    if (n2) std::copy(c->n2.begin(), c->n2.end(), n2);
    if (nm) std::copy(c->nm.begin(), c->nm.end(), nm);
    return HB_OK;
}

int hb_marker_stats(hb_ctx *c, double *mave, double *mstd) {
    HB_CHECK(c && c->finalized, HB_ERR_STATE, "hb_marker_stats: call hb_stage_finalize first");
    if (mave) std::copy(c->mave_h.begin(), c->mave_h.end(), mave);
    if (mstd) std::copy(c->mstd_h.begin(), c->mstd_h.end(), mstd);
    return HB_OK;
}

int hb_marker_is_bed(hb_ctx *c, uint8_t *flags) {
    HB_CHECK(c && flags, HB_ERR_ARG, "null argument");
    std::copy(c->is_bed.begin(), c->is_bed.end(), flags);
    return HB_OK;
}

uint64_t hb_genotype_bytes(hb_ctx *c) { return c ? c->geno_bytes : 0; }

int hb_export_bed(hb_ctx *c, uint32_t m, uint8_t *out) {
    HB_TRY(check_range(c, m, 1));
    HB_CHECK(out && c->staged[m], HB_ERR_STATE, "hb_export_bed: marker %u not staged", m);
    HB_CUDA(cudaSetDevice(c->dev));
    const size_t ostride = (size_t)c->S * c->L / 4;
    DevBuf<uint8_t> d;
    HB_TRY(d.alloc(ostride));
    HB_CUDA(cudaMemsetAsync(d.p, 0xFF, ostride, c->stream));
    k_record_to_bed<<<1, 256, 0, c->stream>>>(c->d_rec.p, m, c->S, c->L, d.p, ostride);
    HB_CUDA(cudaGetLastError());
    const size_t nb = (size_t)(c->N + 3) / 4;
    std::vector<uint8_t> tmp(ostride);
    HB_CUDA(cudaMemcpyAsync(tmp.data(), d.p, ostride, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(out, tmp.data(), nb);
    // PLINK pads the last byte with 00; internal pad positions hold 11
    if (c->N % 4) out[nb - 1] &= (uint8_t)((1u << (2 * (c->N % 4))) - 1u);
    return HB_OK;
}

int hb_export_sparse(hb_ctx *c, uint32_t m_first, uint32_t n,
                     uint32_t *I1, uint64_t *N1S, uint64_t *N1L,
                     uint32_t *I2, uint64_t *N2S, uint64_t *N2L,
                     uint32_t *IM, uint64_t *NMS, uint64_t *NML) {
    HB_TRY(check_range(c, m_first, n));
    HB_CHECK(N1S && N1L && N2S && N2L && NMS && NML, HB_ERR_ARG, "hb_export_sparse: null start/length arrays");
    HB_CUDA(cudaSetDevice(c->dev));
    uint64_t t1 = 0, t2 = 0, tm = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t m = m_first + i;
        HB_CHECK(c->staged[m], HB_ERR_STATE, "hb_export_sparse: marker %u not staged", m);
        N1S[i] = t1; N1L[i] = c->n1[m]; t1 += c->n1[m];
        N2S[i] = t2; N2L[i] = c->n2[m]; t2 += c->n2[m];
        NMS[i] = tm; NML[i] = c->nm[m]; tm += c->nm[m];
    }
    const size_t ostride = (size_t)c->S * c->L / 4;
    size_t chm = std::max<size_t>(1, ((size_t)256 << 20) / ostride);
    DevBuf<uint8_t> d;
    DevBuf<uint32_t> dI[3];
    DevBuf<uint64_t> dS[3];
    uint32_t *outI[3] = {I1, I2, IM};
    const uint64_t *NS[3] = {N1S, N2S, NMS}, *NL[3] = {N1L, N2L, NML};
    for (uint32_t o = 0; o < n; o += (uint32_t)chm) {
        const uint32_t k = (uint32_t)std::min<size_t>(chm, n - o);
        HB_TRY(d.ensure((size_t)k * ostride));
        HB_CUDA(cudaMemsetAsync(d.p, 0xFF, (size_t)k * ostride, c->stream));
        k_record_to_bed<<<k, 256, 0, c->stream>>>(c->d_rec.p, m_first + o, c->S, c->L, d.p, ostride);
        HB_CUDA(cudaGetLastError());
        uint64_t base[3], cntw[3];
        std::vector<uint64_t> rs[3];
        for (int w = 0; w < 3; w++) {
            base[w] = NS[w][o];
            cntw[w] = NS[w][o + k - 1] + NL[w][o + k - 1] - base[w];
            rs[w].resize(k);
            for (uint32_t i = 0; i < k; i++) rs[w][i] = NS[w][o + i] - base[w];
            HB_TRY(dI[w].ensure(std::max<uint64_t>(cntw[w], 1)));
            HB_TRY(dS[w].ensure(k));
            HB_CUDA(cudaMemcpyAsync(dS[w].p, rs[w].data(), sizeof(uint64_t) * k, cudaMemcpyHostToDevice, c->stream));
        }
        k_bed_to_lists<<<k, 256, 0, c->stream>>>(d.p, ostride, c->N, dI[0].p, dS[0].p, dI[1].p, dS[1].p, dI[2].p, dS[2].p);
        HB_CUDA(cudaGetLastError());
        for (int w = 0; w < 3; w++) {
            if (cntw[w]) {
                HB_CHECK(outI[w], HB_ERR_ARG, "hb_export_sparse: null index output");
                HB_CUDA(cudaMemcpyAsync(outI[w] + base[w], dI[w].p, sizeof(uint32_t) * cntw[w], cudaMemcpyDeviceToHost, c->stream));
            }
        }
        HB_CUDA(cudaStreamSynchronize(c->stream));
    }
    return HB_OK;
}

// ------------------------------------------------------------------------------------
// epsilon + unit-level kernels
// ------------------------------------------------------------------------------------
int hb_set_epsilon(hb_ctx *c, const double *eps) {
    HB_CHECK(c && eps, HB_ERR_ARG, "null argument");
    HB_CUDA(cudaSetDevice(c->dev));
    HB_CUDA(cudaMemsetAsync(c->d_E[c->cur].p, 0, sizeof(double) * c->S * c->L, c->stream));
    HB_CUDA(cudaMemcpyAsync(c->d_E[c->cur].p, eps, sizeof(double) * c->N, cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    c->shift = 0.0;
    c->eps_set = true;
    c->eps_slice_abs = 0.0; c->eps_abs_max = 0.0;
    for (uint32_t s0 = 0; s0 < c->N; s0 += c->L) {
        double a = 0.0;
        for (uint32_t i = s0; i < std::min(c->N, s0 + c->L); i++) { a += fabs(eps[i]); c->eps_abs_max = std::max(c->eps_abs_max, fabs(eps[i])); }
        c->eps_slice_abs = std::max(c->eps_slice_abs, a);
    }
    return HB_OK;
}

int hb_get_epsilon(hb_ctx *c, double *eps) {
    HB_CHECK(c && eps, HB_ERR_ARG, "null argument");
    HB_CUDA(cudaSetDevice(c->dev));
    HB_CUDA(cudaMemcpyAsync(eps, c->d_E[c->cur].p, sizeof(double) * c->N, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->shift != 0.0)
        for (uint32_t i = 0; i < c->N; i++) eps[i] += c->shift;
    return HB_OK;
}

static int upload_markers(hb_ctx *c, const uint32_t *markers, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) HB_CHECK(markers[i] < c->M, HB_ERR_ARG, "marker %u out of range", markers[i]);
    HB_TRY(c->d_order.ensure(n));
    HB_TRY(c->d_u.ensure(n));  // the item table reads u/z of every position, also in the unit modes
    HB_TRY(c->d_z.ensure(n));
    HB_CUDA(cudaMemcpyAsync(c->d_order.p, markers, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
    return HB_OK;
}

int hb_dot_markers(hb_ctx *c, const uint32_t *markers, uint32_t n, double *num) {
    HB_CHECK(c && markers && num, HB_ERR_ARG, "null argument");
    HB_CHECK(c->finalized, HB_ERR_STATE, "hb_dot_markers: call hb_stage_finalize first");
    HB_CUDA(cudaSetDevice(c->dev));
    const uint32_t chunk = 1u << 16;
    for (uint32_t o = 0; o < n; o += chunk) {
        const uint32_t k = std::min(chunk, n - o);
        HB_TRY(upload_markers(c, markers + o, k));
        HB_TRY(ensure_scratch(c, k));
        HB_TRY(c->d_num.ensure(k));
        BrrParams P;
        fill_params(c, P);
        P.mode = MODE_DOT; P.T = 1; P.lmax = k; P.SR = 1;
        HB_TRY(launch_window_kernel(c, P));
        HB_CUDA(cudaMemcpyAsync(num + o, c->d_num.p, sizeof(double) * k, cudaMemcpyDeviceToHost, c->stream));
        HB_CUDA(cudaStreamSynchronize(c->stream));
    }
    return HB_OK;
}

int hb_scaadd_markers(hb_ctx *c, const uint32_t *markers, const double *dbeta, uint32_t n) {
    HB_CHECK(c && markers && dbeta, HB_ERR_ARG, "null argument");
    HB_CHECK(c->finalized, HB_ERR_STATE, "hb_scaadd_markers: call hb_stage_finalize first");
    if (n == 0) return HB_OK;
    HB_CUDA(cudaSetDevice(c->dev));
    HB_TRY(upload_markers(c, markers, n));
    HB_TRY(ensure_scratch(c, n));
    std::vector<double> dbs(n);
    for (uint32_t i = 0; i < n; i++) dbs[i] = dbeta[i] * c->mstd_h[markers[i]];  // mstd*deltaBeta (:1982)
    HB_CUDA(cudaMemcpyAsync(c->d_dB.p, dbs.data(), sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    BrrParams P;
    fill_params(c, P);
    P.mode = MODE_SCAADD; P.T = 1; P.lmax = n; P.SR = 1;
    P.shift_in = 0.0;
    {   // the fixed-point grid must hold the updated residual: bound it by the L1 norm / twice the sum of the updates
        double add_abs = 0.0, add_max = 0.0;
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t m = markers[i];
            add_abs += fabs(dbs[i]) * ((double)c->n1[m] + 2.0 * (double)c->n2[m] + fabs(c->mave_h[m]) * (double)c->nm[m]);
            add_max += 2.0 * fabs(dbs[i]);
        }
        c->eps_slice_abs += add_abs;
        c->eps_abs_max += add_max;
    }
    HB_TRY(launch_window_kernel(c, P));
    double off = 0.0;
    std::vector<double> am(2 * (size_t)c->S);
    HB_CUDA(cudaMemcpyAsync(&off, c->d_small.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaMemcpyAsync(am.data(), c->d_small.p + 1 + 2 * c->S + c->G, sizeof(double) * 2 * c->S, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    HB_TRY(check_kernel_error(c, "hb_scaadd_markers"));
    c->eps_slice_abs = *std::max_element(am.begin(), am.begin() + c->S);
    c->eps_abs_max = *std::max_element(am.begin() + c->S, am.end());
    c->cur ^= 1;
    c->shift += off;
    return HB_OK;
}

// ------------------------------------------------------------------------------------
// BayesRRm chain
// ------------------------------------------------------------------------------------
int hb_brr_init(hb_ctx *c, const double *y, const int32_t *groups, const double *mS, const double *sigmaG0, uint32_t seed) {
    HB_CHECK(c && y && mS, HB_ERR_ARG, "null argument");
    HB_CHECK(c->finalized, HB_ERR_STATE, "hb_brr_init: call hb_stage_finalize first");
    HB_CUDA(cudaSetDevice(c->dev));
    if (c->prefetch.joinable()) c->prefetch.join();  // draws prefetched for a previous chain are void
    c->have_next = false;
    const uint32_t N = c->N, K = c->K, G = c->G;
    // groups (global array of m_total; NULL = single group)
    c->MtotGrp.assign(G, 0);
    c->groups_local.assign(c->M, 0);
    if (groups) {
        for (uint32_t m = 0; m < c->Mtot; m++) {
            HB_CHECK(groups[m] >= 0 && (uint32_t)groups[m] < G, HB_ERR_ARG, "hb_brr_init: group %d of marker %u out of range", groups[m], m);
            c->MtotGrp[groups[m]]++;
        }
        for (uint32_t m = 0; m < c->M; m++) c->groups_local[m] = groups[c->m_start + m];
    } else {
        HB_CHECK(G == 1, HB_ERR_ARG, "hb_brr_init: groups == NULL needs n_groups == 1");
        c->MtotGrp[0] = (int32_t)c->Mtot;
    }
    HB_CUDA(cudaMemcpy(c->d_grp.p, c->groups_local.data(), sizeof(int32_t) * c->M, cudaMemcpyHostToDevice));
    // mixture variances and prior pi (src/BayesRRm.cpp:1097-1110)
    c->cVa.assign((size_t)G * K, 0.0); c->cVaI.assign((size_t)G * K, 0.0); c->pi.assign((size_t)G * K, 0.0);
    for (uint32_t g = 0; g < G; g++) {
        double s = 0.0;
        for (uint32_t k = 1; k < K; k++) {
            const double v = mS[g * K + k];
            HB_CHECK(v > 0.0, HB_ERR_ARG, "hb_brr_init: mixture variance must be > 0 (group %u, k %u)", g, k);
            c->cVa[g * K + k] = v; c->cVaI[g * K + k] = 1.0 / v; s += v;
        }
        c->pi[g * K] = 0.5;
        for (uint32_t k = 1; k < K; k++) c->pi[g * K + k] = 0.5 * c->cVa[g * K + k] / s;
    }
    // RNG spec v1 streams (seeding rule of :1228: seed + 1000*rank)
    c->seed = seed;
    c->task_rng.clear();
    for (uint32_t t = 0; t < c->T; t++) c->task_rng.emplace_back(seed + 1000u * (c->t_first + t));
    c->hyper_rng.seed(seed ^ 0x5bd1e995u);
    c->sigmaG.assign(G, 0.0);
    for (uint32_t g = 0; g < G; g++) {
        c->sigmaG[g] = sigmaG0 ? sigmaG0[g] : c->hyper_rng.res53();  // beta_rng(1,1) == U(0,1)  (:1233)
        if (c->MtotGrp[g] == 0) c->sigmaG[g] = 0.0;  // :1239-1240
    }
    c->active.assign(G, 0);
    for (uint32_t g = 0; g < G; g++) c->active[g] = (c->sigmaG[g] != 0.0) ? 1 : 0;  // adaV, :1592-1597
    // y -> centred, scaled (:371-388); eps = y; sigmaE = sum(eps^2)/N*0.5 (:1575-1579)
    std::vector<double> e(y, y + N);
    double mean = 0.0;
    for (uint32_t i = 0; i < N; i++) mean += e[i];
    mean /= (double)(int)N;
    for (uint32_t i = 0; i < N; i++) e[i] -= mean;
    double sqn = 0.0;
    for (uint32_t i = 0; i < N; i++) sqn += e[i] * e[i];
    HB_CHECK(sqn > 0.0, HB_ERR_ARG, "hb_brr_init: phenotype has zero variance");
    sqn = sqrt((double)(N - 1) / sqn);
    for (uint32_t i = 0; i < N; i++) e[i] *= sqn;
    double s2 = 0.0, s1 = 0.0;
    for (uint32_t i = 0; i < N; i++) { s2 += e[i] * e[i]; s1 += e[i]; }
    c->sigmaE = s2 / (double)N * 0.5;
    HB_TRY(hb_set_epsilon(c, e.data()));
    // slice sums of the initial residual
    for (uint32_t s = 0; s < c->S; s++) {
        double v = 0.0;
        for (uint32_t i = s * c->L; i < std::min(N, (s + 1) * c->L); i++) v += e[i];
        c->slice_sum_h[s] = v;
    }
    c->mu.assign(c->T, 0.0);
    c->bsq.assign(G, 0.0); c->cass.assign((size_t)G * K, 0); c->m0.assign(G, 0);
    HB_CUDA(cudaMemset(c->d_beta.p, 0, sizeof(double) * c->M));
    HB_CUDA(cudaMemset(c->d_acum.p, 0, sizeof(double) * c->M));
    HB_CUDA(cudaMemset(c->d_comp.p, 0, sizeof(int32_t) * c->M));
    // identity marker order per task (markerI, :1615-1617)
    for (void *&q : c->pinned_perm) if (q) { cudaHostUnregister(q); q = nullptr; }
    c->perm.resize(c->M);
    c->perm_next.resize(c->M);
    if (cudaHostRegister(c->perm.data(), sizeof(int32_t) * c->M, cudaHostRegisterDefault) == cudaSuccess) c->pinned_perm[0] = c->perm.data();
    if (cudaHostRegister(c->perm_next.data(), sizeof(int32_t) * c->M, cudaHostRegisterDefault) == cudaSuccess) c->pinned_perm[1] = c->perm_next.data();
    cudaGetLastError();  // registration is an optimisation only
    {
        size_t o = 0;
        for (uint32_t t = 0; t < c->T; t++)
            for (int32_t j = 0; j < c->blkL[c->t_first + t]; j++) c->perm[o++] = j;
    }
    std::vector<int32_t> tl(c->T), to(c->T);
    for (uint32_t t = 0; t < c->T; t++) { tl[t] = c->blkL[c->t_first + t]; to[t] = c->blkS[c->t_first + t] - (int32_t)c->m_start; }
    HB_TRY(c->d_task_len.alloc(c->T)); HB_TRY(c->d_task_off.alloc(c->T));
    HB_CUDA(cudaMemcpy(c->d_task_len.p, tl.data(), sizeof(int32_t) * c->T, cudaMemcpyHostToDevice));
    HB_CUDA(cudaMemcpy(c->d_task_off.p, to.data(), sizeof(int32_t) * c->T, cudaMemcpyHostToDevice));
    const size_t Q = (size_t)c->lmax * c->T;
    HB_TRY(c->d_order.ensure(Q)); HB_TRY(c->d_u.ensure(Q)); HB_TRY(c->d_z.ensure(Q));
    HB_TRY(c->d_wmeta.ensure(Q)); HB_TRY(c->d_dirw.ensure((size_t)c->S * Q));
    c->balance = !getenv("HB_NO_BALANCE");
    {
        std::vector<uint32_t> w(c->M);
        for (uint32_t m = 0; m < c->M; m++) w[m] = c->stored_bed[m] ? c->N / 2 : (c->n1[m] + c->n2[m] + c->nm[m]);
        HB_TRY(c->d_wts.alloc(c->M));
        HB_CUDA(cudaMemcpy(c->d_wts.p, w.data(), sizeof(uint32_t) * c->M, cudaMemcpyHostToDevice));
    }
    HB_TRY(c->d_perm.ensure(c->M)); HB_TRY(c->d_ut.ensure(c->M)); HB_TRY(c->d_zt.ensure(c->M));
    // windows run ahead (brr_kernel.cuh): sync_rate steps, or enough steps for ~32 window positions where the tasks are few
    c->n_ahead = std::min<uint32_t>(kSpecMax, std::max<uint32_t>(c->SR, (32u + c->T - 1) / c->T));
    if (getenv("HB_N_AHEAD")) c->n_ahead = std::min<uint32_t>(kSpecMax, std::max<uint32_t>(c->SR, (uint32_t)atoi(getenv("HB_N_AHEAD"))));
    HB_TRY(ensure_scratch(c, std::max(c->SR, c->n_ahead) * c->T));
    HB_TRY(ensure_pin(c, std::max<size_t>(4 * (size_t)G * K, 16 + 4 * (size_t)c->S + G + (size_t)G * K + 32)));  // hyp tables (4*G*K) and the per-iteration read-back share it
    c->iteration = 0;
    c->order_ready = false;
    c->brr_ready = true;
    c->F = 0; c->gamma.clear(); c->xI.clear();   // fixed effects are attached after the init (hb_brr_set_covariates)
    c->fh_on = false; c->grp_priors.clear(); c->dirichlet.clear();   // so are the prior files and bayesFH (hb_brr_set_group_priors / hb_brr_set_fh)
    { const uint64_t sv = seed; HB_TRY(check_equal_over_ranks(c, &sv, 1, "the seed (give every process the same --seed)")); }
    return HB_OK;
}

int hb_brr_iteration(hb_ctx *c, const hb_brr_tape *tape, hb_brr_iter_out *out) {
    HB_CHECK(c && c->brr_ready, HB_ERR_STATE, "hb_brr_iteration: call hb_brr_init first");
    HB_CUDA(cudaSetDevice(c->dev));
    if (tape) HB_CHECK(tape->zmu && tape->perm && tape->u && tape->z, HB_ERR_ARG, "hb_brr_iteration: tape needs zmu, perm, u and z");
    const uint32_t N = c->N, K = c->K, G = c->G, T = c->T, M = c->M;
    const double dN = (double)N, dNm1 = (double)(N - 1);
    const double v0E = 0.0001, s02E = 0.0001;  // src/BayesRRm.h:30-33
    double v0G = 0.0001, s02G = 0.0001;        // overwritten group by group where a priors file was given (:2545-2548)
    cudaStream_t st = c->stream;
    HB_CUDA(cudaEventRecord(c->ev[0], st));

    // The task streams' draws of an iteration (z for mu :1682, then the shuffle :1692) do not depend on the data, so the
    // next iteration's draws are produced on a worker thread while the GPU runs the marker loop.
    if (c->prefetch.joinable()) c->prefetch.join();
    std::vector<double> zmu(T, 0.0);
    bool drew_ahead = false;   // this iteration uses the order that was drawn (and possibly laid out on the device) ahead
    if (!tape) {
        if (!c->have_next) {  // first iteration (or after a taped one): draw now
            for (uint32_t t = 0; t < T; t++) {
                zmu[t] = c->task_rng[t].normal();
                if (c->cfg.reserved[0] == 0) {  // reserved[0]: --shuf-mark 0
                    size_t o = 0;
                    for (uint32_t tt = 0; tt < t; tt++) o += (size_t)c->blkL[c->t_first + tt];
                    c->task_rng[t].shuffle(c->perm.data() + o, c->blkL[c->t_first + t]);
                }
            }
        } else {
            zmu = c->zmu_next;
            c->perm.swap(c->perm_next);
            drew_ahead = true;
        }
        c->have_next = false;
    }
    // ---- mu (:1675-1686): eps_r + mu_r is the same vector for every task
    double ssum = 0.0;
    for (uint32_t s = 0; s < c->S; s++) ssum += c->slice_sum_h[s];
    const double epssum = ssum + dN * c->shift + dN * c->mu[0];
    const double mu0_old = c->mu[0];
    for (uint32_t t = 0; t < T; t++) {
        const double z = tape ? tape->zmu[t] : zmu[t];
        c->mu[t] = epssum / dN + sqrt(c->sigmaE / dN) * z;
    }
    c->shift += mu0_old - c->mu[0];

    // ---- marker order (:1691-1694) and per-marker draws, laid out in window order q = j*T + t
    if (tape) {
        size_t o = 0;
        for (uint32_t t = 0; t < T; t++) {
            const int32_t len = c->blkL[c->t_first + t];
            for (int32_t j = 0; j < len; j++, o++) {
                HB_CHECK(tape->perm[o] >= 0 && tape->perm[o] < len, HB_ERR_ARG, "hb_brr_iteration: tape perm[%zu]=%d out of range", o, tape->perm[o]);
                c->perm[o] = tape->perm[o];
            }
        }
    }
    const uint32_t Q = c->lmax * T;
    // (the order of this iteration may already be on the device: launched at the end of the previous call, while the host drew
    // the hyper-parameters)
    const bool reuse_order = c->order_ready && !tape && drew_ahead;
    c->order_ready = false;
    if (!reuse_order) {
        HB_CUDA(cudaMemcpyAsync(c->d_perm.p, c->perm.data(), sizeof(int32_t) * M, cudaMemcpyHostToDevice, st));
        if (tape) {
            HB_CUDA(cudaMemcpyAsync(c->d_ut.p, tape->u, sizeof(double) * M, cudaMemcpyHostToDevice, st));
            HB_CUDA(cudaMemcpyAsync(c->d_zt.p, tape->z, sizeof(double) * M, cudaMemcpyHostToDevice, st));
        }
        k_window_order<<<c->lmax, std::min(256u, (T + 31u) & ~31u), 0, st>>>(c->d_perm.p, tape ? c->d_ut.p : nullptr, tape ? c->d_zt.p : nullptr,
                                                         c->d_task_len.p, c->d_task_off.p, T, c->lmax, c->seed, c->iteration,
                                                         c->t_first, c->d_order.p, c->d_u.p, c->d_z.p,
                                                         c->d_rec.p, c->d_mave.p, c->d_mstd.p, c->d_beta.p, c->d_grp.p, c->d_rec_bytes.p,
                                                         c->balance ? c->d_wts.p : nullptr, c->S, c->d_wmeta.p, c->d_dirw.p);
        HB_CUDA(cudaGetLastError());
    }

    // ---- hyper-parameter tables (:1721-1723, 1750, 1863-1876, 1901)
    const size_t gk = (size_t)G * K;
    std::vector<double> hyp(4 * gk, 0.0);
    for (uint32_t g = 0; g < G; g++) {
        const double sigE_G = c->sigmaE / c->sigmaG[g], sigG_E = c->sigmaG[g] / c->sigmaE;
        for (uint32_t k = 0; k < K; k++) {
            hyp[g * K + k] = log(c->pi[g * K + k]);
            if (k == 0 || !c->active[g] || c->fh_on) continue;   // (FH: per-marker values from k_fh_prepare)
            hyp[gk + g * K + k] = 0.5 * log(sigG_E * dNm1 * c->cVa[g * K + k] + 1.0);
            const double den = dNm1 + sigE_G * c->cVaI[g * K + k];
            hyp[2 * gk + g * K + k] = den;
            hyp[3 * gk + g * K + k] = sqrt(c->sigmaE / den);
        }
    }
    memcpy(c->pin, hyp.data(), sizeof(double) * 4 * gk);
    HB_CUDA(cudaMemcpyAsync(c->d_hyp.p, c->pin, sizeof(double) * 4 * gk, cudaMemcpyHostToDevice, st));
    HB_CUDA(cudaMemcpyAsync(c->d_active.p, c->active.data(), G, cudaMemcpyHostToDevice, st));
    HB_CUDA(cudaMemsetAsync(c->d_cass.p, 0, sizeof(int32_t) * gk, st));  // :1697
    HB_CUDA(cudaMemsetAsync(c->d_stats.p, 0, sizeof(unsigned long long) * 32, st));

    // ---- bayesFHMPI: nu_var and the marker's own prior variance (:1727-1731), for all markers at once
    FhParams FQ{};
    int fh_launches = 0;
    if (c->fh_on) {
        if (tape) HB_CHECK(tape->gnu && tape->glam, HB_ERR_ARG, "hb_brr_iteration: a tape for bayesFH needs gnu and glam");
        FQ.M = M; FQ.m_start = c->m_start; FQ.seed = c->seed; FQ.iteration = c->iteration;
        FQ.shape = 0.5 + 0.5 * c->fh_cfg.v0L; FQ.v0L = c->fh_cfg.v0L; FQ.tau = c->fh_tau; FQ.sigmaE = c->sigmaE; FQ.dNm1 = dNm1;
        FQ.c_slab = c->d_fh_c.p; FQ.grp = c->d_grp.p; FQ.beta = c->d_beta.p;
        FQ.lambda = c->d_fh_lambda.p; FQ.nu = c->d_fh_nu.p; FQ.par = c->d_fh_par.p; FQ.part = c->d_fh_part.p;
        HB_CUDA(cudaMemcpyAsync(c->d_fh_c.p, c->fh_c.data(), sizeof(double) * G, cudaMemcpyHostToDevice, st));
        if (tape) {
            HB_CUDA(cudaMemcpyAsync(c->d_fh_g.p, tape->gnu, sizeof(double) * M, cudaMemcpyHostToDevice, st));
            HB_CUDA(cudaMemcpyAsync(c->d_fh_g.p + M, tape->glam, sizeof(double) * M, cudaMemcpyHostToDevice, st));
        }
        FQ.g_tape = tape ? c->d_fh_g.p : nullptr;
        k_fh_prepare<<<std::max(1u, std::min(4u * 148u, (M + 255u) / 256u)), 256, 0, st>>>(FQ);
        HB_CUDA(cudaGetLastError());
        fh_launches = 1;
    }

    // ---- marker loop
    BrrParams P;
    fill_params(c, P);
    P.mode = MODE_CHAIN; P.T = T; P.SR = c->SR; P.lmax = c->lmax; P.n_ahead = c->n_ahead;
    if (c->fh_on) P.fh = c->d_fh_par.p;
    P.meta = c->d_wmeta.p; P.dirw = c->d_dirw.p;
    P.shift_in = c->shift;  // fold the accumulated constant into the stored residual
    P.i_2sigE = 1.0 / (2.0 * c->sigmaE);
    HB_CUDA(cudaEventRecord(c->ev[1], st));
    HB_TRY(launch_window_kernel(c, P));
    HB_CUDA(cudaEventRecord(c->ev[2], st));
    c->cur ^= 1;
    if (!tape) {  // overlap the next iteration's host draws with the marker loop
        c->perm_next = c->perm;
        c->zmu_next.assign(T, 0.0);
        hb_ctx *cc = c;
        c->prefetch = std::thread([cc, T]() {
            size_t o = 0;
            for (uint32_t t = 0; t < T; t++) {
                cc->zmu_next[t] = cc->task_rng[t].normal();
                const int32_t len = cc->blkL[cc->t_first + t];
                if (cc->cfg.reserved[0] == 0) cc->task_rng[t].shuffle(cc->perm_next.data() + o, len);
                o += (size_t)len;
            }
        });
        c->have_next = true;
    }

    // ---- group statistics (:2496-2521)
    double *d_bsq = c->d_small.p + 1 + 2 * c->S;
    HB_TRY(c->d_bsq_part.ensure((size_t)G * kSqChunks));
    k_beta_sqnorm<<<dim3(kSqChunks, G), 256, 0, st>>>(c->d_beta.p, c->d_grp.p, M, c->d_bsq_part.p, c->d_comp.p, c->d_active.p, K, c->d_cass.p);
    k_beta_sqnorm_fin<<<(G + 127) / 128, 128, 0, st>>>(c->d_bsq_part.p, kSqChunks, G, d_bsq);
    HB_CUDA(cudaGetLastError());
    const size_t nsmall = 1 + 4 * (size_t)c->S + G;
    double *pin_small = c->pin;
    int32_t *pin_cass = reinterpret_cast<int32_t *>(c->pin + nsmall);
    unsigned long long *pin_stats = reinterpret_cast<unsigned long long *>(c->pin + nsmall + (gk + 1) / 2 + 1);
    double *pin_fh = c->pin + nsmall + (gk + 1) / 2 + 1 + 32;
    if (c->fh_on) {   // lambda_var from the final effects (:1952) and the scaled sum of squares (:2506-2509)
        FQ.g_tape = tape ? c->d_fh_g.p + M : nullptr;
        k_fh_finish<<<kSqChunks, 256, 0, st>>>(FQ);
        k_beta_sqnorm_fin<<<1, 32, 0, st>>>(c->d_fh_part.p, kSqChunks, 1, c->d_fh_part.p + kSqChunks);
        HB_CUDA(cudaGetLastError());
        HB_CUDA(cudaMemcpyAsync(pin_fh, c->d_fh_part.p + kSqChunks, sizeof(double), cudaMemcpyDeviceToHost, st));
        fh_launches += 2;
    }
    HB_CUDA(cudaMemcpyAsync(pin_small, c->d_small.p, sizeof(double) * nsmall, cudaMemcpyDeviceToHost, st));
    HB_CUDA(cudaMemcpyAsync(pin_cass, c->d_cass.p, sizeof(int32_t) * gk, cudaMemcpyDeviceToHost, st));
    HB_CUDA(cudaMemcpyAsync(pin_stats, c->d_stats.p, sizeof(unsigned long long) * 32, cudaMemcpyDeviceToHost, st));
    HB_CUDA(cudaEventRecord(c->ev[3], st));
    // The next iteration's order is known (worker thread above) and nothing it needs changes any more (effects, records): its
    // window layout runs on the device now, while the host waits for the statistics and draws the hyper-parameters.
    bool order_ahead = false;
    if (!tape && c->have_next && !c->debug_cycles && !getenv("HB_NO_ORDER_AHEAD")) {
        if (c->prefetch.joinable()) c->prefetch.join();
        HB_CUDA(cudaMemcpyAsync(c->d_perm.p, c->perm_next.data(), sizeof(int32_t) * M, cudaMemcpyHostToDevice, st));
        k_window_order<<<c->lmax, std::min(256u, (T + 31u) & ~31u), 0, st>>>(c->d_perm.p, nullptr, nullptr, c->d_task_len.p, c->d_task_off.p, T, c->lmax,
                                                         c->seed, c->iteration + 1, c->t_first, c->d_order.p, c->d_u.p, c->d_z.p,
                                                         c->d_rec.p, c->d_mave.p, c->d_mstd.p, c->d_beta.p, c->d_grp.p, c->d_rec_bytes.p,
                                                         c->balance ? c->d_wts.p : nullptr, c->S, c->d_wmeta.p, c->d_dirw.p);
        HB_CUDA(cudaGetLastError());
        order_ahead = true;
    }
    HB_CUDA(cudaEventSynchronize(c->ev[3]));

    if (c->debug_cycles) {  // developer aid: spread of the per-CTA phase cycles
        fprintf(stderr, "[hb] fixed-point grid 2^-%d\n", c->last_sh);
        {
            const double nwin = (double)std::max<unsigned long long>(1, pin_stats[1]);
            const char *nm8[8] = {"upd:count", "upd:stage", "upd:apply", "xchg:push", "xchg:wait", "draw:poll", "draw:math", "draw:store"};
            const char *nmp[8] = {"table", "dot", "publish+draw", "barrier", "update(local)", "sums", "-", "update(peers)"};
            fprintf(stderr, "[hb] CTA 0 thread 0 cycles per window, phases:");
            for (int i = 0; i < 8; i++) fprintf(stderr, " %s=%.0f", nmp[i], (double)pin_stats[8 + i] / nwin);
            fprintf(stderr, "\n");
            fprintf(stderr, "[hb] CTA 0 thread 0 cycles per window:");
            for (int i = 0; i < 8; i++) fprintf(stderr, " %s=%.0f", nm8[i], (double)pin_stats[16 + i] / nwin);
            fprintf(stderr, "\n");
        }
        const size_t nc = (size_t)c->S * c->R;
        std::vector<unsigned long long> cy(nc * 16);
        cudaMemcpy(cy.data(), c->d_ctacyc.p, sizeof(unsigned long long) * nc * 16, cudaMemcpyDeviceToHost);
        // globaltimer stamps (ns) of window 10 at the end of each phase, relative to the earliest "table done"
        unsigned long long t0 = ~0ull;
        for (size_t b = 0; b < nc; b++) t0 = std::min(t0, cy[b * 16 + 0]);
        const int ord[14] = {0, 1, 11, 12, 13, 14, 8, 9, 10, 2, 3, 4, 7, 5};
        const char *nm[14] = {"tab", "dot", "publish", "w0 draw start", "w0 poll", "w0 math", "draws of warp 0", "draws of warp 1", "next table", "pub", "bar", "upd0", "updA", "sum"};
        if (getenv("HB_DEBUG_CYCLES")[0] == '2')   // one row per CTA (block = r * S + c)
            for (size_t b = 0; b < nc; b++) {
                fprintf(stderr, "[hb cta %3zu r %2zu c %2zu]", b, b / c->S, b % c->S);
                for (int i = 0; i < 14; i++) fprintf(stderr, " %s=%lld", nm[i], (long long)(cy[b * 16 + ord[i]] - t0));
                fprintf(stderr, "\n");
            }
        for (int i = 0; i < 14; i++) {
            double mn = 1e300, mx = 0, sm = 0;
            for (size_t b = 0; b < nc; b++) { const double v = (double)(cy[b * 16 + ord[i]] - t0); mn = std::min(mn, v); mx = std::max(mx, v); sm += v; }
            fprintf(stderr, "[hb window 10, ns since first CTA started its dot] end of %-16s min %8.0f mean %8.0f max %8.0f\n", nm[i], mn, sm / (double)nc, mx);
        }
    }
    HB_TRY(check_kernel_error(c, "hb_brr_iteration"));
    c->seq_base += pin_stats[1];
    const double off = pin_small[0];
    c->shift = off;  // the launch folded the old shift into E; the new constant is this launch's base terms
    double s1 = 0.0, s2 = 0.0;
    for (uint32_t s = 0; s < c->S; s++) { c->slice_sum_h[s] = pin_small[1 + s]; s1 += pin_small[1 + s]; s2 += pin_small[1 + c->S + s]; }
    c->eps_slice_abs = 0.0; c->eps_abs_max = 0.0;   // the next launch's fixed-point grid
    for (uint32_t s = 0; s < c->S; s++) {
        c->eps_slice_abs = std::max(c->eps_slice_abs, pin_small[1 + 2 * c->S + G + s]);
        c->eps_abs_max = std::max(c->eps_abs_max, pin_small[1 + 3 * c->S + G + s]);
    }
    // sum (E+off)^2 over the N individuals
    const double e_sqn = s2 + 2.0 * off * s1 + dN * off * off;
    for (uint32_t g = 0; g < G; g++) c->bsq[g] = pin_small[1 + 2 * c->S + g];
    for (size_t x = 0; x < gk; x++) c->cass[x] = pin_cass[x];
    if (c->fh_on) c->fh_sbsqn = pin_fh[0];
    double e_sqn_g = e_sqn;
    double mu_g0 = c->mu[0];
    unsigned long long changed_all = pin_stats[5];
    if (c->nranks > 1) {
        // group statistics summed over the GPUs (MPI_Allreduce of beta_squaredNorm and cass, :2517-2518); e_sqn is that
        // of global task 0, whose sigmaE draw the reference broadcasts (:2705)
        const size_t nr = G + gk + 4;
        std::vector<double> red(nr, 0.0);
        for (uint32_t g = 0; g < G; g++) red[g] = c->bsq[g];
        for (size_t x = 0; x < gk; x++) red[G + x] = (double)c->cass[x];
        red[G + gk] = (c->t_first == 0) ? e_sqn : 0.0;
        red[G + gk + 1] = (double)pin_stats[5];
        red[G + gk + 2] = (c->t_first == 0) ? c->mu[0] : 0.0;   // mu of global task 0 (the fixed effects use its residual)
        red[G + gk + 3] = c->fh_on ? c->fh_sbsqn : 0.0;         // bayesFH: scaled sum of squares over all markers (see hb_brr_set_fh)
        HB_CUDA(cudaMemcpyAsync(c->d_red.p, red.data(), sizeof(double) * nr, cudaMemcpyHostToDevice, st));
        HB_NCCL(ncclAllReduce(c->d_red.p, c->d_red.p + nr, nr, ncclDouble, ncclSum, c->nccl, st));
        HB_CUDA(cudaMemcpyAsync(red.data(), c->d_red.p + nr, sizeof(double) * nr, cudaMemcpyDeviceToHost, st));
        HB_CUDA(cudaStreamSynchronize(st));
        for (uint32_t g = 0; g < G; g++) c->bsq[g] = red[g];
        for (size_t x = 0; x < gk; x++) c->cass[x] = (int32_t)llround(red[G + x]);
        e_sqn_g = red[G + gk];
        changed_all = (unsigned long long)llround(red[G + gk + 1]);
        mu_g0 = red[G + gk + 2];
        if (c->fh_on) c->fh_sbsqn = red[G + gk + 3];
    }

    // ---- hyper-parameters (:2525-2578, 2685-2731)
    for (uint32_t g = 0; g < G; g++) {
        c->m0[g] = 0;
        if (c->MtotGrp[g] == 0) continue;
        c->m0[g] = c->MtotGrp[g] - c->cass[g * K];
        int rowsum = 0;
        for (uint32_t k = 0; k < K; k++) rowsum += c->cass[g * K + k];
        if (c->m0[g] == 0 || rowsum == 0) {  // :2534-2542
            c->active[g] = 0;
            c->sigmaG[g] = 0.0;
            continue;
        }
        if (!c->grp_priors.empty()) { v0G = c->grp_priors[2 * g]; s02G = c->grp_priors[2 * g + 1]; }   // :2545-2548
        const double m0 = (double)c->m0[g];
        if (c->fh_on) {   // :2557-2565 (hypTau and tau are drawn again in every group's pass, as the reference does)
            if (tape && tape->fh_hyper) {
                c->fh_hypTau = tape->fh_hyper[3 * g]; c->fh_tau = tape->fh_hyper[3 * g + 1]; c->fh_c[g] = tape->fh_hyper[3 * g + 2];
            } else {
                const hb_fh_config &f = c->fh_cfg;
                c->fh_hypTau = c->hyper_rng.inv_gamma_rate(0.5 + 0.5 * f.v0t, 1.0 / (f.tau0 * f.tau0) + 1.0 / c->fh_tau);
                c->fh_tau = c->hyper_rng.inv_gamma_rate(0.5 * (m0 + f.v0t), f.v0t / c->fh_hypTau + (0.5 * c->fh_sbsqn));
                c->fh_c[g] = c->hyper_rng.inv_scaled_chisq(f.v0c + m0, (c->bsq[g] * m0 + f.v0c * f.s02c) / (f.v0c + m0));
            }
            c->sigmaG[g] = c->bsq[g];
        }
        if (tape && tape->sigmaG && tape->pi) {
            if (!c->fh_on) c->sigmaG[g] = tape->sigmaG[g];
            for (uint32_t k = 0; k < K; k++) c->pi[g * K + k] = tape->pi[g * K + k];
        } else {
            if (!c->fh_on) c->sigmaG[g] = c->hyper_rng.inv_scaled_chisq(v0G + m0, (c->bsq[g] * m0 + v0G * s02G) / (v0G + m0));  // :2570
            double s = 0.0;
            for (uint32_t k = 0; k < K; k++) {   // dirichlet(cass + dirc), dirc = 1 unless --dPriorsFile (:1184-1185, :2551-2554, :2576-2577)
                const double dk = c->dirichlet.empty() ? 1.0 : c->dirichlet[g * K + k];
                c->pi[g * K + k] = c->hyper_rng.gamma((double)c->cass[g * K + k] + dk); s += c->pi[g * K + k];
            }
            for (uint32_t k = 0; k < K; k++) c->pi[g * K + k] /= s;
        }
    }
    // ---- fixed effects (:2648-2681): gamma and epsilon, then the residual's statistics again
    int cov_launches = 0;
    if (c->F > 0) {
        const uint32_t F = c->F;
        std::vector<double> zc(F);
        if (tape && tape->xI) { for (uint32_t i = 0; i < F; i++) { HB_CHECK(tape->xI[i] >= 0 && (uint32_t)tape->xI[i] < F, HB_ERR_ARG, "hb_brr_iteration: tape xI out of range"); c->xI[i] = tape->xI[i]; } }
        else c->hyper_rng.shuffle(c->xI.data(), (int)F);                                  // :2653
        for (uint32_t i = 0; i < F; i++) zc[i] = (tape && tape->zcov) ? tape->zcov[i] : c->hyper_rng.normal();
        HB_CUDA(cudaMemcpyAsync(c->d_xI.p, c->xI.data(), sizeof(int32_t) * F, cudaMemcpyHostToDevice, st));
        HB_CUDA(cudaMemcpyAsync(c->d_zcov.p, zc.data(), sizeof(double) * F, cudaMemcpyHostToDevice, st));
        HB_CUDA(cudaMemsetAsync(c->d_covbar.p, 0, sizeof(uint32_t), st));
        CovParams Q;
        Q.N = N; Q.S = c->S; Q.L = c->L; Q.F = F;
        Q.E = c->d_E[c->cur].p;
        Q.shift = c->shift + (c->mu[0] - mu_g0);      // residual of GLOBAL task 0 (on one GPU: of task 0)
        Q.X = c->d_X.p; Q.xI = c->d_xI.p; Q.z = c->d_zcov.p; Q.gamma = c->d_gamma.p;
        Q.sigmaE = c->sigmaE; Q.denom = dNm1 + c->sigmaE / 1.0;                            // sigmaF = s02F = 1 (:1605, :2656, :2670)
        Q.part = c->d_covpart.p; Q.bar = c->d_covbar.p;
        Q.slice_sum = c->d_small.p + 1; Q.slice_sq = c->d_small.p + 1 + c->S;
        Q.slice_abs = c->d_small.p + 1 + 2 * c->S + G; Q.slice_max = c->d_small.p + 1 + 3 * c->S + G;
        void *args[] = {(void *)&Q};
        const unsigned nb = (unsigned)std::min<uint32_t>((uint32_t)c->n_sms, 148u);
        HB_CUDA(cudaLaunchCooperativeKernel((const void *)k_cov_gibbs, dim3(nb), dim3(256), args, 0, st));
        cov_launches = 1;
        HB_CUDA(cudaMemcpyAsync(pin_small, c->d_small.p, sizeof(double) * nsmall, cudaMemcpyDeviceToHost, st));
        HB_CUDA(cudaMemcpyAsync(c->gamma.data(), c->d_gamma.p, sizeof(double) * F, cudaMemcpyDeviceToHost, st));
        HB_CUDA(cudaStreamSynchronize(st));
        double t1 = 0.0, t2 = 0.0;
        c->eps_slice_abs = 0.0; c->eps_abs_max = 0.0;
        for (uint32_t s = 0; s < c->S; s++) {
            c->slice_sum_h[s] = pin_small[1 + s]; t1 += pin_small[1 + s]; t2 += pin_small[1 + c->S + s];
            c->eps_slice_abs = std::max(c->eps_slice_abs, pin_small[1 + 2 * c->S + G + s]);
            c->eps_abs_max = std::max(c->eps_abs_max, pin_small[1 + 3 * c->S + G + s]);
        }
        // sum (E + shift_g0)^2 with the residual of global task 0 (identical on every GPU: the replicas are bit-identical)
        const double sh0 = Q.shift;
        e_sqn_g = t2 + 2.0 * sh0 * t1 + dN * sh0 * sh0;
    }
    if (tape && tape->sigmaE) c->sigmaE = tape->sigmaE[0];
    else c->sigmaE = c->hyper_rng.inv_scaled_chisq(v0E + dN, (e_sqn_g + v0E * s02E) / (v0E + dN));  // :2690
    c->iteration++;
    c->order_ready = order_ahead;

    if (out) {
        memset(out, 0, sizeof(*out));
        out->sigmaE = c->sigmaE; out->e_sqn = e_sqn_g; out->epssum = epssum;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); out->loop_ms = ms;
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[3]); out->iter_ms = ms;
        out->n_sync = pin_stats[0]; out->n_windows = pin_stats[1];
        out->n_launches = 4 + (uint64_t)cov_launches + (uint64_t)fh_launches;
        out->nnz_processed = pin_stats[2]; out->nnz_updated = pin_stats[3];
        out->bed_markers = pin_stats[4]; out->markers_changed = changed_all;
        for (int i = 0; i < 8; i++) out->phase_cycles[i] = pin_stats[8 + i];
        out->windows_ahead = pin_stats[6]; out->draws_repeated = pin_stats[7];
    }
    return HB_OK;
}

int hb_brr_get_hyper(hb_ctx *c, double *sigmaG, double *pi, double *sigmaE, double *mu_tasks_local, double *bsq,
                     int32_t *cass, int32_t *m0) {
    HB_CHECK(c && c->brr_ready, HB_ERR_STATE, "hb_brr_get_hyper: call hb_brr_init first");
    if (sigmaG) std::copy(c->sigmaG.begin(), c->sigmaG.end(), sigmaG);
    if (pi) std::copy(c->pi.begin(), c->pi.end(), pi);
    if (sigmaE) *sigmaE = c->sigmaE;
    if (mu_tasks_local) std::copy(c->mu.begin(), c->mu.end(), mu_tasks_local);
    if (bsq) std::copy(c->bsq.begin(), c->bsq.end(), bsq);
    if (cass) std::copy(c->cass.begin(), c->cass.end(), cass);
    if (m0) std::copy(c->m0.begin(), c->m0.end(), m0);
    return HB_OK;
}

int hb_brr_get_state(hb_ctx *c, double *beta, int32_t *components, double *acum) {
    HB_CHECK(c && c->brr_ready, HB_ERR_STATE, "hb_brr_get_state: call hb_brr_init first");
    HB_CUDA(cudaSetDevice(c->dev));
    if (beta) HB_CUDA(cudaMemcpyAsync(beta, c->d_beta.p, sizeof(double) * c->M, cudaMemcpyDeviceToHost, c->stream));
    if (components) HB_CUDA(cudaMemcpyAsync(components, c->d_comp.p, sizeof(int32_t) * c->M, cudaMemcpyDeviceToHost, c->stream));
    if (acum) HB_CUDA(cudaMemcpyAsync(acum, c->d_acum.p, sizeof(double) * c->M, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return HB_OK;
}

// Asynchronous form for writers that thin every iteration (.bet/.cpn/.acu, :2768-2785): the three arrays are snapshotted on the
// device (3 x M, ~10 us) and copied to the caller's (pinned) buffers on a second stream while the next hb_brr_iteration runs.
// The buffers belong to the library until hb_brr_state_wait returns; a second call waits for the first copy itself.
int hb_brr_get_state_async(hb_ctx *c, double *beta, int32_t *components, double *acum) {
    HB_CHECK(c && c->brr_ready, HB_ERR_STATE, "hb_brr_get_state_async: call hb_brr_init first");
    HB_CUDA(cudaSetDevice(c->dev));
    if (!c->copy_stream) {
        HB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        HB_CUDA(cudaEventCreateWithFlags(&c->ev_snap, cudaEventDisableTiming));
        HB_CUDA(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
        HB_TRY(c->d_snap_beta.alloc(c->M)); HB_TRY(c->d_snap_acum.alloc(c->M)); HB_TRY(c->d_snap_comp.alloc(c->M));
    }
    if (c->copy_pending) HB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));   // the previous copy still reads the snapshot
    if (beta) HB_CUDA(cudaMemcpyAsync(c->d_snap_beta.p, c->d_beta.p, sizeof(double) * c->M, cudaMemcpyDeviceToDevice, c->stream));
    if (components) HB_CUDA(cudaMemcpyAsync(c->d_snap_comp.p, c->d_comp.p, sizeof(int32_t) * c->M, cudaMemcpyDeviceToDevice, c->stream));
    if (acum) HB_CUDA(cudaMemcpyAsync(c->d_snap_acum.p, c->d_acum.p, sizeof(double) * c->M, cudaMemcpyDeviceToDevice, c->stream));
    HB_CUDA(cudaEventRecord(c->ev_snap, c->stream));
    HB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_snap, 0));
    if (beta) HB_CUDA(cudaMemcpyAsync(beta, c->d_snap_beta.p, sizeof(double) * c->M, cudaMemcpyDeviceToHost, c->copy_stream));
    if (components) HB_CUDA(cudaMemcpyAsync(components, c->d_snap_comp.p, sizeof(int32_t) * c->M, cudaMemcpyDeviceToHost, c->copy_stream));
    if (acum) HB_CUDA(cudaMemcpyAsync(acum, c->d_snap_acum.p, sizeof(double) * c->M, cudaMemcpyDeviceToHost, c->copy_stream));
    HB_CUDA(cudaEventRecord(c->ev_copy, c->copy_stream));
    c->copy_pending = true;
    return HB_OK;
}

int hb_brr_state_wait(hb_ctx *c) {
    HB_CHECK(c && c->brr_ready, HB_ERR_STATE, "hb_brr_state_wait: call hb_brr_init first");
    if (!c->copy_pending) return HB_OK;
    HB_CUDA(cudaSetDevice(c->dev));
    HB_CUDA(cudaEventSynchronize(c->ev_copy));
    c->copy_pending = false;
    return HB_OK;
}

int hb_brr_set_state(hb_ctx *c, const double *beta, const int32_t *components) {
    HB_CHECK(c && c->brr_ready, HB_ERR_STATE, "hb_brr_set_state: call hb_brr_init first");
    c->order_ready = false;   // the window layout carries the effects
    HB_CUDA(cudaSetDevice(c->dev));
    if (beta) HB_CUDA(cudaMemcpyAsync(c->d_beta.p, beta, sizeof(double) * c->M, cudaMemcpyHostToDevice, c->stream));
    if (components) HB_CUDA(cudaMemcpyAsync(c->d_comp.p, components, sizeof(int32_t) * c->M, cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return HB_OK;
}

// The reference's own restart (src/BayesRRm.cpp:842-928): the chain state is read back from the OUTPUT files of the last save
// point -- .csv (sigmaG, sigmaE, pi), .xbet / .xcpn (effects, components), .mus.<task>, .eps.<task>, .mrk.<task> -- which this
// library's host writes in the reference's layouts. Everything those files hold is put back; the random streams are not in
// them (the reference's .rng files hold Boost engines), so they are re-seeded from (seed, iteration): the chain continues from the
// saved state with fresh draws, not bit-for-bit (hb_brr_load_state does that).
int hb_brr_restore_outputs(hb_ctx *c, uint32_t iterations_done, const double *sigmaG, const double *pi, double sigmaE,
                           const double *mu_tasks_local, const double *beta, const int32_t *components, const double *eps_task0,
                           const int32_t *perm_local) {
    HB_CHECK(c && c->brr_ready && !c->bw_ready, HB_ERR_STATE, "hb_brr_restore_outputs: call hb_brr_init first");
    HB_CHECK(sigmaG && pi && mu_tasks_local && beta && components && eps_task0, HB_ERR_ARG, "hb_brr_restore_outputs: null argument");
    HB_CUDA(cudaSetDevice(c->dev));
    if (c->prefetch.joinable()) c->prefetch.join();
    c->have_next = false;
    c->order_ready = false;
    c->iteration = iterations_done;
    for (uint32_t g = 0; g < c->G; g++) {
        c->sigmaG[g] = sigmaG[g];
        if (g < c->active.size()) c->active[g] = (sigmaG[g] != 0.0) ? 1 : 0;   // adaV follows sigmaG (:1592-1597)
        double ps = 0.0;
        for (uint32_t k = 0; k < c->K; k++) { c->pi[g * c->K + k] = pi[g * c->K + k]; ps += pi[g * c->K + k]; }
        HB_CHECK(fabs(ps - 1.0) < 1e-6, HB_ERR_ARG, "hb_brr_restore_outputs: pi of group %u sums to %.9g", g, ps);
    }
    HB_CHECK(sigmaE > 0.0, HB_ERR_ARG, "hb_brr_restore_outputs: sigmaE = %g", sigmaE);
    c->sigmaE = sigmaE;
    for (uint32_t t = 0; t < c->T; t++) c->mu[t] = mu_tasks_local[t];
    HB_TRY(hb_set_epsilon(c, eps_task0));                    // residual of local task 0; the other tasks differ by mu (hb_brr_get_task_epsilon)
    for (uint32_t s0 = 0, sl = 0; sl < c->S; s0 += c->L, sl++) {
        double v = 0.0;
        for (uint32_t i = s0; i < std::min(c->N, s0 + c->L); i++) v += eps_task0[i];
        c->slice_sum_h[sl] = v;
    }
    HB_TRY(hb_brr_set_state(c, beta, components));
    HB_CUDA(cudaMemset(c->d_acum.p, 0, sizeof(double) * c->M));
    if (perm_local) {
        size_t o = 0;
        for (uint32_t t = 0; t < c->T; t++) {
            const int32_t len = c->blkL[c->t_first + t];
            for (int32_t j = 0; j < len; j++, o++) {
                HB_CHECK(perm_local[o] >= 0 && perm_local[o] < len, HB_ERR_ARG, "hb_brr_restore_outputs: marker order of task %u: %d out of range", t, perm_local[o]);
                c->perm[o] = perm_local[o];
            }
        }
    }
    const uint32_t mix = iterations_done * 0x9E3779B9u;
    for (uint32_t t = 0; t < c->T; t++) c->task_rng[t].seed((c->seed + 1000u * (c->t_first + t)) ^ mix);
    c->hyper_rng.seed((c->seed ^ 0x5bd1e995u) ^ mix);
    return HB_OK;
}

int hb_brr_get_task_epsilon(hb_ctx *c, uint32_t task_local, double *eps) {
    HB_CHECK(c && c->brr_ready && eps, HB_ERR_STATE, "hb_brr_get_task_epsilon: call hb_brr_init first");
    HB_CHECK(task_local < c->T, HB_ERR_ARG, "task %u out of range", task_local);
    HB_TRY(hb_get_epsilon(c, eps));  // eps of local task 0
    const double d = c->mu[0] - c->mu[task_local];
    if (d != 0.0)
        for (uint32_t i = 0; i < c->N; i++) eps[i] += d;
    return HB_OK;
}

int hb_brr_get_task_perm(hb_ctx *c, uint32_t task_local, int32_t *perm) {
    HB_CHECK(c && c->brr_ready && perm, HB_ERR_STATE, "hb_brr_get_task_perm: call hb_brr_init first");
    HB_CHECK(task_local < c->T, HB_ERR_ARG, "task %u out of range", task_local);
    size_t o = 0;
    for (uint32_t t = 0; t < task_local; t++) o += (size_t)c->blkL[c->t_first + t];
    std::copy(c->perm.begin() + o, c->perm.begin() + o + c->blkL[c->t_first + task_local], perm);
    return HB_OK;
}

int hb_brr_set_covariates(hb_ctx *c, const double *X, uint32_t n_cov) {
    HB_CHECK(c && c->brr_ready, HB_ERR_STATE, "hb_brr_set_covariates: call hb_brr_init first");
    HB_CHECK(n_cov == 0 || X, HB_ERR_ARG, "hb_brr_set_covariates: null matrix");
    HB_CHECK(c->cfg.model == 0, HB_ERR_ARG, "hb_brr_set_covariates: fixed effects are available for BayesRRm only");
    HB_CUDA(cudaSetDevice(c->dev));
    c->F = n_cov;
    c->gamma.assign(n_cov, 0.0);                      // gamma.setZero(), :1092
    c->xI.resize(n_cov);
    for (uint32_t i = 0; i < n_cov; i++) c->xI[i] = (int32_t)i;   // :1113-1117
    if (n_cov == 0) return HB_OK;
    const size_t n_all = (size_t)c->S * c->L;
    std::vector<double> col(n_all * n_cov, 0.0);      // column-major on the slice layout, padded individuals 0
    for (uint32_t k = 0; k < c->N; k++)
        for (uint32_t f = 0; f < n_cov; f++) col[(size_t)f * n_all + k] = X[(size_t)k * n_cov + f];
    HB_TRY(c->d_X.alloc(n_all * n_cov));
    HB_CUDA(cudaMemcpy(c->d_X.p, col.data(), sizeof(double) * n_all * n_cov, cudaMemcpyHostToDevice));
    HB_TRY(c->d_gamma.alloc(n_cov)); HB_TRY(c->d_zcov.alloc(n_cov)); HB_TRY(c->d_xI.alloc(n_cov));
    HB_TRY(c->d_covpart.alloc(2 * 148)); HB_TRY(c->d_covbar.alloc(1));
    HB_CUDA(cudaMemset(c->d_gamma.p, 0, sizeof(double) * n_cov));
    return HB_OK;
}

int hb_brr_get_gamma(hb_ctx *c, double *gamma, int32_t *xI) {
    HB_CHECK(c && c->brr_ready, HB_ERR_STATE, "hb_brr_get_gamma: call hb_brr_init first");
    if (gamma) std::copy(c->gamma.begin(), c->gamma.end(), gamma);
    if (xI) std::copy(c->xI.begin(), c->xI.end(), xI);
    return HB_OK;
}

int hb_brr_set_group_priors(hb_ctx *c, const double *v0G_s02G, const double *dirichlet) {
    HB_CHECK(c && c->brr_ready, HB_ERR_STATE, "hb_brr_set_group_priors: call hb_brr_init first");
    HB_CHECK(c->cfg.model == 0, HB_ERR_ARG, "hb_brr_set_group_priors: BayesRRm / BayesFH only");
    c->grp_priors.clear(); c->dirichlet.clear();
    if (v0G_s02G) {
        for (uint32_t g = 0; g < c->G; g++)
            HB_CHECK(v0G_s02G[2 * g] > 0.0 && v0G_s02G[2 * g + 1] > 0.0, HB_ERR_ARG, "hb_brr_set_group_priors: v0G and s02G of group %u must be positive", g);
        c->grp_priors.assign(v0G_s02G, v0G_s02G + 2 * (size_t)c->G);
    }
    if (dirichlet) {
        for (size_t x = 0; x < (size_t)c->G * c->K; x++)
            HB_CHECK(dirichlet[x] > 0.0, HB_ERR_ARG, "hb_brr_set_group_priors: Dirichlet parameter %zu must be positive", x);
        c->dirichlet.assign(dirichlet, dirichlet + (size_t)c->G * c->K);
    }
    return HB_OK;
}

int hb_brr_set_fh(hb_ctx *c, const hb_fh_config *cfg, const double *state0) {
    HB_CHECK(c && c->brr_ready, HB_ERR_STATE, "hb_brr_set_fh: call hb_brr_init first");
    HB_CHECK(c->cfg.model == 0, HB_ERR_ARG, "hb_brr_set_fh: bayesFH is a variant of BayesRRm");
    HB_CUDA(cudaSetDevice(c->dev));
    if (!cfg) { c->fh_on = false; return HB_OK; }
    // The reference neither reduces the scaled sum of squares over its ranks nor broadcasts tau / c_slab (src/BayesRRm.cpp:2503-2510,
    // 2557-2565): with several ranks every rank follows its own FH parameters (flagged, not reproduced). Here ONE set of parameters:
    // the sum runs over all markers of all tasks and GPUs (all-reduced with the group statistics), and tau / hypTau / c_slab are
    // drawn on every GPU from the common hyper-parameter stream.
    HB_CHECK(cfg->v0L > 0.0 && cfg->v0t > 0.0 && cfg->v0c > 0.0 && cfg->s02c > 0.0 && cfg->tau0 > 0.0, HB_ERR_ARG,
             "hb_brr_set_fh: v0L, v0t, v0c, s02c and tau0 must be positive");
    const uint32_t G = c->G, M = c->M;
    c->fh_cfg = *cfg;
    c->fh_c.assign(G, 0.0);
    if (state0) {
        c->fh_hypTau = state0[0]; c->fh_tau = state0[1];
        for (uint32_t g = 0; g < G; g++) c->fh_c[g] = state0[2 + g];
        HB_CHECK(c->fh_tau > 0.0, HB_ERR_ARG, "hb_brr_set_fh: tau must be positive");
    } else {   // :1147-1154, from the hyper-parameter stream
        c->fh_hypTau = c->hyper_rng.inv_gamma_rate(0.5, 1.0 / (cfg->tau0 * cfg->tau0));
        c->fh_tau = c->hyper_rng.inv_gamma_rate(0.5 * cfg->v0t, cfg->v0t / c->fh_hypTau);
        for (uint32_t g = 0; g < G; g++) c->fh_c[g] = c->hyper_rng.inv_scaled_chisq(cfg->v0c, cfg->s02c);
    }
    double cs = 0.0;
    for (uint32_t g = 0; g < G; g++) cs += c->fh_c[g];
    HB_TRY(c->d_fh_lambda.alloc(M)); HB_TRY(c->d_fh_nu.alloc(M)); HB_TRY(c->d_fh_par.alloc(3 * (size_t)M));
    HB_TRY(c->d_fh_c.alloc(G)); HB_TRY(c->d_fh_part.alloc(kSqChunks + 1)); HB_TRY(c->d_fh_g.alloc(2 * (size_t)M));
    std::vector<double> lam(M, cs / (double)c->Mtot);                                    // :1161
    HB_CUDA(cudaMemcpy(c->d_fh_lambda.p, lam.data(), sizeof(double) * M, cudaMemcpyHostToDevice));
    HB_CUDA(cudaMemset(c->d_fh_nu.p, 0, sizeof(double) * M));
    c->fh_sbsqn = 0.0;
    c->fh_on = true;
    return HB_OK;
}

int hb_brr_get_fh(hb_ctx *c, double *scalars3, double *c_slab, double *lambda_var, double *nu_var) {
    HB_CHECK(c && c->brr_ready && c->fh_on, HB_ERR_STATE, "hb_brr_get_fh: call hb_brr_set_fh first");
    HB_CUDA(cudaSetDevice(c->dev));
    if (scalars3) { scalars3[0] = c->fh_hypTau; scalars3[1] = c->fh_tau; scalars3[2] = c->fh_sbsqn; }
    if (c_slab) std::copy(c->fh_c.begin(), c->fh_c.end(), c_slab);
    HB_CUDA(cudaStreamSynchronize(c->stream));
    if (lambda_var) HB_CUDA(cudaMemcpy(lambda_var, c->d_fh_lambda.p, sizeof(double) * c->M, cudaMemcpyDeviceToHost));
    if (nu_var) HB_CUDA(cudaMemcpy(nu_var, c->d_fh_nu.p, sizeof(double) * c->M, cudaMemcpyDeviceToHost));
    return HB_OK;
}

int hb_comm_get_unique_id(uint8_t id[HB_NCCL_ID_BYTES]) {
    HB_CHECK(id, HB_ERR_ARG, "null argument");
    static_assert(sizeof(ncclUniqueId) <= HB_NCCL_ID_BYTES, "ncclUniqueId larger than HB_NCCL_ID_BYTES");
    ncclUniqueId u;
    HB_NCCL(ncclGetUniqueId(&u));
    memset(id, 0, HB_NCCL_ID_BYTES);
    memcpy(id, &u, sizeof(u));
    return HB_OK;
}

// One process per GPU. Sets up (1) an NCCL communicator for the per-iteration statistics that hydra all-reduces
// (src/BayesRRm.cpp:2517-2518) and the bootstrap, and (2) the NVLink peer mapping of every GPU's inbox, through
// which the marker kernel exchanges the changed markers of a synchronisation window (replaces :2051, :2456).
int hb_comm_init(hb_ctx *c, const uint8_t id[HB_NCCL_ID_BYTES], int rank, int nranks) {
    HB_CHECK(c && id, HB_ERR_ARG, "null argument");
    HB_CHECK(nranks >= 1 && nranks <= (int)kMaxRanks && rank >= 0 && rank < nranks, HB_ERR_ARG, "hb_comm_init: rank %d of %d (max %u GPUs)", rank, nranks, kMaxRanks);
    HB_CHECK(!c->nccl, HB_ERR_STATE, "hb_comm_init: already initialised");
    HB_CUDA(cudaSetDevice(c->dev));
    if (nranks == 1) return HB_OK;
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    HB_NCCL(ncclCommInitRank(&c->nccl, nranks, u, rank));
    c->rank = rank; c->nranks = nranks;
    const char *mb = getenv("HB_INBOX_MB");
    c->inbox_stride = ((size_t)(mb ? atoi(mb) : 16) << 20);
    HB_CHECK(c->inbox_stride >= kInboxHeader + (1u << 20), HB_ERR_ARG, "hb_comm_init: HB_INBOX_MB too small");
    c->comm_bytes = 2 * (size_t)nranks * c->inbox_stride + 256;
    cudaError_t e = cudaMalloc((void **)&c->inbox, c->comm_bytes);
    HB_CHECK(e == cudaSuccess, HB_ERR_NOMEM, "hb_comm_init: inbox of %zu bytes: %s", c->comm_bytes, cudaGetErrorString(e));
    HB_CUDA(cudaMemset(c->inbox, 0, c->comm_bytes));
    c->flags = reinterpret_cast<unsigned long long *>(c->inbox + c->comm_bytes - 256);
    // exchange the IPC handles of the inboxes with NCCL itself (no MPI / torch needed by the C ABI)
    cudaIpcMemHandle_t mine;
    HB_CUDA(cudaIpcGetMemHandle(&mine, c->inbox));
    DevBuf<unsigned char> d_h;
    HB_TRY(d_h.alloc((size_t)nranks * sizeof(mine)));
    HB_CUDA(cudaMemcpy(d_h.p + (size_t)rank * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice));
    HB_NCCL(ncclAllGather(d_h.p + (size_t)rank * sizeof(mine), d_h.p, sizeof(mine), ncclChar, c->nccl, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<cudaIpcMemHandle_t> all(nranks);
    HB_CUDA(cudaMemcpy(all.data(), d_h.p, (size_t)nranks * sizeof(mine), cudaMemcpyDeviceToHost));
    for (int h = 0; h < nranks; h++) {
        if (h == rank) continue;
        void *ptr = nullptr;
        e = cudaIpcOpenMemHandle(&ptr, all[h], cudaIpcMemLazyEnablePeerAccess);
        HB_CHECK(e == cudaSuccess, HB_ERR_CUDA, "hb_comm_init: cannot map the inbox of rank %d over NVLink/PCIe (%s); the GPUs must be peers on one node", h, cudaGetErrorString(e));
        c->peer_inbox[h] = static_cast<unsigned char *>(ptr);
    }
    // nobody may write into an inbox before everybody has mapped (and cleared) theirs; and all GPUs must run the same
    // grid, because the arrival counters of the exchange count CTAs
    DevBuf<int32_t> d_g;
    HB_TRY(d_g.alloc(2));
    const int32_t mine_g[2] = {(int32_t)(c->S * c->R), -(int32_t)(c->S * c->R)};
    int32_t mx[2] = {0, 0};
    HB_CUDA(cudaMemcpy(d_g.p, mine_g, sizeof(mine_g), cudaMemcpyHostToDevice));
    HB_NCCL(ncclAllReduce(d_g.p, d_g.p, 2, ncclInt32, ncclMax, c->nccl, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    HB_CUDA(cudaMemcpy(mx, d_g.p, sizeof(mx), cudaMemcpyDeviceToHost));
    HB_CHECK(mx[0] == -mx[1], HB_ERR_ARG, "hb_comm_init: the GPUs run different grids (%d vs %d CTAs); use the same n_slices / max_ctas", mx[0], -mx[1]);
    if (c->brr_ready) { const uint64_t sv = c->seed; HB_TRY(check_equal_over_ranks(c, &sv, 1, "the seed (give every process the same --seed)")); }
    return HB_OK;
}

int hb_comm_check_equal(hb_ctx *c, const uint64_t *vals, uint32_t n, const char *what) {
    HB_CHECK(c && vals && n > 0 && n <= 64, HB_ERR_ARG, "hb_comm_check_equal: bad argument");
    HB_CUDA(cudaSetDevice(c->dev));
    return check_equal_over_ranks(c, vals, n, what ? what : "a value");
}

}  // extern "C"


// ---- restart (hydra's --restart: src/BayesRRm.cpp:842-928 reads .csv/.bet/.cpn/.eps/.mrk/.mus/.rng back).
// The complete state of the BayesRRm chain of this GPU as one opaque, self-describing blob: hyper-parameters, per-task mu,
// the residual, marker effects / components / Acum, the host random streams (task streams, hyper-parameter stream, the
// already drawn order of the next iteration). A run continued from the blob is bit-identical to the uninterrupted run.
namespace hb {
struct BlobW {
    std::vector<unsigned char> b;
    template <class T> void put(const T *p, size_t n) { const unsigned char *q = reinterpret_cast<const unsigned char *>(p); b.insert(b.end(), q, q + n * sizeof(T)); }
    template <class T> void one(const T &v) { put(&v, 1); }
    void str(const std::string &s) { const uint64_t n = s.size(); one(n); put(s.data(), s.size()); }
};
struct BlobR {
    const unsigned char *p, *e;
    bool ok = true;
    template <class T> void get(T *d, size_t n) {
        if (!ok || (size_t)(e - p) < n * sizeof(T)) { ok = false; return; }
        memcpy(d, p, n * sizeof(T)); p += n * sizeof(T);
    }
    template <class T> T one() { T v{}; get(&v, 1); return v; }
    std::string str() { const uint64_t n = one<uint64_t>(); std::string s; if (ok && (uint64_t)(e - p) >= n) { s.assign(reinterpret_cast<const char *>(p), n); p += n; } else ok = false; return s; }
};
static std::string rng_text(const HostRng &r) { std::ostringstream o; o << r.eng; return o.str(); }
static bool rng_from_text(HostRng &r, const std::string &t) { std::istringstream i(t); i >> r.eng; return !i.fail(); }
constexpr uint32_t kRestartMagic = 0x32524248u;  // "HBR2"
}  // namespace hb

extern "C" {

// .rng.<task> of the reference (src/distributions_boost.cpp:38-44: `file << rng`, the text form of a boost::mt19937). Here the
// text form of the task's std::mt19937 (the same generator, libstdc++'s operator<<: 624 state words and the position).
int hb_brr_get_task_rng(hb_ctx *c, uint32_t task_local, char *buf, size_t cap, size_t *need) {
    HB_CHECK(c && c->brr_ready && need, HB_ERR_STATE, "hb_brr_get_task_rng: call hb_brr_init first");
    HB_CHECK(task_local < c->T, HB_ERR_ARG, "hb_brr_get_task_rng: task %u of %u", task_local, c->T);
    if (c->prefetch.joinable()) c->prefetch.join();  // the worker thread owns the task streams while it runs
    const std::string t = rng_text(c->task_rng[task_local]);
    *need = t.size() + 1;
    if (!buf) return HB_OK;
    HB_CHECK(cap >= t.size() + 1, HB_ERR_ARG, "hb_brr_get_task_rng: buffer of %zu bytes, %zu needed", cap, t.size() + 1);
    memcpy(buf, t.c_str(), t.size() + 1);
    return HB_OK;
}

int hb_brr_save_state(hb_ctx *c, void *buf, size_t cap, size_t *need) {
    HB_CHECK(c && c->brr_ready && need, HB_ERR_STATE, "hb_brr_save_state: call hb_brr_init first");
    HB_CUDA(cudaSetDevice(c->dev));
    if (c->prefetch.joinable()) c->prefetch.join();  // the worker thread owns the task streams while it runs
    const size_t nE = (size_t)c->S * c->L;
    BlobW w;
    const uint32_t head[12] = {kRestartMagic, (uint32_t)HB_ABI_VERSION, c->N, c->M, c->T, c->G, c->K, c->S, c->L, c->seed, c->iteration, c->have_next ? 1u : 0u};
    w.put(head, 12);
    w.one(c->shift); w.one(c->sigmaE); w.one(c->seq_base);
    const double bwv[4] = {c->bw_mu, c->bw_alpha, c->bw_sumSigmaG, c->bw_ready ? 1.0 : 0.0};  // BayesW scalars (the same blob serves both models)
    w.put(bwv, 4);
    w.put(c->sigmaG.data(), c->G); w.put(c->pi.data(), (size_t)c->G * c->K); w.put(c->mu.data(), c->T); w.put(c->bsq.data(), c->G);
    w.put(c->cass.data(), (size_t)c->G * c->K); w.put(c->m0.data(), c->G);
    w.put(c->slice_sum_h.data(), c->S);
    w.put(c->perm.data(), c->M);
    if (c->have_next) { w.put(c->perm_next.data(), c->M); w.put(c->zmu_next.data(), c->T); }
    for (uint32_t t = 0; t < c->T; t++) w.str(rng_text(c->task_rng[t]));
    w.str(rng_text(c->hyper_rng));
    std::vector<double> hd(std::max(nE, (size_t)c->M));
    std::vector<int32_t> hi(c->M);
    HB_CUDA(cudaMemcpy(hd.data(), c->d_E[c->bw_ready ? 0 : c->cur].p, sizeof(double) * nE, cudaMemcpyDeviceToHost)); w.put(hd.data(), nE);
    HB_CUDA(cudaMemcpy(hd.data(), c->d_beta.p, sizeof(double) * c->M, cudaMemcpyDeviceToHost)); w.put(hd.data(), c->M);
    HB_CUDA(cudaMemcpy(hd.data(), c->d_acum.p, sizeof(double) * c->M, cudaMemcpyDeviceToHost)); w.put(hd.data(), c->M);
    HB_CUDA(cudaMemcpy(hi.data(), c->d_comp.p, sizeof(int32_t) * c->M, cudaMemcpyDeviceToHost)); w.put(hi.data(), c->M);
    { const uint32_t F = c->F; w.one(F); w.put(c->gamma.data(), F); w.put(c->xI.data(), F); }   // fixed effects (.gam / .xiv of the reference)
    {   // bayesFH: global / slab scales and the local scales
        const uint32_t fh = c->fh_on ? 1u : 0u;
        w.one(fh);
        if (fh) {
            const double sc[3] = {c->fh_hypTau, c->fh_tau, c->fh_sbsqn};
            w.put(sc, 3); w.put(c->fh_c.data(), c->G);
            HB_CUDA(cudaMemcpy(hd.data(), c->d_fh_lambda.p, sizeof(double) * c->M, cudaMemcpyDeviceToHost)); w.put(hd.data(), c->M);
            HB_CUDA(cudaMemcpy(hd.data(), c->d_fh_nu.p, sizeof(double) * c->M, cudaMemcpyDeviceToHost)); w.put(hd.data(), c->M);
        }
    }
    *need = w.b.size();
    if (!buf) return HB_OK;
    HB_CHECK(cap >= w.b.size(), HB_ERR_ARG, "hb_brr_save_state: buffer of %zu bytes, %zu needed", cap, w.b.size());
    memcpy(buf, w.b.data(), w.b.size());
    return HB_OK;
}

int hb_brr_load_state(hb_ctx *c, const void *buf, size_t n) {
    HB_CHECK(c && c->brr_ready && buf, HB_ERR_STATE, "hb_brr_load_state: call hb_brr_init (same data, groups, mixtures, tasks) first");
    HB_CUDA(cudaSetDevice(c->dev));
    if (c->prefetch.joinable()) c->prefetch.join();
    BlobR r{static_cast<const unsigned char *>(buf), static_cast<const unsigned char *>(buf) + n};
    uint32_t head[12] = {0};
    r.get(head, 12);
    HB_CHECK(r.ok && head[0] == kRestartMagic, HB_ERR_ARG, "hb_brr_load_state: not a hydra_b200 restart state");
    HB_CHECK(head[1] == (uint32_t)HB_ABI_VERSION, HB_ERR_ARG, "hb_brr_load_state: the state was written by ABI version %u of the library, this is version %d "
             "(restart from the output files instead: remove the .rst.<rank> files)", head[1], HB_ABI_VERSION);
    HB_CHECK(head[2] == c->N && head[3] == c->M && head[4] == c->T && head[5] == c->G && head[6] == c->K && head[7] == c->S && head[8] == c->L,
             HB_ERR_ARG, "hb_brr_load_state: the state belongs to another problem (N %u M %u tasks %u groups %u mixtures %u slices %u x %u; here %u %u %u %u %u %u x %u)",
             head[2], head[3], head[4], head[5], head[6], head[7], head[8], c->N, c->M, c->T, c->G, c->K, c->S, c->L);
    const size_t nE = (size_t)c->S * c->L;
    c->seed = head[9]; c->iteration = head[10]; c->have_next = head[11] != 0;
    c->order_ready = false;
    c->shift = r.one<double>(); c->sigmaE = r.one<double>(); c->seq_base = r.one<unsigned long long>();
    double bwv[4] = {0, 0, 0, 0};
    r.get(bwv, 4);
    HB_CHECK(r.ok && (bwv[3] != 0.0) == c->bw_ready, HB_ERR_ARG, "hb_brr_load_state: the state belongs to the other model (BayesRRm / BayesW)");
    if (c->bw_ready) { c->bw_mu = bwv[0]; c->bw_alpha = bwv[1]; c->bw_sumSigmaG = bwv[2]; }
    r.get(c->sigmaG.data(), c->G); r.get(c->pi.data(), (size_t)c->G * c->K); r.get(c->mu.data(), c->T); r.get(c->bsq.data(), c->G);
    r.get(c->cass.data(), (size_t)c->G * c->K); r.get(c->m0.data(), c->G);
    for (uint32_t g = 0; g < c->G && g < c->active.size(); g++) c->active[g] = (c->sigmaG[g] != 0.0) ? 1 : 0;  // adaV follows the restored sigmaG (:1592-1597)
    r.get(c->slice_sum_h.data(), c->S);
    r.get(c->perm.data(), c->M);
    if (c->have_next) { c->perm_next.resize(c->M); c->zmu_next.assign(c->T, 0.0); r.get(c->perm_next.data(), c->M); r.get(c->zmu_next.data(), c->T); }
    for (uint32_t t = 0; t < c->T; t++) HB_CHECK(rng_from_text(c->task_rng[t], r.str()), HB_ERR_ARG, "hb_brr_load_state: bad task stream %u", t);
    HB_CHECK(rng_from_text(c->hyper_rng, r.str()), HB_ERR_ARG, "hb_brr_load_state: bad hyper-parameter stream");
    std::vector<double> hd(std::max(nE, (size_t)c->M));
    std::vector<int32_t> hi(c->M);
    r.get(hd.data(), nE); HB_CHECK(r.ok, HB_ERR_ARG, "hb_brr_load_state: truncated state");
    HB_CUDA(cudaMemcpy(c->d_E[c->bw_ready ? 0 : c->cur].p, hd.data(), sizeof(double) * nE, cudaMemcpyHostToDevice));
    c->eps_slice_abs = 0.0; c->eps_abs_max = 0.0;
    for (uint32_t s0 = 0; s0 < c->N; s0 += c->L) {
        double a = 0.0;
        for (uint32_t i = s0; i < std::min(c->N, s0 + c->L); i++) { a += fabs(hd[i]); c->eps_abs_max = std::max(c->eps_abs_max, fabs(hd[i])); }
        c->eps_slice_abs = std::max(c->eps_slice_abs, a);
    }
    r.get(hd.data(), c->M); HB_CUDA(cudaMemcpy(c->d_beta.p, hd.data(), sizeof(double) * c->M, cudaMemcpyHostToDevice));
    r.get(hd.data(), c->M); HB_CUDA(cudaMemcpy(c->d_acum.p, hd.data(), sizeof(double) * c->M, cudaMemcpyHostToDevice));
    r.get(hi.data(), c->M); HB_CHECK(r.ok, HB_ERR_ARG, "hb_brr_load_state: truncated state");
    HB_CUDA(cudaMemcpy(c->d_comp.p, hi.data(), sizeof(int32_t) * c->M, cudaMemcpyHostToDevice));
    {
        const uint32_t F = r.one<uint32_t>();
        HB_CHECK(r.ok && F == c->F, HB_ERR_ARG, "hb_brr_load_state: the state has %u fixed effects, this chain %u (call hb_brr_set_covariates first)", F, c->F);
        r.get(c->gamma.data(), F); r.get(c->xI.data(), F);
        HB_CHECK(r.ok, HB_ERR_ARG, "hb_brr_load_state: truncated state");
        if (F && !c->bw_ready) HB_CUDA(cudaMemcpy(c->d_gamma.p, c->gamma.data(), sizeof(double) * F, cudaMemcpyHostToDevice));   // (BayesW keeps gamma on the host)
    }
    {
        const uint32_t fh = r.one<uint32_t>();
        HB_CHECK(r.ok && (fh != 0) == c->fh_on, HB_ERR_ARG, "hb_brr_load_state: the state %s a bayesFH chain, this chain %s (call hb_brr_set_fh first)",
                 fh ? "is" : "is not", c->fh_on ? "is" : "is not");
        if (fh) {
            double sc[3] = {0, 0, 0};
            r.get(sc, 3); r.get(c->fh_c.data(), c->G);
            c->fh_hypTau = sc[0]; c->fh_tau = sc[1]; c->fh_sbsqn = sc[2];
            r.get(hd.data(), c->M); HB_CHECK(r.ok, HB_ERR_ARG, "hb_brr_load_state: truncated state");
            HB_CUDA(cudaMemcpy(c->d_fh_lambda.p, hd.data(), sizeof(double) * c->M, cudaMemcpyHostToDevice));
            r.get(hd.data(), c->M); HB_CHECK(r.ok, HB_ERR_ARG, "hb_brr_load_state: truncated state");
            HB_CUDA(cudaMemcpy(c->d_fh_nu.p, hd.data(), sizeof(double) * c->M, cudaMemcpyHostToDevice));
        }
    }
    c->eps_set = true;
    return HB_OK;
}

}  // extern "C"

#include "hydra_b200_bw.inc"
