"""Host-side mirror of the reference's interface for the per-marker Gibbs hot path.

hydra has no plugin API; its seam is class BayesRRm (reference src/BayesRRm.h:116-150) and the
Data loaders (src/data.hpp:88-357).  `GenotypeStore` plays the role of `Data` for the genotype
representations (load_data_from_bed_file / _sparse_files / _mixed_representations) and
`BayesRRm` that of the sampler (sparse_dotprod, sparse_scaadd, runMpiGibbs' iteration body).
Everything numeric happens in libhydra_b200.so on the GPU, through the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import arr, check, ptr

_REPR = {"sparse": capi.REPR_SPARSE, "bed": capi.REPR_BED, "mixed": capi.REPR_MIXED}


class GenotypeStore:
    """Device-resident genotypes of the markers owned by this GPU + the residual vector epsilon."""

    def __init__(self, n_ind, m_total, *, na_inds=None, tasks=1, task_first=0, tasks_local=None, sync_rate=1,
                 n_groups=1, n_mix=4, repr_mode="sparse", threshold_fnz=0.06, device=0, n_slices=0, max_ctas=0,
                 block_starts=None, block_lens=None, shuffle=True, model="bayesRRm"):
        self._lib = capi.load()
        cfg = capi.HbConfig()
        self._na = arr(na_inds if na_inds is not None else np.zeros(0), np.uint32)
        self._bs = arr(block_starts, np.int32)
        self._bl = arr(block_lens, np.int32)
        cfg.device, cfg.n_ind_raw, cfg.n_na = device, n_ind, len(self._na)
        cfg.na_inds = self._na.ctypes.data if len(self._na) else None
        cfg.m_total, cfg.n_tasks_total, cfg.task_first = m_total, tasks, task_first
        cfg.n_tasks_local = tasks if tasks_local is None else tasks_local
        cfg.block_starts = None if self._bs is None else self._bs.ctypes.data
        cfg.block_lens = None if self._bl is None else self._bl.ctypes.data
        cfg.sync_rate, cfg.n_groups, cfg.n_mix = sync_rate, n_groups, n_mix
        cfg.repr_mode, cfg.threshold_fnz = _REPR[repr_mode], threshold_fnz
        cfg.n_slices, cfg.max_ctas, cfg.model = n_slices, max_ctas, {"bayesRRm": 0, "bayesW": 1}[model]
        cfg.reserved[0] = 0 if shuffle else 1
        self._h = C.c_void_p()
        check(self._lib.hb_create(C.byref(cfg), C.byref(self._h)))
        v = [C.c_uint32() for _ in range(7)]
        check(self._lib.hb_get_layout(self._h, *[C.byref(x) for x in v]))
        (self.n_ind, self.m_start, self.m_local, self.n_slices, self.slice_len, self.n_cta_groups, self.lmax) = [x.value for x in v]
        self.n_ind_raw, self.m_total, self.tasks, self.tasks_local, self.task_first = n_ind, m_total, tasks, cfg.n_tasks_local, task_first
        self.n_groups, self.n_mix, self.sync_rate = n_groups, n_mix, sync_rate

    # -- life cycle
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.hb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def comm_init(self, dist=None, unique_id=None, rank=None, world=None):
        """Join the GPUs of one node (one process each): NCCL communicator for the per-iteration statistics and the
        NVLink peer mapping used by the marker kernel to exchange changed markers (replaces the MPI_Allreduce of deltaEps,
        src/BayesRRm.cpp:2051, 2456, 2517-2518). `dist` = an initialised torch.distributed (any backend) used only to
        broadcast the NCCL unique id; or pass unique_id/rank/world explicitly."""
        uid = np.zeros(capi.NCCL_ID_BYTES, np.uint8)
        if dist is not None:
            import torch
            rank, world = dist.get_rank(), dist.get_world_size()
            if rank == 0:
                check(self._lib.hb_comm_get_unique_id(ptr(uid)))
            t = torch.from_numpy(uid)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.broadcast(t, src=0)
            uid = t.cpu().numpy().copy()
        else:
            uid[:] = np.frombuffer(bytes(unique_id), np.uint8)[: capi.NCCL_ID_BYTES]
        check(self._lib.hb_comm_init(self._h, ptr(uid), C.c_int(rank), C.c_int(world)))
        self.rank, self.world = rank, world

    def task_blocks(self):
        s = np.zeros(self.tasks, np.int32)
        l = np.zeros(self.tasks, np.int32)
        check(self._lib.hb_get_task_blocks(self._h, ptr(s), ptr(l)))
        return s, l

    # -- staging (Data::load_data_from_*)
    def load_data_from_bed(self, bed_cols, m_first=0):
        """bed_cols: (n, ceil(n_ind_raw/4)) uint8, PLINK columns without the 3-byte header."""
        b = arr(bed_cols, np.uint8)
        assert b.ndim == 2 and b.shape[1] == (self.n_ind_raw + 3) // 4, b.shape
        check(self._lib.hb_stage_bed(self._h, C.c_uint32(m_first), C.c_uint32(b.shape[0]), ptr(b)))

    def load_data_from_sparse(self, I1, N1S, N1L, I2, N2S, N2L, IM, NMS, NML, m_first=0):
        a = [arr(I1, np.uint32), arr(N1S, np.uint64), arr(N1L, np.uint64), arr(I2, np.uint32), arr(N2S, np.uint64),
             arr(N2L, np.uint64), arr(IM, np.uint32), arr(NMS, np.uint64), arr(NML, np.uint64)]
        n = len(a[1])
        check(self._lib.hb_stage_sparse(self._h, C.c_uint32(m_first), C.c_uint32(n), *[ptr(x) if len(x) else None for x in a]))

    def load_synthetic(self, seed, thresholds, attempts=None, m_first=0):
        t = arr(thresholds, np.uint32)
        at = arr(attempts, np.uint32)
        check(self._lib.hb_stage_synth(self._h, C.c_uint32(m_first), C.c_uint32(t.shape[0]), C.c_uint32(seed), ptr(t), ptr(at)))

    def finalize(self):
        check(self._lib.hb_stage_finalize(self._h))

    def marker_counts(self):
        o = [np.zeros(self.m_local, np.uint32) for _ in range(3)]
        check(self._lib.hb_marker_counts(self._h, *[ptr(x) for x in o]))
        return o

    def marker_stats(self):
        a, s = np.zeros(self.m_local), np.zeros(self.m_local)
        check(self._lib.hb_marker_stats(self._h, ptr(a), ptr(s)))
        return a, s

    def marker_is_bed(self):
        f = np.zeros(self.m_local, np.uint8)
        check(self._lib.hb_marker_is_bed(self._h, ptr(f)))
        return f

    @property
    def genotype_bytes(self):
        return int(self._lib.hb_genotype_bytes(self._h))

    def export_sparse(self, m_first=0, n=None):
        n = self.m_local - m_first if n is None else n
        c1, c2, cm = self.marker_counts()
        sl = slice(m_first, m_first + n)
        I = [np.zeros(max(int(c[sl].sum()), 1), np.uint32) for c in (c1, c2, cm)]
        SL = [np.zeros(n, np.uint64) for _ in range(6)]
        check(self._lib.hb_export_sparse(self._h, C.c_uint32(m_first), C.c_uint32(n), ptr(I[0]), ptr(SL[0]), ptr(SL[1]),
                                         ptr(I[1]), ptr(SL[2]), ptr(SL[3]), ptr(I[2]), ptr(SL[4]), ptr(SL[5])))
        return (I[0][: int(c1[sl].sum())], SL[0], SL[1], I[1][: int(c2[sl].sum())], SL[2], SL[3], I[2][: int(cm[sl].sum())], SL[4], SL[5])

    def export_bed(self, m):
        out = np.zeros((self.n_ind + 3) // 4, np.uint8)
        check(self._lib.hb_export_bed(self._h, C.c_uint32(m), ptr(out)))
        return out

    # -- epsilon and the unit kernels (BayesRRm::sparse_dotprod / sparse_scaadd, LUT variants)
    def set_epsilon(self, eps):
        e = arr(eps, np.float64)
        assert e.shape == (self.n_ind,)
        check(self._lib.hb_set_epsilon(self._h, ptr(e)))

    def get_epsilon(self):
        e = np.zeros(self.n_ind)
        check(self._lib.hb_get_epsilon(self._h, ptr(e)))
        return e

    def sparse_dotprod(self, markers):
        """num_j = mstd_j * (x_j . eps) for centred columns (src/BayesRRm.cpp:316-342 / :1757-1809)."""
        m = arr(markers, np.uint32)
        out = np.zeros(len(m))
        check(self._lib.hb_dot_markers(self._h, ptr(m), C.c_uint32(len(m)), ptr(out)))
        return out

    def sparse_scaadd(self, markers, dbeta):
        """eps += sum_j dbeta_j * mstd_j * (x_j - mave_j) (src/BayesRRm.cpp:250-281, :1976-2010, :2460-2471)."""
        m = arr(markers, np.uint32)
        d = arr(dbeta, np.float64)
        assert len(m) == len(d)
        check(self._lib.hb_scaadd_markers(self._h, ptr(m), ptr(d), C.c_uint32(len(m))))


FH_DEFAULTS = dict(v0L=3.0, v0t=3.0, v0c=3.0, s02c=1.0, tau0=1.0)   # src/options.hpp:91-96


class BayesRRm:
    """The BayesRRm chain on one GPU (src/BayesRRm.cpp:933-2939, marker loop :1709-2490)."""

    def __init__(self, store: GenotypeStore, y, mS, groups=None, sigmaG0=None, seed=0, covariates=None, group_priors=None,
                 dirichlet_priors=None, fh=None):
        """`group_priors` (n_groups, 2) = hydra's --groupPriorsFile, `dirichlet_priors` (n_groups, n_mix) = --dPriorsFile;
        `fh` = dict(v0L, v0t, v0c, s02c, tau0[, state0]) (missing keys: src/options.hpp:91-96) runs --bayesType bayesFHMPI."""
        self.store = store
        self._lib = store._lib
        G, K = store.n_groups, store.n_mix
        mS = np.asarray(mS, dtype=np.float64).reshape(G, -1)
        if mS.shape[1] == K - 1:  # the zero component is implicit in .mS files (src/data.cpp:1981-2007)
            mS = np.concatenate([np.zeros((G, 1)), mS], axis=1)
        assert mS.shape == (G, K), (mS.shape, G, K)
        self.mS = np.ascontiguousarray(mS)
        y = arr(y, np.float64)
        assert y.shape == (store.n_ind,)
        g = arr(groups, np.int32)
        s0 = arr(sigmaG0, np.float64)
        check(self._lib.hb_brr_init(store._h, ptr(y), ptr(g), ptr(self.mS), ptr(s0), C.c_uint32(seed & 0xFFFFFFFF)))
        self.iteration_index = 0
        self.n_cov = 0
        if covariates is not None:   # hydra's --covariates (src/BayesRRm.cpp:2648-2681): X (n_ind, n_cov), used as given
            X = arr(covariates, np.float64)
            assert X.ndim == 2 and X.shape[0] == store.n_ind, X.shape
            check(self._lib.hb_brr_set_covariates(store._h, ptr(X), C.c_uint32(X.shape[1])))
            self.n_cov = X.shape[1]
        if group_priors is not None or dirichlet_priors is not None:
            gp = None if group_priors is None else arr(np.asarray(group_priors, np.float64).reshape(G, 2), np.float64)
            dp = None if dirichlet_priors is None else arr(np.asarray(dirichlet_priors, np.float64).reshape(G, K), np.float64)
            check(self._lib.hb_brr_set_group_priors(store._h, ptr(gp), ptr(dp)))
        self.fh = None
        if fh is not None:
            self.fh = dict(FH_DEFAULTS, **{k: float(fh[k]) for k in FH_DEFAULTS if k in fh})
            cfg = capi.HbFhConfig(**self.fh)
            s0 = None if fh.get("state0") is None else arr(np.asarray(fh["state0"], np.float64).reshape(2 + G), np.float64)
            check(self._lib.hb_brr_set_fh(store._h, C.byref(cfg), ptr(s0)))

    def fh_state(self):
        """bayesFH: dict(hypTau, tau, scaledBSQN, c_slab, lambda_var, nu_var) after the last iteration."""
        s = self.store
        sc, cs, lam, nu = np.zeros(3), np.zeros(s.n_groups), np.zeros(s.m_local), np.zeros(s.m_local)
        check(self._lib.hb_brr_get_fh(s._h, ptr(sc), ptr(cs), ptr(lam), ptr(nu)))
        return dict(hypTau=sc[0], tau=sc[1], scaledBSQN=sc[2], c_slab=cs, lambda_var=lam, nu_var=nu)

    def iteration(self, tape=None):
        """One Gibbs iteration. tape = dict(zmu, perm, u, z[, sigmaG, pi, sigmaE][, gnu, glam, fh_hyper]) for deterministic replay."""
        out = capi.HbBrrIterOut()
        keep = []
        tp = None
        if tape is not None:
            t = capi.HbBrrTape()
            for name, dt in (("zmu", np.float64), ("perm", np.int32), ("u", np.float64), ("z", np.float64),
                             ("sigmaG", np.float64), ("pi", np.float64), ("sigmaE", np.float64), ("xI", np.int32), ("zcov", np.float64),
                             ("gnu", np.float64), ("glam", np.float64), ("fh_hyper", np.float64)):
                v = tape.get(name)
                if v is not None:
                    a = arr(np.atleast_1d(v), dt)
                    keep.append(a)
                    setattr(t, name, a.ctypes.data)
            tp = C.byref(t)
        check(self._lib.hb_brr_iteration(self.store._h, tp, C.byref(out)))
        self.iteration_index += 1
        d = {n: getattr(out, n) for n, _ in out._fields_}
        d["phase_cycles"] = list(out.phase_cycles)
        return d

    def restore_outputs(self, iterations_done, sigmaG, pi, sigmaE, mu_tasks, beta, components, eps_task0, perm=None):
        """The reference's restart from its output files (src/BayesRRm.cpp:842-928): see hb_brr_restore_outputs."""
        s = self.store
        a = [arr(sigmaG, np.float64), arr(pi, np.float64), arr(np.atleast_1d(mu_tasks), np.float64), arr(beta, np.float64),
             arr(components, np.int32), arr(eps_task0, np.float64)]
        p = arr(perm, np.int32) if perm is not None else None
        check(self._lib.hb_brr_restore_outputs(s._h, C.c_uint32(iterations_done), ptr(a[0]), ptr(a[1]), C.c_double(float(sigmaE)), ptr(a[2]),
                                               ptr(a[3]), ptr(a[4]), ptr(a[5]), ptr(p) if p is not None else None))
        self.iteration_index = iterations_done

    def hyper(self):
        s = self.store
        G, K = s.n_groups, s.n_mix
        sigmaG, pi, sigmaE = np.zeros(G), np.zeros((G, K)), C.c_double()
        mu, bsq, cass, m0 = np.zeros(s.tasks_local), np.zeros(G), np.zeros((G, K), np.int32), np.zeros(G, np.int32)
        check(self._lib.hb_brr_get_hyper(s._h, ptr(sigmaG), ptr(pi), C.byref(sigmaE), ptr(mu), ptr(bsq), ptr(cass), ptr(m0)))
        return dict(sigmaG=sigmaG, pi=pi, sigmaE=sigmaE.value, mu=mu, bsq=bsq, cass=cass, m0=m0)

    def state(self, out=None):
        """beta, components, Acum of the local markers. `out` = preallocated (beta f64, components i32, acum f64) arrays,
        e.g. views of pinned host memory, to make the device-to-host copies asynchronous DMA transfers."""
        s = self.store
        if out is None:
            out = (np.zeros(s.m_local), np.zeros(s.m_local, np.int32), np.zeros(s.m_local))
        beta, comp, acum = out
        assert beta.dtype == np.float64 and comp.dtype == np.int32 and acum.dtype == np.float64
        assert beta.flags.c_contiguous and comp.flags.c_contiguous and acum.flags.c_contiguous and len(beta) == len(comp) == len(acum) == s.m_local
        check(self._lib.hb_brr_get_state(s._h, ptr(beta), ptr(comp), ptr(acum)))
        return beta, comp, acum

    def state_async(self, out):
        """As state(out=...), but returns at once: the copies run on a second stream while the next iteration() computes. `out`
        (page-locked arrays for a true DMA) must not be touched until state_wait() has returned."""
        s = self.store
        beta, comp, acum = out
        assert beta.dtype == np.float64 and comp.dtype == np.int32 and acum.dtype == np.float64
        assert beta.flags.c_contiguous and comp.flags.c_contiguous and acum.flags.c_contiguous and len(beta) == len(comp) == len(acum) == s.m_local
        self._async_keep = out
        check(self._lib.hb_brr_get_state_async(s._h, ptr(beta), ptr(comp), ptr(acum)))

    def state_wait(self):
        check(self._lib.hb_brr_state_wait(self.store._h))
        out, self._async_keep = getattr(self, "_async_keep", None), None
        return out

    def gamma(self):
        """Fixed effects and their current order (the reference's .gam / .xiv files)."""
        g, x = np.zeros(self.n_cov), np.zeros(self.n_cov, np.int32)
        if self.n_cov:
            check(self._lib.hb_brr_get_gamma(self.store._h, ptr(g), ptr(x)))
        return g, x

    def set_state(self, beta=None, components=None):
        check(self._lib.hb_brr_set_state(self.store._h, ptr(arr(beta, np.float64)), ptr(arr(components, np.int32))))

    def save_state(self):
        """The complete chain state of this GPU (hydra --restart); bytes for load_state of a BayesRRm built on the same inputs."""
        need = C.c_size_t(0)
        check(self._lib.hb_brr_save_state(self.store._h, None, C.c_size_t(0), C.byref(need)))
        buf = (C.c_ubyte * need.value)()
        check(self._lib.hb_brr_save_state(self.store._h, buf, C.c_size_t(need.value), C.byref(need)))
        return bytes(buf)

    def load_state(self, blob):
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        check(self._lib.hb_brr_load_state(self.store._h, buf, C.c_size_t(len(blob))))

    def task_epsilon(self, task_local=0):
        e = np.zeros(self.store.n_ind)
        check(self._lib.hb_brr_get_task_epsilon(self.store._h, C.c_uint32(task_local), ptr(e)))
        return e

    def task_perm(self, task_local=0):
        _, l = self.store.task_blocks()
        p = np.zeros(int(l[self.store.task_first + task_local]), np.int32)
        check(self._lib.hb_brr_get_task_perm(self.store._h, C.c_uint32(task_local), ptr(p)))
        return p


class BayesW:
    """The BayesW (Weibull survival) chain on one GPU (src/BayesW.cpp:905-1907). `store` must be created with model="bayesW"."""

    def __init__(self, store: GenotypeStore, y, failure, mS, groups=None, quad_points=25, seed=0, covariates=None):
        self.store = store
        self.n_cov = 0
        self._lib = store._lib
        G, K = store.n_groups, store.n_mix
        mS = np.asarray(mS, dtype=np.float64).reshape(G, -1)
        if mS.shape[1] == K - 1:
            mS = np.concatenate([np.zeros((G, 1)), mS], axis=1)
        assert mS.shape == (G, K), (mS.shape, G, K)
        self.mS = np.ascontiguousarray(mS)
        y = arr(y, np.float64)
        f = arr(failure, np.float64)
        assert y.shape == (store.n_ind,) and f.shape == (store.n_ind,)
        g = arr(groups, np.int32)
        check(self._lib.hb_bw_init(store._h, ptr(y), ptr(f), ptr(g), ptr(self.mS), C.c_uint32(quad_points), C.c_uint32(seed & 0xFFFFFFFF)))
        if covariates is not None:   # fixed effects, N x F as read from the covariate file (src/BayesW.cpp:1366-1413)
            X = np.ascontiguousarray(np.asarray(covariates, np.float64).reshape(store.n_ind, -1))
            self.n_cov = X.shape[1]
            check(self._lib.hb_bw_set_covariates(store._h, ptr(X), C.c_uint32(self.n_cov)))

    def gamma(self):
        g, xI = np.zeros(self.n_cov), np.zeros(self.n_cov, np.int32)
        check(self._lib.hb_bw_get_gamma(self.store._h, ptr(g), ptr(xI)))
        return g, xI

    def iteration(self, tape=None):
        out = capi.HbBwIterOut()
        keep = []
        tp = None
        if tape is not None:
            t = capi.HbBwTape()
            for name, dt in (("perm", np.int32), ("p", np.float64), ("sigmaG", np.float64), ("pi", np.float64), ("xI", np.int32)):
                v = tape.get(name)
                if v is not None:
                    a = arr(np.atleast_1d(v), dt)
                    keep.append(a)
                    setattr(t, name, a.ctypes.data)
            tp = C.byref(t)
        check(self._lib.hb_bw_iteration(self.store._h, tp, C.byref(out)))
        return {n: getattr(out, n) for n, _ in out._fields_}

    def hyper(self):
        s = self.store
        G, K = s.n_groups, s.n_mix
        sigmaG, pi, mu, alpha = np.zeros(G), np.zeros((G, K)), C.c_double(), C.c_double()
        bsq, cass, m0 = np.zeros(G), np.zeros((G, K), np.int32), np.zeros(G, np.int32)
        check(self._lib.hb_bw_get_hyper(s._h, ptr(sigmaG), ptr(pi), C.byref(mu), C.byref(alpha), ptr(bsq), ptr(cass), ptr(m0)))
        return dict(sigmaG=sigmaG, pi=pi, mu=mu.value, alpha=alpha.value, bsq=bsq, cass=cass, m0=m0)

    def state(self):
        s = self.store
        beta, comp = np.zeros(s.m_local), np.zeros(s.m_local, np.int32)
        check(self._lib.hb_brr_get_state(s._h, ptr(beta), ptr(comp), None))
        return beta, comp

    def epsilon(self):
        return self.store.get_epsilon()

    def marker_stats(self):
        s = self.store
        sd, sf = np.zeros(s.m_local), np.zeros(s.m_local)
        check(self._lib.hb_bw_marker_stats(s._h, ptr(sd), ptr(sf)))
        return sd, sf


def bw_vi_sums(store, markers, beta_old, alpha):
    m = arr(markers, np.uint32)
    b = arr(beta_old, np.float64)
    out = np.zeros((len(m), 4))
    check(store._lib.hb_bw_vi_sums(store._h, ptr(m), ptr(b), C.c_uint32(len(m)), C.c_double(alpha), ptr(out)))
    return out


def bw_sum_exp(store, a, b):
    o = C.c_double()
    check(store._lib.hb_bw_sum_exp(store._h, C.c_double(a), C.c_double(b), C.byref(o)))
    return o.value


def bw_marginal_likelihoods(store, quad_points, pars, prior, cVa):
    pars, prior, cVa = arr(pars, np.float64), arr(prior, np.float64), arr(cVa, np.float64)
    post = np.zeros(len(prior))
    check(store._lib.hb_bw_marginal_likelihoods(store._h, C.c_uint32(quad_points), ptr(pars), ptr(prior), ptr(cVa), C.c_uint32(len(cVa)), ptr(post)))
    return post


def bw_arms_beta(store, pars, C_k, sum_sigmaG, beta_old, seed, task, iteration, j):
    pars = arr(pars, np.float64)
    out = np.zeros(4)
    check(store._lib.hb_bw_arms_beta(store._h, ptr(pars), C.c_double(C_k), C.c_double(sum_sigmaG), C.c_double(beta_old), C.c_uint32(seed),
                                     C.c_uint32(task), C.c_uint32(iteration), C.c_uint32(j), ptr(out)))
    return dict(beta=out[0], err=int(out[1]), neval=int(out[2]), nrand=int(out[3]))
