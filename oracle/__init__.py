"""ctypes front-end of the CPU ORACLE (oracle/hydra_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline / --impl reference legs.  The product package
(hydra_b200/) must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, C.CDLL] = {}

c_sz = C.c_size_t
P = C.POINTER


def build(fast: bool = False, force: bool = False) -> str:
    name = "libhydra_oracle_fast.so" if fast else "libhydra_oracle.so"
    path = os.path.join(_HERE, name)
    args = ["make", "-C", _HERE, name] + (["-B"] if force else [])  # make tracks the sources
    subprocess.run(args, check=True, capture_output=True)
    return path


def lib(fast: bool = False, rebuild: bool = False) -> C.CDLL:
    key = "fast" if fast else "strict"
    if key not in _LIBS or rebuild:
        L = C.CDLL(build(fast, force=rebuild))
        L.ho_sparse_dotprod.restype = C.c_double
        L.ho_lut_dotprod.restype = C.c_double
        L.ho_mt_res53.restype = C.c_double
        L.ho_mt_normal.restype = C.c_double
        L.ho_mt_gamma.restype = C.c_double
        L.ho_mt_gamma.argtypes = [C.c_void_p, C.c_double]
        L.ho_mt_u32.restype = C.c_uint32
        L.ho_bw_beta_dens.restype = C.c_double
        L.ho_bw_gh_integral.restype = C.c_double
        L.ho_bw_mu_dens.restype = C.c_double
        L.ho_bw_alpha_dens.restype = C.c_double
        _LIBS[key] = L
    return _LIBS[key]


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# --------------------------------------------------------------------------- LUT
def lut_build():
    a = np.empty(1024, np.float64)
    b = np.empty(1024, np.float64)
    lib().ho_lut_build(_p(a), _p(b))
    return a, b


# --------------------------------------------------------------------------- staging
@dataclass
class SparseLists:
    """Reference sparse representation (src/data.cpp:1224-1290)."""

    I1: np.ndarray
    N1S: np.ndarray
    N1L: np.ndarray
    I2: np.ndarray
    N2S: np.ndarray
    N2L: np.ndarray
    IM: np.ndarray
    NMS: np.ndarray
    NML: np.ndarray


def snp_len_byt(n: int) -> int:
    return (n + 3) // 4


def sparse_fill_indices(bed: np.ndarray, nind: int) -> SparseLists:
    """bed: (M, NB) uint8 column-major PLINK bytes (header stripped)."""
    bed = _c(bed, np.uint8)
    M, NB = bed.shape
    n1 = c_sz()
    n2 = c_sz()
    nm = c_sz()
    L = lib()
    L.ho_sparse_get_sizes_from_raw(_p(bed), C.c_uint(M), C.c_uint(NB), C.c_uint(nind), C.byref(n1), C.byref(n2), C.byref(nm))
    I1 = np.empty(max(n1.value, 1), np.uint32)
    I2 = np.empty(max(n2.value, 1), np.uint32)
    IM = np.empty(max(nm.value, 1), np.uint32)
    S = [np.zeros(M, np.uint64) for _ in range(6)]
    L.ho_sparse_fill_indices(_p(bed), C.c_uint(M), C.c_uint(NB), C.c_uint(nind),
                             _p(S[0]), _p(S[1]), _p(I1), _p(S[2]), _p(S[3]), _p(I2), _p(S[4]), _p(S[5]), _p(IM))
    return SparseLists(I1[: n1.value], S[0], S[1], I2[: n2.value], S[2], S[3], IM[: nm.value], S[4], S[5])


def correct_for_missing_phenotype(sp: SparseLists, na_inds: np.ndarray, usebed=None) -> None:
    na = _c(na_inds, np.uint32)
    M = len(sp.N1S)
    ub = None if usebed is None else _c(usebed, np.uint8)
    L = lib()
    for (S, Ln, I) in ((sp.N1S, sp.N1L, sp.I1), (sp.N2S, sp.N2L, sp.I2), (sp.NMS, sp.NML, sp.IM)):
        L.ho_sparse_correct_for_missing_phenotype(_p(S), _p(Ln), _p(I), C.c_int(M), _p(ub), _p(na), C.c_int(len(na)))


def bed_marker_from_sparse(nbytes: int, i1, i2, im) -> np.ndarray:
    out = np.empty(nbytes, np.uint8)
    i1 = _c(i1, np.uint32)
    i2 = _c(i2, np.uint32)
    im = _c(im, np.uint32)
    lib().ho_bed_marker_from_sparse(_p(out), c_sz(nbytes), _p(i1), c_sz(len(i1)), _p(i2), c_sz(len(i2)), _p(im), c_sz(len(im)))
    return out


def marker_stats_brr(N: int, n1l, n2l, nml):
    n1l = _c(n1l, np.uint64)
    n2l = _c(n2l, np.uint64)
    nml = _c(nml, np.uint64)
    M = len(n1l)
    mave = np.empty(M)
    mstd = np.empty(M)
    lib().ho_marker_stats_brr(C.c_int(M), C.c_uint(N), _p(n1l), _p(n2l), _p(nml), _p(mave), _p(mstd))
    return mave, mstd


def center_and_scale(y):
    y = np.array(y, dtype=np.float64, copy=True)
    lib().ho_center_and_scale(_p(y), C.c_int(len(y)))
    return y


# --------------------------------------------------------------------------- kernels
def sparse_dotprod(eps, sp: SparseLists, m: int, mave: float, mstd: float, fast=False) -> float:
    eps = _c(eps, np.float64)
    return lib(fast).ho_sparse_dotprod(
        _p(eps), _p(sp.I1), c_sz(int(sp.N1S[m])), c_sz(int(sp.N1L[m])), _p(sp.I2), c_sz(int(sp.N2S[m])), c_sz(int(sp.N2L[m])),
        _p(sp.IM), c_sz(int(sp.NMS[m])), c_sz(int(sp.NML[m])), C.c_double(mave), C.c_double(mstd), C.c_int(len(eps)))


def lut_dotprod(raw, eps, mave: float, mstd: float) -> float:
    raw = _c(raw, np.uint8)
    eps = _c(eps, np.float64)
    return lib().ho_lut_dotprod(_p(raw), _p(eps), C.c_int(len(eps)), C.c_double(mave), C.c_double(mstd))


def sparse_scaadd(N: int, dmult: float, sp: SparseLists, m: int, mave: float, mstd: float):
    out = np.empty(N)
    lib().ho_sparse_scaadd(
        _p(out), C.c_double(dmult), _p(sp.I1), c_sz(int(sp.N1S[m])), c_sz(int(sp.N1L[m])), _p(sp.I2), c_sz(int(sp.N2S[m])),
        c_sz(int(sp.N2L[m])), _p(sp.IM), c_sz(int(sp.NMS[m])), c_sz(int(sp.NML[m])), C.c_double(mave), C.c_double(mstd), C.c_int(N))
    return out


def lut_scaadd(N: int, raw, dbeta: float, mave: float, mstd: float):
    raw = _c(raw, np.uint8)
    out = np.empty(N)
    lib().ho_lut_scaadd(_p(out), _p(raw), C.c_double(dbeta), C.c_double(mave), C.c_double(mstd), C.c_int(N))
    return out


def define_blocks(Mtot: int, T: int):
    s = np.zeros(T, np.int32)
    l = np.zeros(T, np.int32)
    lib().ho_define_blocks_of_markers(C.c_int(Mtot), _p(s), _p(l), C.c_uint(T))
    return s, l


# --------------------------------------------------------------------------- RNG spec v1
class MT:
    def __init__(self, seed: int):
        self.buf = C.create_string_buffer(lib().ho_sizeof_mt())
        lib().ho_mt_seed(self.buf, C.c_uint32(seed & 0xFFFFFFFF))

    def u32(self):
        return lib().ho_mt_u32(self.buf)

    def res53(self):
        return lib().ho_mt_res53(self.buf)

    def normal(self):
        return lib().ho_mt_normal(self.buf)

    def gamma(self, a):
        return lib().ho_mt_gamma(self.buf, C.c_double(a))


def philox4x32(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().ho_philox4x32(c, k, o)
    return list(o)


def marker_draws(seed, task, iteration, j):
    u = C.c_double()
    z = C.c_double()
    lib().ho_marker_draws(C.c_uint32(seed), C.c_uint32(task), C.c_uint32(iteration), C.c_uint32(j), C.byref(u), C.byref(z))
    return u.value, z.value


class TapeMaker:
    """Positional draw tape of RNG spec v1."""

    def __init__(self, seed: int, T: int, Mtot: int, shuffle: bool = True):
        self.seed, self.T, self.Mtot, self.shuffle = seed, T, Mtot, shuffle
        sz = lib().ho_sizeof_mt()
        self.streams = C.create_string_buffer(sz * T)
        for r in range(T):
            lib().ho_mt_seed(C.byref(self.streams, sz * r), C.c_uint32((seed + 1000 * r) & 0xFFFFFFFF))
        s, l = define_blocks(Mtot, T)
        self.perm_state = np.concatenate([np.arange(n, dtype=np.int32) for n in l]) if Mtot else np.zeros(0, np.int32)
        self.it = 0

    def next(self):
        T, M = self.T, self.Mtot
        zmu = np.empty(T)
        perm = np.empty(M, np.int32)
        u = np.empty(M)
        z = np.empty(M)
        lib().ho_tape_iteration(self.streams, C.c_uint32(self.seed), C.c_int(T), C.c_int(M), C.c_uint32(self.it),
                                _p(self.perm_state), _p(zmu), _p(perm), _p(u), _p(z), C.c_int(1 if self.shuffle else 0))
        self.it += 1
        return zmu, perm, u, z

    def make(self, n_iter: int):
        zs, ps, us, zz = [], [], [], []
        for _ in range(n_iter):
            a, b, c, d = self.next()
            zs.append(a), ps.append(b), us.append(c), zz.append(d)
        return dict(zmu=np.array(zs), perm=np.array(ps, dtype=np.int32), u=np.array(us), z=np.array(zz))


# --------------------------------------------------------------------------- synthetic data
def synth_thresholds(p: float, pmiss: float = 0.001):
    t = (C.c_uint32 * 3)()
    lib().ho_synth_thresholds(C.c_double(p), C.c_double(pmiss), t)
    return np.array(list(t), dtype=np.uint32)


def synth_bed(seed: int, N: int, thresholds: np.ndarray, attempts=None, j0: int = 0, fast=False) -> np.ndarray:
    """thresholds: (M,3) uint32 -> (M, ceil(N/4)) uint8 BED columns."""
    thresholds = _c(thresholds, np.uint32)
    M = thresholds.shape[0]
    out = np.empty((M, snp_len_byt(N)), np.uint8)
    L = lib(fast)
    for j in range(M):
        att = 0 if attempts is None else int(attempts[j])
        L.ho_synth_bed_marker(C.c_uint32(seed), C.c_uint32(j0 + j), C.c_uint32(att), _p(thresholds[j]), C.c_uint32(N), _p(out[j]))
    return out


# --------------------------------------------------------------------------- chain
class _BrrArgs(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in ("N", "Mtot", "T", "K", "G", "sync_rate", "n_iter", "iter0")]
        + [(n, C.c_void_p) for n in ("I1", "N1S", "N1L", "I2", "N2S", "N2L", "IM", "NMS", "NML", "usebed", "bed")]
        + [("snpLenByt", c_sz)]
        + [(n, C.c_void_p) for n in ("y_raw", "groups", "mS", "tape_zmu", "tape_perm", "tape_u", "tape_z",
                                     "tape_sigmaG0", "tape_sigmaG", "tape_sigmaE", "tape_pi")]
        + [("hyper_seed", C.c_uint32)]
        + [(n, C.c_void_p) for n in ("out_beta", "out_comp", "out_acum", "out_mu", "out_eps", "out_sigmaG", "out_sigmaE",
                                     "out_pi", "out_bsq", "out_cass", "out_esqn", "out_epssum", "out_nsync", "out_loop_seconds")]
        + [("X", C.c_void_p), ("F", C.c_int32), ("tape_xI", C.c_void_p), ("tape_zcov", C.c_void_p), ("out_gamma", C.c_void_p)]
        + [("group_priors", C.c_void_p), ("dirichlet", C.c_void_p), ("fh", C.c_int32)]
        + [(n, C.c_double) for n in ("fh_v0L", "fh_v0t", "fh_v0c", "fh_s02c", "fh_tau0")]
        + [(n, C.c_void_p) for n in ("fh_state0", "tape_gnu", "tape_glam", "tape_fh_hyper")]
        + [("fh_seed", C.c_uint32)]
        + [(n, C.c_void_p) for n in ("out_fh", "out_lambda", "out_nu")]
    )


FH_DEFAULTS = dict(v0L=3.0, v0t=3.0, v0c=3.0, s02c=1.0, tau0=1.0)   # src/options.hpp:91-96


def fh_gamma(seed, marker, iteration, which, a):
    """Standard Gamma(a, 1) variate behind the nu_var (which=0) / lambda_var (which=1) draw of a marker (RNG spec v1)."""
    f = lib().ho_fh_gamma
    f.restype = C.c_double
    return f(C.c_uint32(seed), C.c_uint32(marker), C.c_uint32(iteration), C.c_int(which), C.c_double(a))


def brr_chain(N, Mtot, T, K, G, sync_rate, n_iter, sp: SparseLists, y_raw, groups, mS, tape, sigmaG0,
              usebed=None, bed=None, hyper=None, hyper_seed=0, want_eps=True, fast=False, want_marker_out=True, covariates=None,
              fh=None, group_priors=None, dirichlet=None):
    """Run the BayesRRm oracle chain. `tape` = dict(zmu, perm, u, z); `hyper` =
    dict(sigmaG, sigmaE, pi) to replay hyper-parameter VALUES, or None to draw
    them with mt19937(hyper_seed) (RNG spec v1).  Returns dict of per-iteration outputs.
    `fh` = dict(v0L, v0t, v0c, s02c, tau0[, state0 (2+G), seed]) switches bayesFHMPI on; the tape may then carry gnu / glam
    (n_iter, Mtot) and fh_hyper (n_iter, G, 3). `group_priors` (G, 2) and `dirichlet` (G, K) are the two prior files."""
    keep = []

    def k(a, dt):
        if a is None:
            return None
        a = _c(a, dt)
        keep.append(a)
        return a.ctypes.data

    out = dict(
        beta=np.zeros((n_iter, Mtot)) if want_marker_out else None,
        comp=np.zeros((n_iter, Mtot), np.int32) if want_marker_out else None,
        acum=np.zeros((n_iter, Mtot)) if want_marker_out else None,
        mu=np.zeros((n_iter, T)),
        eps=np.zeros((n_iter, T, N)) if want_eps else None,
        sigmaG=np.zeros((n_iter, G)), sigmaE=np.zeros(n_iter), pi=np.zeros((n_iter, G, K)),
        bsq=np.zeros((n_iter, G)), cass=np.zeros((n_iter, G, K), np.int32), esqn=np.zeros(n_iter),
        epssum=np.zeros((n_iter, T)), nsync=np.zeros(n_iter, np.int64), loop_seconds=np.zeros(n_iter),
    )
    a = _BrrArgs()
    a.N, a.Mtot, a.T, a.K, a.G, a.sync_rate, a.n_iter, a.iter0 = N, Mtot, T, K, G, sync_rate, n_iter, 0
    a.I1, a.N1S, a.N1L = k(sp.I1 if len(sp.I1) else np.zeros(1), np.uint32), k(sp.N1S, np.uint64), k(sp.N1L, np.uint64)
    a.I2, a.N2S, a.N2L = k(sp.I2 if len(sp.I2) else np.zeros(1), np.uint32), k(sp.N2S, np.uint64), k(sp.N2L, np.uint64)
    a.IM, a.NMS, a.NML = k(sp.IM if len(sp.IM) else np.zeros(1), np.uint32), k(sp.NMS, np.uint64), k(sp.NML, np.uint64)
    a.usebed = k(usebed, np.uint8)
    a.bed = k(bed, np.uint8)
    a.snpLenByt = snp_len_byt(N)
    a.y_raw, a.groups, a.mS = k(y_raw, np.float64), k(groups, np.int32), k(mS, np.float64)
    a.tape_zmu, a.tape_perm = k(tape["zmu"], np.float64), k(tape["perm"], np.int32)
    a.tape_u, a.tape_z = k(tape["u"], np.float64), k(tape["z"], np.float64)
    a.tape_sigmaG0 = k(sigmaG0, np.float64)
    if hyper is not None:
        a.tape_sigmaG, a.tape_sigmaE, a.tape_pi = k(hyper["sigmaG"], np.float64), k(hyper["sigmaE"], np.float64), k(hyper["pi"], np.float64)
    a.hyper_seed = hyper_seed & 0xFFFFFFFF
    if covariates is not None:   # fixed effects (src/BayesRRm.cpp:2648-2681): X (N, F); tape may carry xI (n_iter, F) and zcov (n_iter, F)
        Xc = np.ascontiguousarray(covariates, dtype=np.float64)
        assert Xc.shape[0] == N
        a.X, a.F = k(Xc, np.float64), Xc.shape[1]
        a.tape_xI, a.tape_zcov = k(tape.get("xI"), np.int32), k(tape.get("zcov"), np.float64)
        out["gamma"] = np.zeros((n_iter, Xc.shape[1]))
        a.out_gamma = out["gamma"].ctypes.data
    a.group_priors, a.dirichlet = k(group_priors, np.float64), k(dirichlet, np.float64)
    if fh is not None:
        p = dict(FH_DEFAULTS, **{q: fh[q] for q in FH_DEFAULTS if q in fh})
        a.fh = 1
        a.fh_v0L, a.fh_v0t, a.fh_v0c, a.fh_s02c, a.fh_tau0 = p["v0L"], p["v0t"], p["v0c"], p["s02c"], p["tau0"]
        a.fh_state0 = k(fh.get("state0"), np.float64)
        a.fh_seed = int(fh.get("seed", 0)) & 0xFFFFFFFF
        a.tape_gnu, a.tape_glam, a.tape_fh_hyper = k(tape.get("gnu"), np.float64), k(tape.get("glam"), np.float64), k(tape.get("fh_hyper"), np.float64)
        out["fh"] = np.zeros((n_iter, 3 + G))
        out["lambda"] = np.zeros((n_iter, Mtot))
        out["nu"] = np.zeros((n_iter, Mtot))
        a.out_fh, a.out_lambda, a.out_nu = out["fh"].ctypes.data, out["lambda"].ctypes.data, out["nu"].ctypes.data
    for name in ("beta", "comp", "acum", "mu", "eps", "sigmaG", "sigmaE", "pi", "bsq", "cass", "esqn", "epssum", "nsync", "loop_seconds"):
        v = out[name]
        setattr(a, "out_" + name, None if v is None else v.ctypes.data)
    rc = lib(fast).ho_brr_chain(C.byref(a))
    if rc != 0:
        raise RuntimeError(f"ho_brr_chain failed rc={rc}")
    return out


def num_threads(fast=True):
    return lib(fast).ho_num_threads()


def set_num_threads(n, fast=True):
    """OpenMP threads of the timing build (torchrun exports OMP_NUM_THREADS=1; the CPU baseline must not inherit that)."""
    lib(fast).ho_set_num_threads(C.c_int(int(n)))


# --------------------------------------------------------------------------- BayesW
_ARMS = None


def arms_ref():
    """The reference's own ARMS object code (oracle/_ref/libarms_ref.so, built by oracle/build_ref.sh from
    /root/reference/src/BayesW_arms.cpp with rand() wrapped). None if it was never built."""
    global _ARMS
    if _ARMS is None:
        p = os.path.join(_HERE, "_ref", "libarms_ref.so")
        if not os.path.exists(p):
            return None
        _ARMS = C.CDLL(p)
    return _ARMS


class _BwPars(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("alpha", "sigmaG", "sum_failure", "vi_sum", "vi_0", "vi_1", "vi_2", "mean", "sd", "mean_sd_ratio", "mixture_value")]


def bw_pars(pars10, mixture_value=0.0):
    return _BwPars(*[float(x) for x in pars10], float(mixture_value))


def bw_marginal_likelihoods(quad_points, pars10, prior, cVa):
    prior, cVa = _c(prior, np.float64), _c(cVa, np.float64)
    post = np.zeros(len(prior))
    p = bw_pars(pars10)
    lib().ho_bw_marginal_likelihoods(C.c_int(quad_points), _p(prior), _p(cVa), C.c_int(len(cVa)), C.byref(p), _p(post))
    post[0] = prior[0] * 1.77245385090552
    return post


def bw_sample_beta(pars10, C_k, sum_sigmaG, beta_old, seed, task, iteration, j):
    a = arms_ref()
    p = bw_pars(pars10, C_k)
    bn, ne, nr = C.c_double(), C.c_int(), C.c_int()
    fn = C.cast(a.ho_ref_arms, C.c_void_p)
    err = lib().ho_bw_sample_beta(fn, C.byref(p), C.c_double(sum_sigmaG), C.c_double(beta_old), C.c_uint32(seed), C.c_uint32(task),
                                  C.c_uint32(iteration), C.c_uint32(j), C.byref(bn), C.byref(ne), C.byref(nr))
    return dict(beta=bn.value, err=err, neval=ne.value, nrand=nr.value)


def bw_marker_stats(N, sp: SparseLists, fail):
    M = len(sp.N1S)
    fail = _c(fail, np.float64)
    mave, mstd, sf = np.zeros(M), np.zeros(M), np.zeros(M)
    lib().ho_bw_marker_stats(C.c_int(M), C.c_uint(N), _p(sp.I1 if len(sp.I1) else np.zeros(1, np.uint32)), _p(sp.N1S), _p(sp.N1L),
                             _p(sp.I2 if len(sp.I2) else np.zeros(1, np.uint32)), _p(sp.N2S), _p(sp.N2L), _p(sp.NML), _p(fail), _p(mave), _p(mstd), _p(sf))
    return mave, mstd, sf


class _BwArgs(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("N", "Mtot", "T", "K", "G", "sync_rate", "n_iter", "quad_points")]
                + [(n, C.c_void_p) for n in ("I1", "N1S", "N1L", "I2", "N2S", "N2L", "IM", "NMS", "NML", "y", "fail", "groups", "mS",
                                             "tape_perm", "tape_p", "tape_sigmaG", "tape_pi")]
                + [("seed", C.c_uint32), ("hyper_seed", C.c_uint32), ("arms", C.c_void_p)]
                + [(n, C.c_void_p) for n in ("out_beta", "out_comp", "out_eps", "out_mu", "out_alpha", "out_sigmaG", "out_pi", "out_bsq",
                                             "out_cass", "out_nsync", "out_err")]
                + [("X", C.c_void_p), ("F", C.c_int32), ("tape_xI", C.c_void_p), ("out_gamma", C.c_void_p)])


def bw_chain(N, Mtot, T, K, G, sync_rate, n_iter, quad_points, sp: SparseLists, y, fail, groups, mS, tape, seed, hyper=None, hyper_seed=0,
             covariates=None):
    """BayesW oracle chain; tape = dict(perm, p[, xI]); hyper = dict(sigmaG, pi) to replay values, else drawn with mt19937(hyper_seed);
    covariates = N x F matrix of fixed effects (src/BayesW.cpp:1366-1413), tape["xI"] = their order per iteration (n_iter x F)."""
    keep = []

    def k(a, dt):
        if a is None:
            return None
        a = _c(a, dt)
        keep.append(a)
        return a.ctypes.data

    out = dict(beta=np.zeros((n_iter, Mtot)), comp=np.zeros((n_iter, Mtot), np.int32), eps=np.zeros((n_iter, N)), mu=np.zeros(n_iter),
               alpha=np.zeros(n_iter), sigmaG=np.zeros((n_iter, G)), pi=np.zeros((n_iter, G, K)), bsq=np.zeros((n_iter, G)),
               cass=np.zeros((n_iter, G, K), np.int32), nsync=np.zeros(n_iter, np.int64), err=np.zeros(1, np.int32))
    a = _BwArgs()
    a.N, a.Mtot, a.T, a.K, a.G, a.sync_rate, a.n_iter, a.quad_points = N, Mtot, T, K, G, sync_rate, n_iter, quad_points
    a.I1, a.N1S, a.N1L = k(sp.I1 if len(sp.I1) else np.zeros(1), np.uint32), k(sp.N1S, np.uint64), k(sp.N1L, np.uint64)
    a.I2, a.N2S, a.N2L = k(sp.I2 if len(sp.I2) else np.zeros(1), np.uint32), k(sp.N2S, np.uint64), k(sp.N2L, np.uint64)
    a.IM, a.NMS, a.NML = k(sp.IM if len(sp.IM) else np.zeros(1), np.uint32), k(sp.NMS, np.uint64), k(sp.NML, np.uint64)
    a.y, a.fail, a.groups, a.mS = k(y, np.float64), k(fail, np.float64), k(groups, np.int32), k(mS, np.float64)
    a.tape_perm, a.tape_p = k(tape["perm"], np.int32), k(tape["p"], np.float64)
    if hyper is not None:
        a.tape_sigmaG, a.tape_pi = k(hyper["sigmaG"], np.float64), k(hyper["pi"], np.float64)
    a.seed, a.hyper_seed = seed & 0xFFFFFFFF, hyper_seed & 0xFFFFFFFF
    if covariates is not None:
        Xc = np.asarray(covariates, np.float64).reshape(N, -1)
        a.F = Xc.shape[1]
        a.X = k(np.asfortranarray(Xc).T.copy(), np.float64)      # column-major: F contiguous columns of N
        a.tape_xI = k(tape.get("xI"), np.int32)
        out["gamma"] = np.zeros((n_iter, a.F))
        a.out_gamma = out["gamma"].ctypes.data
    a.arms = C.cast(arms_ref().ho_ref_arms, C.c_void_p)
    for name in ("beta", "comp", "eps", "mu", "alpha", "sigmaG", "pi", "bsq", "cass", "nsync", "err"):
        setattr(a, "out_" + name, out[name].ctypes.data)
    rc = lib().ho_bw_chain(C.byref(a))
    if rc != 0:
        raise RuntimeError(f"ho_bw_chain: ARMS error code {rc}")
    return out


# --------------------------------------------------------------------------- reference object code (pins)
_REFK = None


def ref_kernels():
    """Function bodies of the reference (src/BayesRRm.cpp:60-413, 1757-1849, 1976-2019; src/data.cpp:826-865, 1112-1290)
    compiled from where they lie by oracle/build_ref.sh into oracle/_ref/libref_kernels.so. None if it was never built."""
    global _REFK
    if _REFK is None:
        p = os.path.join(_HERE, "_ref", "libref_kernels.so")
        if not os.path.exists(p):
            return None
        _REFK = C.CDLL(p)
        for n in ("rk_sum_vector_elements_f64", "rk_sparse_dotprod", "rk_loop_num_bed"):
            getattr(_REFK, n).restype = C.c_double
    return _REFK


def ref_sparse_fill_indices(bed: np.ndarray, nind: int) -> SparseLists:
    """The reference's own Data::sparse_data_get_sizes_from_raw + sparse_data_fill_indices (object code)."""
    bed = _c(bed, np.uint8)
    M, NB = bed.shape
    R = ref_kernels()
    n1, n2, nm = c_sz(), c_sz(), c_sz()
    R.rk_sparse_get_sizes_from_raw(_p(bed), C.c_uint(M), C.c_uint(NB), C.c_uint(nind), C.byref(n1), C.byref(n2), C.byref(nm))
    I1, I2, IM = (np.empty(max(n.value, 1), np.uint32) for n in (n1, n2, nm))
    S = [np.zeros(M, np.uint64) for _ in range(6)]
    R.rk_sparse_fill_indices(_p(bed), C.c_uint(M), C.c_uint(NB), C.c_uint(nind), _p(S[0]), _p(S[1]), _p(I1), _p(S[2]), _p(S[3]), _p(I2),
                             _p(S[4]), _p(S[5]), _p(IM))
    return SparseLists(I1[: n1.value], S[0], S[1], I2[: n2.value], S[2], S[3], IM[: nm.value], S[4], S[5])


def ref_correct_for_missing_phenotype(sp: SparseLists, na_inds) -> None:
    na = _c(na_inds, np.uint32)
    M = len(sp.N1S)
    R = ref_kernels()
    for (S, Ln, I) in ((sp.N1S, sp.N1L, sp.I1), (sp.N2S, sp.N2L, sp.I2), (sp.NMS, sp.NML, sp.IM)):
        R.rk_sparse_correct_for_missing_phenotype(_p(S), _p(Ln), _p(I), C.c_int(M), None, _p(na), C.c_int(len(na)))


def ref_sparse_dotprod(eps, sp: SparseLists, m: int, mave: float, mstd: float) -> float:
    eps = _c(eps, np.float64)
    return ref_kernels().rk_sparse_dotprod(
        _p(eps), _p(sp.I1), c_sz(int(sp.N1S[m])), c_sz(int(sp.N1L[m])), _p(sp.I2), c_sz(int(sp.N2S[m])), c_sz(int(sp.N2L[m])),
        _p(sp.IM), c_sz(int(sp.NMS[m])), c_sz(int(sp.NML[m])), C.c_double(mave), C.c_double(mstd), C.c_int(len(eps)))


def ref_sparse_scaadd(N: int, dmult: float, sp: SparseLists, m: int, mave: float, mstd: float):
    out = np.empty(N)
    ref_kernels().rk_sparse_scaadd(
        _p(out), C.c_double(dmult), _p(sp.I1), c_sz(int(sp.N1S[m])), c_sz(int(sp.N1L[m])), _p(sp.I2), c_sz(int(sp.N2S[m])),
        c_sz(int(sp.N2L[m])), _p(sp.IM), c_sz(int(sp.NMS[m])), c_sz(int(sp.NML[m])), C.c_double(mave), C.c_double(mstd), C.c_int(N))
    return out


def ref_lut_dotprod(raw, eps, mave: float, mstd: float) -> float:
    """The `if (USEBED[marker])` branch of the marker loop, src/BayesRRm.cpp:1757-1840 (object code)."""
    raw, eps = _c(raw, np.uint8), _c(eps, np.float64)
    return ref_kernels().rk_loop_num_bed(_p(raw), _p(eps), C.c_int(len(eps)), C.c_double(mave), C.c_double(mstd))


def ref_lut_scaadd(N: int, raw, dbeta: float, mave: float, mstd: float):
    raw = _c(raw, np.uint8)
    out = np.zeros(N)
    ref_kernels().rk_loop_deltaeps_bed(_p(raw), _p(out), C.c_double(dbeta), C.c_int(N), C.c_double(mave), C.c_double(mstd))
    return out


def ref_bed_marker_from_sparse(Ntot: int, i1, i2, im) -> np.ndarray:
    out = np.empty(snp_len_byt(Ntot), np.uint8)
    i1, i2, im = _c(i1, np.uint32), _c(i2, np.uint32), _c(im, np.uint32)
    ref_kernels().rk_get_bed_marker_from_sparse(_p(out), C.c_int(Ntot), _p(i1), c_sz(len(i1)), _p(i2), c_sz(len(i2)), _p(im), c_sz(len(im)))
    return out
