"""CPU tests of the oracle itself: pinned against the reference's golden LUT tables and against
independent numpy restatements of the integer rules."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from helpers import bed_from_lists, random_bed, reference_lists, simulate_y

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_lut_tables_equal_reference_dotp_lut_h():
    # tests/golden/lut_ref.npz was parsed from the reference's src/dotp_lut.h (make_golden.py)
    g = np.load(os.path.join(GOLD, "lut_ref.npz"))
    a, b = oracle.lut_build()
    assert np.array_equal(a, g["a"].astype(np.float64))
    assert np.array_equal(b, g["b"].astype(np.float64))


def test_survey_worked_example_bed_bytes():
    # N=6, genotypes [0,1,2,NA,1,0] -> bytes 0x4B 0x0E, I1=[1,4], I2=[2], IM=[3]
    bed = np.array([[0x4B, 0x0E]], np.uint8)
    sp = oracle.sparse_fill_indices(bed, 6)
    assert sp.I1.tolist() == [1, 4] and sp.I2.tolist() == [2] and sp.IM.tolist() == [3]
    mave, mstd = oracle.marker_stats_brr(6, sp.N1L, sp.N2L, sp.NML)
    assert mave[0] == pytest.approx(0.8)
    back = oracle.bed_marker_from_sparse(2, sp.I1, sp.I2, sp.IM)
    assert back[0] == 0x4B and (back[1] & 0x0F) == 0x0E


@pytest.mark.parametrize("N", [1, 5, 64, 1003])
def test_fill_indices_matches_dense_decode(N):
    rng = np.random.default_rng(N)
    M = 17
    bed, g = random_bed(rng, M, N, pmiss=0.05)
    sp = oracle.sparse_fill_indices(bed, N)
    for m in range(M):
        for I, S, L, val in ((sp.I1, sp.N1S, sp.N1L, 1), (sp.I2, sp.N2S, sp.N2L, 2), (sp.IM, sp.NMS, sp.NML, -1)):
            assert np.array_equal(I[int(S[m]): int(S[m] + L[m])], np.flatnonzero(g[m] == val))


def test_na_compaction_matches_numpy():
    rng = np.random.default_rng(3)
    N, M = 500, 11
    bed, g = random_bed(rng, M, N, pmiss=0.05)
    na = np.sort(rng.choice(N, 37, replace=False)).astype(np.uint32)
    sp = reference_lists(bed, N, na)
    keep = np.setdiff1d(np.arange(N), na)
    gc = g[:, keep]
    for m in range(M):
        assert np.array_equal(sp.I1[int(sp.N1S[m]): int(sp.N1S[m] + sp.N1L[m])], np.flatnonzero(gc[m] == 1))
        assert np.array_equal(sp.IM[int(sp.NMS[m]): int(sp.NMS[m] + sp.NML[m])], np.flatnonzero(gc[m] == -1))
    # list -> BED -> list round trip
    b2 = bed_from_lists(sp, len(keep))
    sp2 = oracle.sparse_fill_indices(b2, len(keep))
    assert np.array_equal(sp2.N1L, sp.N1L) and np.array_equal(sp2.N2L, sp.N2L) and np.array_equal(sp2.NML, sp.NML)


def test_sparse_and_lut_kernels_agree_with_dense_algebra():
    rng = np.random.default_rng(8)
    N, M = 1001, 9
    bed, g = random_bed(rng, M, N, pmiss=0.03)
    sp = oracle.sparse_fill_indices(bed, N)
    mave, mstd = oracle.marker_stats_brr(N, sp.N1L, sp.N2L, sp.NML)
    eps = rng.normal(size=N)
    for m in range(M):
        x = np.where(g[m] < 0, 0.0, (g[m] - mave[m])) * mstd[m]
        x[g[m] < 0] = 0.0
        d1 = oracle.sparse_dotprod(eps, sp, m, mave[m], mstd[m])
        d2 = oracle.lut_dotprod(bed[m], eps, mave[m], mstd[m])
        assert d1 == pytest.approx(float(x @ eps), rel=1e-10, abs=1e-10)
        assert d2 == pytest.approx(float(x @ eps), rel=1e-10, abs=1e-10)
        np.testing.assert_allclose(oracle.sparse_scaadd(N, 0.37, sp, m, mave[m], mstd[m]), 0.37 * x, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(oracle.lut_scaadd(N, bed[m], 0.37, mave[m], mstd[m]), 0.37 * x, rtol=1e-12, atol=1e-15)
        # standardised column: unit sample variance over the non-missing, as the reference's mstd intends
        assert float(x @ x) == pytest.approx(N - 1, rel=1e-9)


def test_blocks_of_markers():
    s, l = oracle.define_blocks(10, 4)
    assert l.tolist() == [3, 3, 2, 2] and s.tolist() == [0, 3, 6, 8]


def test_mt19937_known_answer():
    # MT19937 reference output for init_genrand(5489): first word 3499211612, 10000th word 4123659995
    mt = oracle.MT(5489)
    first = mt.u32()
    for _ in range(9998):
        mt.u32()
    assert first == 3499211612 and mt.u32() == 4123659995


def test_philox_known_answer():
    # Random123 kat_vectors: philox4x32-10, counter/key all zero and all ones
    assert oracle.philox4x32([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox4x32([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]


def test_chain_is_task_layout_invariant_when_sync_every_marker_single_task():
    """Sanity of the chain restatement: with 1 task and sync_rate 1 the residual always equals y - X beta - mu."""
    rng = np.random.default_rng(4)
    N, M, K = 400, 60, 4
    bed, g = random_bed(rng, M, N, pmiss=0.0)
    sp = oracle.sparse_fill_indices(bed, N)
    y = simulate_y(rng, g)
    mS = np.array([[0.0, 0.001, 0.01, 0.1]])
    tape = oracle.TapeMaker(1, 1, M).make(3)
    out = oracle.brr_chain(N, M, 1, K, 1, 1, 3, sp, y, np.zeros(M, np.int32), mS, tape, np.array([0.5]), hyper_seed=9)
    mave, mstd = oracle.marker_stats_brr(N, sp.N1L, sp.N2L, sp.NML)
    X = (g - mave[:, None]) * mstd[:, None]
    ys = oracle.center_and_scale(y)
    for it in range(3):
        want = ys - X.T @ out["beta"][it] - out["mu"][it, 0]
        np.testing.assert_allclose(out["eps"][it, 0], want, atol=1e-9)
    # BED route and list route of the oracle give the same chain
    out2 = oracle.brr_chain(N, M, 1, K, 1, 1, 3, sp, y, np.zeros(M, np.int32), mS, tape, np.array([0.5]), hyper_seed=9,
                            usebed=np.ones(M, np.uint8), bed=bed)
    assert np.array_equal(out2["comp"], out["comp"])
    np.testing.assert_allclose(out2["beta"], out["beta"], rtol=1e-9, atol=1e-14)


def test_oracle_fixed_effects_recover_simulated_gamma():
    """--covariates (src/BayesRRm.cpp:2648-2681): with standardised covariates the posterior mean of gamma is the simulated
    fixed effect on the scale of the standardised phenotype."""
    rng = np.random.default_rng(1)
    N, M, T, K = 800, 60, 2, 4
    bed, g = random_bed(rng, M, N)
    sp = reference_lists(bed, N)
    X = rng.normal(size=(N, 3))
    X = (X - X.mean(0)) / X.std(0, ddof=1)
    y = simulate_y(rng, g, n_causal=5) + X @ np.array([0.5, -0.3, 0.0])
    sd = (y - y.mean()).std(ddof=1)
    tape = oracle.TapeMaker(3, T, M).make(60)
    ref = oracle.brr_chain(N, M, T, K, 1, 3, 60, sp, y, np.zeros(M, np.int32), np.array([[0, .001, .01, .1]]), tape, np.array([.5]),
                           hyper_seed=9, covariates=X, want_eps=False)
    gm = ref["gamma"][20:].mean(0)
    np.testing.assert_allclose(gm, np.array([0.5, -0.3, 0.0]) / sd, atol=0.06)


def test_bayesw_oracle_fixed_effects():
    """BayesW oracle with covariates (src/BayesW.cpp:1366-1413): the fixed effects leave 0 towards the simulated values, the
    density restated in C equals the formula of :119-129, and a chain without covariates is unchanged by F = 0."""
    if oracle.arms_ref() is None:
        pytest.skip("oracle/_ref/libarms_ref.so not built")
    rng = np.random.default_rng(3)
    N, M, K, G, T, SR, n_iter, quad, seed = 400, 24, 4, 1, 1, 1, 12, 25, 9
    bed, g = random_bed(rng, M, N, pmiss=0.0)
    sp = reference_lists(bed, N)
    X = np.column_stack([rng.integers(0, 2, N).astype(np.float64), rng.normal(size=N)])
    true = np.array([0.05, -0.03])
    w = np.log(rng.exponential(size=N))
    y = 4.1 + X @ true + w / 10.0 + 0.577215664901532 / 10.0
    fail = np.ones(N)
    tm = oracle.TapeMaker(seed, T, M).make(n_iter)
    xI = np.tile(np.arange(2, dtype=np.int32), (n_iter, 1))
    mS = np.array([[0.0, 0.001, 0.01, 0.1]])
    ref = oracle.bw_chain(N, M, T, K, G, SR, n_iter, quad, sp, y, fail, np.zeros(M, np.int32), mS, dict(perm=tm["perm"], p=tm["u"], xI=xI), seed,
                          hyper_seed=1, covariates=X)
    gam = ref["gamma"][n_iter // 2:].mean(axis=0)
    assert np.all(np.abs(gam - true) < 0.02), gam          # +-0.075 per step: a dozen iterations reach 0.05 / -0.03
    # density (:119-129) against numpy
    eps = rng.normal(size=N) * 0.1
    x, alpha, sff, sig = 0.013, 9.5, float(X[:, 0] @ fail), 100.0
    L = oracle.lib()
    L.ho_bw_gamma_dens.restype = C.c_double
    got = L.ho_bw_gamma_dens(C.c_double(x), eps.ctypes.data_as(C.c_void_p), np.ascontiguousarray(X[:, 0]).ctypes.data_as(C.c_void_p), C.c_int(N),
                             C.c_double(alpha), C.c_double(sff), C.c_double(sig))
    want = -alpha * x * sff - np.exp((eps - X[:, 0] * x) * alpha - 0.577215664901532).sum() - x * x / (2 * sig)
    np.testing.assert_allclose(got, want, rtol=1e-12)
