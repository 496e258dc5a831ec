"""GPU parity tests of the BayesW path (through the C ABI) against the CPU oracle, whose ARMS is the reference's own
object code (oracle/_ref/libarms_ref.so)."""
import numpy as np
import pytest

import oracle
from helpers import random_bed, reference_lists

pytestmark = pytest.mark.gpu
RTOL = 1e-10
EUM = 0.577215664901532


def _weibull_data(rng, g, mu=4.1, alpha=10.0, h2=0.3, cens=0.1):
    M, N = g.shape
    x = np.where(g < 0, 0, g).astype(np.float64)
    x = (x - x.mean(1, keepdims=True)) / (x.std(1, keepdims=True) + 1e-12)
    causal = rng.choice(M, size=max(3, M // 10), replace=False)
    b = rng.normal(0, np.sqrt(h2 * (np.pi ** 2 / 6) / alpha ** 2 / len(causal)), size=len(causal))
    w = np.log(rng.exponential(size=N))           # Gumbel(min) noise of a Weibull log-time
    y = mu + x[causal].T @ b + w / alpha + EUM / alpha
    fail = (rng.random(N) > cens).astype(np.float64)
    return y, fail


def _store(N, M, **kw):
    import hydra_b200
    return hydra_b200.GenotypeStore(N, M, model="bayesW", **kw)


@pytest.mark.parametrize("repr_mode,n_slices", [("sparse", 0), ("bed", 3)])
def test_bayesw_unit_kernels(repr_mode, n_slices):
    import hydra_b200
    from hydra_b200 import sampler
    rng = np.random.default_rng(11)
    N, M = 1203, 40
    bed, g = random_bed(rng, M, N, pmiss=0.02)
    sp = reference_lists(bed, N)
    y, fail = _weibull_data(rng, g)
    with _store(N, M, n_groups=1, n_mix=4, repr_mode=repr_mode, n_slices=n_slices) as st:
        st.load_data_from_bed(bed)
        st.finalize()
        bw = hydra_b200.BayesW(st, y, fail, [[0.001, 0.01, 0.1]], quad_points=25, seed=3)
        sd, sf = bw.marker_stats()
        mave, omstd, osf = oracle.bw_marker_stats(N, sp, fail)
        assert np.array_equal(st.marker_stats()[0], mave)
        np.testing.assert_allclose(sd, omstd, rtol=1e-15)
        np.testing.assert_allclose(sf, osf, rtol=1e-13, atol=1e-12)
        # sums of vi by genotype class with the marker's own effect removed (src/BayesW.cpp:1499-1525), against the dense formula
        eps = rng.normal(size=N) * 0.1
        st.set_epsilon(eps)
        alpha = 9.3
        markers = np.arange(M, dtype=np.uint32)
        bold = np.where(rng.random(M) < 0.5, 0.0, rng.normal(size=M) * 0.02)
        got = sampler.bw_vi_sums(st, markers, bold, alpha)
        for m in range(M):
            delta = np.where(g[m] < 0, 0.0, (g[m] - mave[m]) / sd[m]) * bold[m]
            delta[g[m] < 0] = 0.0
            v = np.exp(alpha * (eps + delta) - EUM)
            want = np.array([v.sum(), v[g[m] == 1].sum(), v[g[m] == 2].sum()])
            np.testing.assert_allclose(got[m, :3], want, rtol=1e-11)
            np.testing.assert_allclose(got[m, 3], want[0] - want[1] - want[2], rtol=1e-9)
        # N-sum of exp inside the mu / alpha log-densities
        a, b = 7.7, -0.3
        np.testing.assert_allclose(sampler.bw_sum_exp(st, a, b), np.exp(a * eps + b).sum(), rtol=1e-12)


@pytest.mark.parametrize("quad", [3, 7, 11, 25])
def test_bayesw_marginal_likelihoods_match_oracle(quad):
    from hydra_b200 import sampler
    rng = np.random.default_rng(quad)
    with _store(64, 4, n_groups=1, n_mix=4) as st:
        for _ in range(20):
            p = 0.01 + 0.49 * rng.random()
            mean, sd = 2 * p, np.sqrt(2 * p * (1 - p))
            Nn = 5000
            v1, v2, v0 = Nn * 2 * p * (1 - p) * rng.uniform(0.8, 1.2), Nn * p * p * rng.uniform(0.8, 1.2), Nn * (1 - p) ** 2 * rng.uniform(0.8, 1.2)
            pars = [rng.uniform(5, 12), rng.uniform(0.005, 0.03), rng.normal() * 30, v0 + v1 + v2, v0, v1, v2, mean, sd, mean / sd]
            prior, cVa = [0.9, 0.05, 0.03, 0.02], [0.001, 0.01, 0.1]
            got = sampler.bw_marginal_likelihoods(st, quad, pars, prior, cVa)
            want = oracle.bw_marginal_likelihoods(quad, pars, prior, cVa)
            np.testing.assert_allclose(got, want, rtol=1e-11)   # measured max 9.1e-13 over 200 cases (scripts/bw_error_probe.py)


def test_bayesw_device_arms_matches_reference_arms():
    if oracle.arms_ref() is None:
        pytest.skip("oracle/_ref/libarms_ref.so not built")
    from hydra_b200 import sampler
    rng = np.random.default_rng(5)
    same = 0
    with _store(64, 4, n_groups=1, n_mix=4) as st:
        for t in range(60):
            p = 0.01 + 0.49 * rng.random()
            mean, sd = 2 * p, np.sqrt(2 * p * (1 - p))
            Nn = 5000
            v1, v2, v0 = Nn * 2 * p * (1 - p), Nn * p * p, Nn * (1 - p) ** 2
            pars = [rng.uniform(5, 12), rng.uniform(0.005, 0.03), rng.normal() * 30, v0 + v1 + v2, v0, v1, v2, mean, sd, mean / sd]
            Ck, ssg, bold = float(rng.choice([0.001, 0.01, 0.1])), 0.02, float(rng.choice([0.0, 0.01, -0.02]))
            got = sampler.bw_arms_beta(st, pars, Ck, ssg, bold, 77, 3, 5, t)
            want = oracle.bw_sample_beta(pars, Ck, ssg, bold, 77, 3, 5, t)
            assert got["err"] == want["err"] == 0
            # same uniforms, same control flow: identical counts; the sample agrees to the precision of exp/log
            if got["nrand"] == want["nrand"] and got["neval"] == want["neval"]:
                same += 1
                np.testing.assert_allclose(got["beta"], want["beta"], rtol=1e-9, atol=1e-15)
    # same uniforms, same control flow in ALL cases (400/400 in scripts/bw_error_probe.py): the accept/reject decisions of the
    # device ARMS are those of the reference's object code. The sample itself is the inverse of the piecewise-exponential
    # envelope's cumulative, x = xl + log(1 + slope * area_needed * exp(-yl)) / slope (src/BayesW_arms.cpp:390-415): a 1-ulp
    # difference between CUDA's and glibc's exp/log is amplified by 1 / (slope * dx) ~ 1e5-1e7 there, hence 1e-9 (measured
    # max 2.2e-9 over 400 cases) on the SAMPLE although every log-density agrees to 1e-13.
    assert same == 60


@pytest.mark.parametrize("T,SR,G,repr_mode,replay_hyper", [(1, 1, 1, "sparse", True), (3, 4, 2, "sparse", True), (2, 3, 1, "bed", False)])
def test_bayesw_chain_replay(T, SR, G, repr_mode, replay_hyper):
    if oracle.arms_ref() is None:
        pytest.skip("oracle/_ref/libarms_ref.so not built")
    import hydra_b200
    N, M, K, n_iter, seed, quad = 900, 90, 4, 3, 17, 25
    rng = np.random.default_rng(T * 10 + SR)
    bed, g = random_bed(rng, M, N, pmiss=0.01)
    sp = reference_lists(bed, N)
    y, fail = _weibull_data(rng, g)
    groups = (np.arange(M) % G).astype(np.int32)
    mS = np.tile(np.array([0.0, 0.001, 0.01, 0.1]), (G, 1))
    tm = oracle.TapeMaker(seed, T, M).make(n_iter)
    tape = dict(perm=tm["perm"], p=tm["u"])
    hseed = (seed ^ 0x5bd1e995) & 0xFFFFFFFF
    ref = oracle.bw_chain(N, M, T, K, G, SR, n_iter, quad, sp, y, fail, groups, mS, tape, seed, hyper_seed=hseed)
    with _store(N, M, tasks=T, sync_rate=SR, n_groups=G, n_mix=K, repr_mode=repr_mode) as st:
        st.load_data_from_bed(bed)
        st.finalize()
        bw = hydra_b200.BayesW(st, y, fail, mS, groups=groups, quad_points=quad, seed=seed)
        for it in range(n_iter):
            tp = dict(perm=tape["perm"][it], p=tape["p"][it])
            if replay_hyper:
                tp.update(sigmaG=ref["sigmaG"][it], pi=ref["pi"][it])
            o = bw.iteration(tp)
            beta, comp = bw.state()
            h = bw.hyper()
            # tolerances = north_star's 1e-10 where the measured difference allows it (scripts/bw_error_probe.py: mu 1.7e-14,
            # alpha 5.7e-13, beta 2.8e-11, epsilon 2.9e-13 absolute, sigmaG 1.8e-13, pi 0 on these chains)
            np.testing.assert_allclose(o["mu"], ref["mu"][it], rtol=1e-11, err_msg=f"mu it {it}")
            np.testing.assert_allclose(o["alpha"], ref["alpha"][it], rtol=1e-11, err_msg=f"alpha it {it}")
            assert np.array_equal(comp, ref["comp"][it]), f"components differ at iteration {it}"
            np.testing.assert_allclose(beta, ref["beta"][it], rtol=1e-10, atol=1e-14, err_msg=f"beta it {it}")
            assert np.array_equal(h["cass"], ref["cass"][it])
            assert o["n_sync"] == ref["nsync"][it]
            np.testing.assert_allclose(bw.epsilon(), ref["eps"][it], rtol=1e-10, atol=1e-12, err_msg=f"eps it {it}")
            np.testing.assert_allclose(h["sigmaG"], ref["sigmaG"][it], rtol=1e-10)
            np.testing.assert_allclose(h["pi"], ref["pi"][it], rtol=1e-10)


def test_bayesw_chain_replay_with_covariates():
    """Fixed effects of BayesW (src/BayesW.cpp:1366-1413, gamma_dens :119-129): each gamma by ARMS between mu and alpha, the
    density's N-sum on the device; order of the covariates from the tape, ARMS uniforms from the Philox stream 'ARMG'."""
    if oracle.arms_ref() is None:
        pytest.skip("oracle/_ref/libarms_ref.so not built")
    import hydra_b200
    N, M, K, G, T, SR, F, n_iter, seed, quad = 900, 90, 4, 1, 2, 3, 3, 4, 23, 25
    rng = np.random.default_rng(5)
    bed, g = random_bed(rng, M, N, pmiss=0.01)
    sp = reference_lists(bed, N)
    X = np.column_stack([rng.integers(0, 2, N).astype(np.float64), rng.normal(size=N), rng.normal(size=N) * 0.5])
    y, fail = _weibull_data(rng, g)
    y = y + X @ np.array([0.03, -0.02, 0.01])
    groups = np.zeros(M, np.int32)
    mS = np.array([[0.0, 0.001, 0.01, 0.1]])
    tm = oracle.TapeMaker(seed, T, M).make(n_iter)
    xI = np.stack([rng.permutation(F) for _ in range(n_iter)]).astype(np.int32)
    tape = dict(perm=tm["perm"], p=tm["u"], xI=xI)
    ref = oracle.bw_chain(N, M, T, K, G, SR, n_iter, quad, sp, y, fail, groups, mS, tape, seed, hyper_seed=(seed ^ 0x5bd1e995) & 0xFFFFFFFF,
                          covariates=X)
    assert np.abs(ref["gamma"][-1]).max() > 1e-3          # the fixed effects moved
    with _store(N, M, tasks=T, sync_rate=SR, n_groups=G, n_mix=K, repr_mode="sparse") as st:
        st.load_data_from_bed(bed)
        st.finalize()
        bw = hydra_b200.BayesW(st, y, fail, mS, groups=groups, quad_points=quad, seed=seed, covariates=X)
        for it in range(n_iter):
            o = bw.iteration(dict(perm=tape["perm"][it], p=tape["p"][it], xI=xI[it], sigmaG=ref["sigmaG"][it], pi=ref["pi"][it]))
            gam, order = bw.gamma()
            assert np.array_equal(order, xI[it])
            # a gamma is a point of [gamma_old - 0.075, gamma_old + 0.075] found by inverting the ARMS envelope: its error scales with
            # that range (measured 3.4e-12 absolute), not with its own size (values of 1e-3 here)
            np.testing.assert_allclose(gam, ref["gamma"][it], rtol=1e-10, atol=2e-11, err_msg=f"gamma it {it}")
            np.testing.assert_allclose(o["mu"], ref["mu"][it], rtol=1e-11, err_msg=f"mu it {it}")
            np.testing.assert_allclose(o["alpha"], ref["alpha"][it], rtol=1e-11, err_msg=f"alpha it {it}")
            beta, comp = bw.state()
            assert np.array_equal(comp, ref["comp"][it]), f"components differ at iteration {it}"
            # gamma's absolute error enters epsilon as x * d(gamma) (|x| up to 3) and from there the ARMS draws of the effects
            np.testing.assert_allclose(beta, ref["beta"][it], rtol=1e-10, atol=1e-11, err_msg=f"beta it {it}")
            assert o["n_sync"] == ref["nsync"][it]
            np.testing.assert_allclose(bw.epsilon(), ref["eps"][it], rtol=1e-10, atol=5e-11, err_msg=f"eps it {it}")
