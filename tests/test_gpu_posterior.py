"""Normal-mode (own RNG) statistical checks at the example scale of BASELINE configs 1-3 (N=5000, M=10000 stand-in data:
the reference's .bed files are missing from its checkout): posterior means recover the simulated truth within Monte-Carlo
error, for the layouts the configs name (1 task sync 1 on BED input; grouped mixtures with 4 tasks sync 10; BayesW)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _simulate(N, M, n_causal, h2, seed):
    import hydra_b200
    from hydra_b200 import synth
    st = hydra_b200.GenotypeStore(N, M, repr_mode="bed")
    p = synth.maf_spectrum(M, 0.05, 0.5, seed=seed)
    st.load_synthetic(seed, synth.thresholds(p, 0.001))
    st.finalize()
    rng = np.random.default_rng(seed)
    causal = np.sort(rng.choice(M, n_causal, replace=False)).astype(np.uint32)
    b = rng.normal(0, np.sqrt(h2 / n_causal), n_causal)
    st.set_epsilon(np.zeros(N))
    st.sparse_scaadd(causal, b)
    g = st.get_epsilon()
    bed = np.stack([st.export_bed(m) for m in range(M)])
    st.close()
    return bed, g, causal, b


@pytest.mark.parametrize("tasks,sync_rate,groups,repr_mode", [(1, 1, 1, "bed"), (4, 10, 2, "sparse")])
def test_bayesrr_recovers_heritability_and_effects(tasks, sync_rate, groups, repr_mode):
    import hydra_b200
    N, M, n_causal, h2 = 5000, 10000, 200, 0.5
    bed, g, causal, b = _simulate(N, M, n_causal, h2, seed=7)
    rng = np.random.default_rng(1)
    scale = np.sqrt(h2 / g.var())
    g = g * scale
    y = g + rng.normal(0, np.sqrt(1 - h2), N)
    h2_real = g.var() / y.var()                                  # 0.508 for this draw of the noise
    grp = (np.arange(M) % groups).astype(np.int32)
    with hydra_b200.GenotypeStore(N, M, tasks=tasks, sync_rate=sync_rate, n_groups=groups, n_mix=4, repr_mode=repr_mode) as st:
        st.load_data_from_bed(bed)
        st.finalize()
        brr = hydra_b200.BayesRRm(st, y, [[0.001, 0.01, 0.1]] * groups, groups=grp, seed=1222)
        n_it, burn = 700, 300
        h2s, bsum = [], np.zeros(M)
        for it in range(n_it):
            brr.iteration()
            if it >= burn:
                h = brr.hyper()
                h2s.append(h["sigmaG"].sum() / (h["sigmaG"].sum() + h["sigmaE"]))
                bsum += brr.state()[0]
        h2_hat = float(np.mean(h2s))
        bmean = bsum / (n_it - burn)
        # measured on the B200 (both layouts, 400 / 700 / 1000 iterations): h2 0.512-0.522 with a posterior sd of 0.023, r 0.958-0.960,
        # slope 0.878-0.888, null / causal 0.010-0.011 (cf. example/normal.h2: 0.5136 on the reference's data)
        assert abs(h2_hat - h2_real) < 0.04, (h2_hat, h2_real)
        r = np.corrcoef(bmean[causal], b)[0, 1]
        assert r > 0.93, r                                       # posterior means track the simulated effects
        bt = b * scale
        slope = float(np.dot(bmean[causal], bt) / np.dot(bt, bt))
        assert 0.85 < slope < 1.05, slope                        # regression of posterior mean on truth: mild shrinkage, no bias
        assert np.abs(bmean[np.setdiff1d(np.arange(M), causal)]).mean() < 0.03 * np.abs(bmean[causal]).mean()


def test_bayesw_recovers_weibull_parameters():
    import hydra_b200
    N, M, n_causal = 5000, 1500, 60
    bed, g, causal, b = _simulate(N, M, n_causal, 1.0, seed=11)
    rng = np.random.default_rng(2)
    mu, alpha, var_g = 4.1, 10.0, 0.01675                         # example/Weibull.h2: var_g 0.01675, mu 4.1, alpha 10
    g = g * np.sqrt(var_g / g.var())
    w = np.log(rng.exponential(size=N))
    y = mu + g + w / alpha + 0.577215664901532 / alpha
    fail = np.ones(N)
    with hydra_b200.GenotypeStore(N, M, tasks=4, sync_rate=5, n_groups=1, n_mix=4, repr_mode="sparse", model="bayesW") as st:
        st.load_data_from_bed(bed)
        st.finalize()
        bw = hydra_b200.BayesW(st, y, fail, [[0.001, 0.01, 0.1]], quad_points=25, seed=5)
        mus, alphas, sg = [], [], []
        for it in range(120):
            o = bw.iteration()
            if it >= 50:
                mus.append(o["mu"]); alphas.append(o["alpha"]); sg.append(bw.hyper()["sigmaG"].sum())
        assert abs(np.mean(mus) - mu) < 0.02, np.mean(mus)         # E[exp(alpha*eps - EuMasc)] = 1 puts the intercept at mu
        assert 8.0 < np.mean(alphas) < 12.5, np.mean(alphas)
        assert 0.2 * var_g < np.mean(sg) < 5 * var_g, np.mean(sg)


def test_bayesfh_recovers_effects_and_learns_local_scales():
    """bayesFHMPI in normal mode (device RNG): the horseshoe-type prior finds the causal markers, its local scales lambda_var grow on
    them, sigmaG (= sum beta^2, src/BayesRRm.cpp:2565) gives the heritability. Calibrated on the CPU oracle run with the same
    data and streams (h2 0.497 against a realised 0.498, r 0.963, slope 0.944, null / causal 0.0012, lambda 0.73 against 0.15)."""
    import hydra_b200
    from helpers import random_bed
    N, M, n_causal, h2, seed, n_it, burn = 2000, 1000, 50, 0.5, 1222, 300, 100
    rng = np.random.default_rng(5)
    bed, g = random_bed(rng, M, N, maf_lo=0.05, pmiss=0.0)
    x = g.astype(np.float64)
    x = (x - x.mean(1, keepdims=True)) / x.std(1, keepdims=True)
    causal = np.sort(rng.choice(M, n_causal, replace=False))
    b = rng.normal(0, np.sqrt(h2 / n_causal), n_causal)
    gv = x[causal].T @ b
    y = gv + rng.normal(0, np.sqrt(gv.var() * (1 - h2) / h2), N)
    with hydra_b200.GenotypeStore(N, M, tasks=4, sync_rate=5, n_groups=1, n_mix=3, repr_mode="sparse") as st:
        st.load_data_from_bed(bed)
        st.finalize()
        brr = hydra_b200.BayesRRm(st, y, [[0.01, 0.1]], sigmaG0=[0.5], seed=seed, fh={})
        bsum, lsum, h2s = np.zeros(M), np.zeros(M), []
        for it in range(n_it):
            brr.iteration()
            if it >= burn:
                h = brr.hyper()
                h2s.append(h["sigmaG"].sum() / (h["sigmaG"].sum() + h["sigmaE"]))
                bsum += brr.state()[0]
                lsum += brr.fh_state()["lambda_var"]
    bm, lm = bsum / (n_it - burn), lsum / (n_it - burn)
    bt = b / y.std(ddof=1)
    null = np.setdiff1d(np.arange(M), causal)
    assert abs(np.mean(h2s) - gv.var() / y.var()) < 0.04, np.mean(h2s)
    assert np.corrcoef(bm[causal], bt)[0, 1] > 0.93
    assert 0.85 < bm[causal] @ bt / (bt @ bt) < 1.05
    assert np.abs(bm[null]).mean() < 0.02 * np.abs(bm[causal]).mean()
    assert lm[causal].mean() > 2.0 * lm[null].mean(), (lm[causal].mean(), lm[null].mean())
