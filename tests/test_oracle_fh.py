"""bayesFHMPI in the oracle (src/BayesRRm.cpp:1125-1163, 1727-1731, 1747-1748, 1869-1872, 1942-1952, 2503-2510, 2557-2565) and
the two prior files (--groupPriorsFile :2545-2548, --dPriorsFile :2551-2554), against an independent dense numpy restatement."""
import numpy as np

import oracle
from helpers import random_bed, simulate_y


def _dense_fh_chain(N, M, K, G, n_iter, g, y, groups, tape, sigmaG0, fh, state0, hyper_seed, group_priors=None, dirichlet=None, fh_seed=0):
    """One task, synchronisation after every marker: eps is always y - X beta - mu. Plain numpy on the dense matrix."""
    n1, n2, nm = (g == 1).sum(1), (g == 2).sum(1), (g < 0).sum(1)
    mave, mstd = oracle.marker_stats_brr(N, n1.astype(np.uint64), n2.astype(np.uint64), nm.astype(np.uint64))
    X = np.where(g < 0, mave[:, None], g) - mave[:, None]
    X = X * mstd[:, None]
    ys = oracle.center_and_scale(y)
    eps = ys.copy()
    sigmaE = float(eps @ eps) / N * 0.5
    sigmaG = np.array(sigmaG0, float)
    pi = np.full((G, K), 0.5 / (K - 1))   # overwritten below
    mt = oracle.MT(hyper_seed)
    v0L, v0t, v0c, s02c, tau0 = fh["v0L"], fh["v0t"], fh["v0c"], fh["s02c"], fh["tau0"]
    hypTau, tau, c = state0[0], state0[1], np.array(state0[2:], float)
    lam = np.full(M, c.sum() / M)
    nu = np.zeros(M)
    beta = np.zeros(M)
    comp = np.zeros(M, np.int32)
    mu = 0.0
    mS = fh["mS"]
    for gg in range(G):
        pi[gg, 0] = 0.5
        pi[gg, 1:] = 0.5 * mS[gg, 1:] / mS[gg, 1:].sum()
    active = sigmaG != 0.0
    Mg = np.bincount(groups, minlength=G)
    out = dict(beta=[], comp=[], lam=[], nu=[], fh=[], sigmaG=[], sigmaE=[], pi=[])
    sh = 0.5 + 0.5 * v0L
    for it in range(n_iter):
        eps = eps + mu
        mu = eps.sum() / N + np.sqrt(sigmaE / N) * tape["zmu"][it, 0]
        eps = eps - mu
        cass = np.zeros((G, K), np.int64)
        for j in range(M):
            m = int(tape["perm"][it, j]); gr = int(groups[m])
            nu[m] = 1.0 / (oracle.fh_gamma(fh_seed, m, it, 0, sh) * (1.0 / (v0L / lam[m] + 1)))
            lt = tau * c[gr] / (tau + c[gr] * lam[m])
            b_old = beta[m]
            if active[gr]:
                den = (N - 1) + sigmaE / lt
                num = float(X[m] @ eps) + b_old * (N - 1)
                muk = num / den
                logL = np.log(pi[gr]).copy()
                logL[1:] += -0.5 * np.log((lt / sigmaE) * (N - 1) + 1.0) + muk * num / (2 * sigmaE)
                p = np.exp(logL - logL.max()); p /= p.sum()
                k = min(int(np.searchsorted(np.cumsum(p), tape["u"][it, j], side="left")), K - 1)
                beta[m] = 0.0 if k == 0 else muk + np.sqrt(sigmaE / den) * tape["z"][it, j]
                comp[m] = k
                cass[gr, k] += 1
            else:
                beta[m] = 0.0
            lam[m] = 1.0 / (oracle.fh_gamma(fh_seed, m, it, 1, sh) * (1.0 / (0.5 * beta[m] ** 2 / tau + v0L / nu[m])))
            eps = eps + (b_old - beta[m]) * X[m]
        bsq = np.array([(beta[groups == gg] ** 2).sum() for gg in range(G)])
        sb = float((beta ** 2 / lam).sum())
        for gg in range(G):
            if Mg[gg] == 0:
                continue
            m0 = Mg[gg] - cass[gg, 0]
            if m0 == 0 or cass[gg].sum() == 0:
                active[gg] = False; sigmaG[gg] = 0.0
                continue
            hypTau = 1.0 / (mt.gamma(0.5 + 0.5 * v0t) * (1.0 / (1.0 / tau0 ** 2 + 1.0 / tau)))
            tau = 1.0 / (mt.gamma(0.5 * (m0 + v0t)) * (1.0 / (v0t / hypTau + 0.5 * sb)))
            dof, sc = v0c + m0, (bsq[gg] * m0 + v0c * s02c) / (v0c + m0)
            c[gg] = 1.0 / (mt.gamma(0.5 * dof) * (1.0 / (0.5 * dof * sc)))
            sigmaG[gg] = bsq[gg]
            dk = np.ones(K) if dirichlet is None else dirichlet[gg]
            w = np.array([mt.gamma(cass[gg, k] + dk[k]) for k in range(K)])
            pi[gg] = w / w.sum()
        esq = float(eps @ eps)
        dof, sc = 0.0001 + N, (esq + 0.0001 * 0.0001) / (0.0001 + N)
        sigmaE = 1.0 / (mt.gamma(0.5 * dof) * (1.0 / (0.5 * dof * sc)))
        out["beta"].append(beta.copy()); out["comp"].append(comp.copy()); out["lam"].append(lam.copy()); out["nu"].append(nu.copy())
        out["fh"].append(np.concatenate([[hypTau, tau, sb], c])); out["sigmaG"].append(sigmaG.copy()); out["sigmaE"].append(sigmaE)
        out["pi"].append(pi.copy())
    return {k: np.array(v) for k, v in out.items()}


def test_fh_gamma_spec_moments_and_shapes_below_one():
    for a in (2.0, 0.75):
        x = np.array([oracle.fh_gamma(5, m, 3, 0, a) for m in range(20000)])
        assert abs(x.mean() - a) < 0.03 and abs(x.var() - a) < 0.08, (a, x.mean(), x.var())
    # the two draws of a marker, and two iterations, are different streams
    assert oracle.fh_gamma(5, 1, 3, 0, 2.0) != oracle.fh_gamma(5, 1, 3, 1, 2.0) != oracle.fh_gamma(5, 1, 4, 1, 2.0)


def test_fh_chain_equals_dense_numpy_restatement():
    rng = np.random.default_rng(11)
    N, M, K, G, n_iter = 300, 40, 3, 2, 4
    bed, g = random_bed(rng, M, N)
    sp = oracle.sparse_fill_indices(bed, N)
    y = simulate_y(rng, g, n_causal=6)
    groups = (np.arange(M) % G).astype(np.int32)
    mS = np.array([[0.0, 0.01, 0.1], [0.0, 0.001, 0.05]])
    tape = oracle.TapeMaker(2, 1, M).make(n_iter)
    fh = dict(oracle.FH_DEFAULTS, seed=77)
    state0 = np.array([1.3, 0.02, 0.4, 0.7])
    dirichlet = np.array([[2.0, 1.0, 1.0], [1.0, 3.0, 0.5]])
    ref = oracle.brr_chain(N, M, 1, K, G, 1, n_iter, sp, y, groups, mS, tape, np.array([0.3, 0.2]), hyper_seed=5,
                           fh=dict(fh, state0=state0), dirichlet=dirichlet)
    want = _dense_fh_chain(N, M, K, G, n_iter, g, y, groups, tape, [0.3, 0.2], dict(fh, mS=mS), state0, 5, dirichlet=dirichlet, fh_seed=77)
    assert np.array_equal(ref["comp"], want["comp"])
    np.testing.assert_allclose(ref["beta"], want["beta"], rtol=1e-8, atol=1e-13)
    np.testing.assert_allclose(ref["lambda"], want["lam"], rtol=1e-8)
    np.testing.assert_allclose(ref["nu"], want["nu"], rtol=1e-8)
    np.testing.assert_allclose(ref["fh"], want["fh"], rtol=1e-8)
    np.testing.assert_allclose(ref["sigmaG"], want["sigmaG"], rtol=1e-8)   # sigmaG = beta_squaredNorm (:2565)
    np.testing.assert_allclose(ref["sigmaE"], want["sigmaE"], rtol=1e-8)
    np.testing.assert_allclose(ref["pi"], want["pi"], rtol=1e-8)
    assert (ref["comp"] > 0).any() and (ref["lambda"] != ref["lambda"][0, 0]).any()


def test_fh_initial_draws_and_taped_hyper():
    """state0 = NULL draws hypTau, tau, c_slab in the reference's order (:1147-1154); a taped fh_hyper is taken as is."""
    rng = np.random.default_rng(3)
    N, M, K, G, n_iter = 200, 24, 3, 1, 2
    bed, g = random_bed(rng, M, N)
    sp = oracle.sparse_fill_indices(bed, N)
    y = simulate_y(rng, g, n_causal=4)
    mS = np.array([[0.0, 0.01, 0.1]])
    tape = oracle.TapeMaker(2, 2, M).make(n_iter)
    fh = dict(oracle.FH_DEFAULTS, v0t=5.0, tau0=0.5, seed=1)
    mt = oracle.MT(21)
    hypTau = 1.0 / (mt.gamma(0.5) * (1.0 / (1.0 / 0.25)))
    tau = 1.0 / (mt.gamma(0.5 * 5.0) * (1.0 / (5.0 / hypTau)))
    c0 = 1.0 / (mt.gamma(0.5 * 3.0) * (1.0 / (0.5 * 3.0 * 1.0)))
    a = oracle.brr_chain(N, M, 2, K, G, 3, n_iter, sp, y, np.zeros(M, np.int32), mS, tape, np.array([0.3]), hyper_seed=21, fh=fh)
    b = oracle.brr_chain(N, M, 2, K, G, 3, n_iter, sp, y, np.zeros(M, np.int32), mS, tape, np.array([0.3]), hyper_seed=21,
                         fh=dict(fh, state0=np.array([hypTau, tau, c0])))
    np.testing.assert_allclose(a["lambda"][0], b["lambda"][0], rtol=1e-12)   # same start => same first iteration
    np.testing.assert_allclose(a["beta"][0], b["beta"][0], rtol=1e-12, atol=0)
    tp = dict(tape, fh_hyper=np.array([[[2.0, 0.03, 0.5]], [[2.5, 0.04, 0.6]]]))
    hyp = dict(sigmaG=np.zeros((n_iter, G)), sigmaE=np.array([0.6, 0.55]), pi=np.tile(np.array([0.8, 0.15, 0.05]), (n_iter, G, 1)))
    t = oracle.brr_chain(N, M, 2, K, G, 3, n_iter, sp, y, np.zeros(M, np.int32), mS, tp, np.array([0.3]), hyper=hyp, fh=fh)
    np.testing.assert_allclose(t["fh"][:, [0, 1, 3]], tp["fh_hyper"][:, 0, :])
    np.testing.assert_allclose(t["sigmaG"], t["bsq"])


def test_group_priors_file_values_enter_the_sigmaG_draw():
    rng = np.random.default_rng(8)
    N, M, K, G, n_iter = 200, 30, 3, 2, 2
    bed, g = random_bed(rng, M, N)
    sp = oracle.sparse_fill_indices(bed, N)
    y = simulate_y(rng, g, n_causal=4)
    groups = (np.arange(M) >= M // 2).astype(np.int32)
    mS = np.array([[0.0, 0.01, 0.1], [0.0, 0.01, 0.1]])
    tape = oracle.TapeMaker(2, 1, M).make(n_iter)
    base = oracle.brr_chain(N, M, 1, K, G, 1, n_iter, sp, y, groups, mS, tape, np.array([0.3, 0.3]), hyper_seed=4)
    same = oracle.brr_chain(N, M, 1, K, G, 1, n_iter, sp, y, groups, mS, tape, np.array([0.3, 0.3]), hyper_seed=4,
                            group_priors=np.full((G, 2), 0.0001), dirichlet=np.ones((G, K)))
    np.testing.assert_array_equal(base["sigmaG"], same["sigmaG"])
    pri = np.array([[4.0, 0.2], [0.0001, 0.0001]])
    other = oracle.brr_chain(N, M, 1, K, G, 1, n_iter, sp, y, groups, mS, tape, np.array([0.3, 0.3]), hyper_seed=4, group_priors=pri)
    # first iteration: same gamma variate shape? no - dof differs; check the formula of group 0 through the MT stream
    mt = oracle.MT(4)
    m0 = int((groups == 0).sum() - other["cass"][0, 0, 0])
    dof, sc = 4.0 + m0, (other["bsq"][0, 0] * m0 + 4.0 * 0.2) / (4.0 + m0)
    want = 1.0 / (mt.gamma(0.5 * dof) * (1.0 / (0.5 * dof * sc)))
    np.testing.assert_allclose(other["sigmaG"][0, 0], want, rtol=1e-12)
