"""GPU parity of bayesFHMPI and of the two prior files against the oracle (through the C ABI).

src/BayesRRm.cpp:1125-1163 (initialisation), :1727-1731 (nu_var, the marker's own prior variance), :1747-1748 / :1869-1872 (its use
in the mixture draw), :1942-1952 (lambda_var), :2503-2510 (scaled sum of squares), :2557-2565 (hypTau, tau, c_slab),
:2545-2554 (--groupPriorsFile / --dPriorsFile). Tolerance: north_star's 1e-10 relative on effects, residuals and the FH scales;
components bit-exact."""
import numpy as np
import pytest

import oracle
from helpers import bed_from_lists, random_bed, reference_lists, simulate_y

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _case(N, M, T, SR, G, K, repr_mode, n_iter, seed, replay, n_slices=0, n_causal=None, fhp=None, dirichlet=None, group_priors=None,
          restart=False, n_cov=0):
    import hydra_b200
    rng = np.random.default_rng(seed)
    bed, g = random_bed(rng, M, N, pmiss=0.01)
    sp = reference_lists(bed, N)
    y = simulate_y(rng, g, n_causal=max(3, M // 10) if n_causal is None else n_causal)
    X = None
    if n_cov:   # fixed effects next to the FH prior: the covariate block runs between the FH hyper-parameters and sigmaE (:2648-2681)
        X = rng.normal(size=(N, n_cov))
        X = (X - X.mean(0)) / X.std(0, ddof=1)
        y = y + X @ rng.normal(0, 0.4, n_cov)
    groups = (np.arange(M) % G).astype(np.int32)
    mS = np.tile(np.array([0.0] + [10.0 ** (-(K - 1 - k)) for k in range(1, K)]), (G, 1))
    sigmaG0 = rng.uniform(0.2, 0.8, size=G)
    tape = oracle.TapeMaker(seed, T, M).make(n_iter)
    if n_cov and replay:
        tape["xI"] = np.array([rng.permutation(n_cov) for _ in range(n_iter)], np.int32)
        tape["zcov"] = rng.normal(size=(n_iter, n_cov))
    fh = dict(oracle.FH_DEFAULTS, **(fhp or {}))
    fnz = (sp.N1L + sp.N2L + sp.NML).astype(np.float64) / N
    usebed = {"sparse": np.zeros(M, np.uint8), "bed": np.ones(M, np.uint8), "mixed": (fnz > 0.35).astype(np.uint8)}[repr_mode]
    hseed = (seed ^ 0x5bd1e995) & 0xFFFFFFFF
    state0 = None
    if replay:   # every FH variate comes from the tape: standard gammas per marker, hyper values per group
        sh = 0.5 + 0.5 * fh["v0L"]
        tape["gnu"] = rng.gamma(sh, size=(n_iter, M))
        tape["glam"] = rng.gamma(sh, size=(n_iter, M))
        tape["fh_hyper"] = np.stack([rng.uniform(0.5, 3.0, (n_iter, G)), rng.uniform(0.005, 0.05, (n_iter, G)), rng.uniform(0.1, 1.0, (n_iter, G))], axis=2)
        state0 = np.concatenate([[1.7, 0.02], rng.uniform(0.2, 0.9, G)])
    ref = oracle.brr_chain(N, M, T, K, G, SR, n_iter, sp, y, groups, mS, tape, sigmaG0, usebed=usebed, bed=bed_from_lists(sp, N),
                           hyper_seed=hseed, fh=dict(fh, state0=state0, seed=seed), dirichlet=dirichlet, group_priors=group_priors, covariates=X)
    totals = {}
    with hydra_b200.GenotypeStore(N, M, tasks=T, sync_rate=SR, n_groups=G, n_mix=K, repr_mode=repr_mode, n_slices=n_slices,
                                  threshold_fnz=0.35) as st:
        st.load_data_from_bed(bed)
        st.finalize()
        mk = lambda: hydra_b200.BayesRRm(st, y, mS, groups=groups, sigmaG0=sigmaG0, seed=seed, fh=dict(fh, state0=state0),
                                         dirichlet_priors=dirichlet, group_priors=group_priors, covariates=X)
        brr = mk()
        blob = None
        for it in range(n_iter):
            if restart and it == n_iter - 1:
                blob = brr.save_state()
            tp = None
            if replay:
                tp = dict(zmu=tape["zmu"][it], perm=tape["perm"][it], u=tape["u"][it], z=tape["z"][it], gnu=tape["gnu"][it], glam=tape["glam"][it],
                          fh_hyper=tape["fh_hyper"][it], sigmaG=ref["sigmaG"][it], pi=ref["pi"][it], sigmaE=ref["sigmaE"][it:it + 1])
                if n_cov:
                    tp.update(xI=tape["xI"][it], zcov=tape["zcov"][it])
            o = brr.iteration(tp)
            if n_cov:
                np.testing.assert_allclose(brr.gamma()[0], ref["gamma"][it], rtol=RTOL, atol=1e-14, err_msg=f"gamma it {it}")
            beta, comp, acum = brr.state()
            h, f = brr.hyper(), brr.fh_state()
            assert np.array_equal(comp, ref["comp"][it]), f"components differ at iteration {it}"
            np.testing.assert_allclose(beta, ref["beta"][it], rtol=RTOL, atol=1e-15, err_msg=f"beta it {it}")
            np.testing.assert_allclose(acum, ref["acum"][it], rtol=1e-9, atol=1e-300)
            np.testing.assert_allclose(f["lambda_var"], ref["lambda"][it], rtol=RTOL, err_msg=f"lambda_var it {it}")
            np.testing.assert_allclose(f["nu_var"], ref["nu"][it], rtol=RTOL, err_msg=f"nu_var it {it}")
            np.testing.assert_allclose([f["hypTau"], f["tau"], f["scaledBSQN"]], ref["fh"][it, :3], rtol=RTOL)
            np.testing.assert_allclose(f["c_slab"], ref["fh"][it, 3:], rtol=RTOL)
            assert np.array_equal(h["cass"], ref["cass"][it])
            np.testing.assert_allclose(h["sigmaG"], ref["sigmaG"][it], rtol=RTOL)   # = beta_squaredNorm (:2565)
            np.testing.assert_allclose(h["pi"], ref["pi"][it], rtol=RTOL)
            np.testing.assert_allclose(h["sigmaE"], ref["sigmaE"][it], rtol=RTOL)
            np.testing.assert_allclose(o["e_sqn"], ref["esqn"][it], rtol=RTOL)
            assert o["n_sync"] == ref["nsync"][it]
            for t in range(T):
                np.testing.assert_allclose(brr.task_epsilon(t), ref["eps"][it, t], rtol=RTOL, atol=1e-12, err_msg=f"eps it {it} task {t}")
            for k in ("windows_ahead", "draws_repeated", "markers_changed"):
                totals[k] = totals.get(k, 0) + o[k]
        if restart:   # the state blob carries tau, hypTau, c_slab and the local scales: the continued run is bit-identical
            want = (brr.state()[0].copy(), brr.fh_state())
            b2 = mk()
            b2.load_state(blob)
            b2.iteration(tp)
            got = (b2.state()[0], b2.fh_state())
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1]["lambda_var"], want[1]["lambda_var"])
            assert got[1]["tau"] == want[1]["tau"] and np.array_equal(got[1]["c_slab"], want[1]["c_slab"])
        assert (ref["comp"] > 0).any()
    return totals


def test_fh_replay_single_task_sync_every_marker():
    _case(N=1200, M=300, T=1, SR=1, G=1, K=3, repr_mode="sparse", n_iter=4, seed=31, replay=True)


def test_fh_replay_groups_tasks_windows_mixed_with_restart():
    t = _case(N=1800, M=1024, T=4, SR=5, G=2, K=4, repr_mode="mixed", n_iter=5, seed=32, replay=True, n_slices=3, n_causal=6, restart=True,
              dirichlet=np.array([[2.0, 1.0, 1.0, 0.5], [1.0, 1.0, 3.0, 1.0]]))
    assert t["windows_ahead"] > 0, t


def test_fh_rng_spec_v1_device_gammas_and_host_hyper_draws():
    """tape = None: the per-marker gammas come from the device's Philox Marsaglia-Tsang, hypTau / tau / c_slab (and their initial
    values, :1147-1154) from the hyper-parameter stream; the oracle draws the same spec on the CPU."""
    _case(N=1000, M=400, T=2, SR=3, G=2, K=3, repr_mode="sparse", n_iter=4, seed=2025, replay=False, restart=True)


@pytest.mark.parametrize("replay", [True, False])
def test_fh_with_covariates_and_group_priors(replay):
    """bayesFHMPI together with --covariates, --groupPriorsFile and --dPriorsFile: draw order of the hyper-parameter stream per
    iteration = per group {hypTau, tau, c_slab, pi}, then the covariates' shuffle and normals, then sigmaE."""
    _case(N=1300, M=300, T=3, SR=4, G=2, K=4, repr_mode="sparse", n_iter=4, seed=77, replay=replay, n_cov=3, restart=True,
          group_priors=np.array([[4.0, 0.2], [2.5, 0.05]]), dirichlet=np.array([[5.0, 1.0, 1.0, 1.0], [1.0, 2.0, 2.0, 0.5]]))


def test_fh_rng_spec_shape_below_one():
    _case(N=700, M=120, T=1, SR=2, G=1, K=3, repr_mode="bed", n_iter=3, seed=9, replay=False, fhp=dict(v0L=0.6, v0t=2.0, tau0=0.5))


def test_group_and_dirichlet_priors_enter_the_hyper_draws():
    """--groupPriorsFile / --dPriorsFile (BayesRRm): v0G, s02G per group in the sigmaG draw, Dirichlet parameters in the pi draw."""
    import hydra_b200
    N, M, T, SR, G, K, n_iter, seed = 900, 200, 2, 3, 2, 4, 4, 77
    rng = np.random.default_rng(seed)
    bed, g = random_bed(rng, M, N)
    sp = reference_lists(bed, N)
    y = simulate_y(rng, g, n_causal=20)
    groups = (np.arange(M) % G).astype(np.int32)
    mS = np.tile(np.array([0.0, 0.001, 0.01, 0.1]), (G, 1))
    sigmaG0 = np.array([0.3, 0.6])
    pri = np.array([[4.0, 0.2], [2.5, 0.05]])
    dp = np.array([[5.0, 1.0, 1.0, 1.0], [1.0, 2.0, 2.0, 0.5]])
    tape = oracle.TapeMaker(seed, T, M).make(n_iter)
    ref = oracle.brr_chain(N, M, T, K, G, SR, n_iter, sp, y, groups, mS, tape, sigmaG0, hyper_seed=(seed ^ 0x5bd1e995) & 0xFFFFFFFF,
                           group_priors=pri, dirichlet=dp)
    with hydra_b200.GenotypeStore(N, M, tasks=T, sync_rate=SR, n_groups=G, n_mix=K) as st:
        st.load_data_from_bed(bed)
        st.finalize()
        brr = hydra_b200.BayesRRm(st, y, mS, groups=groups, sigmaG0=sigmaG0, seed=seed, group_priors=pri, dirichlet_priors=dp)
        for it in range(n_iter):
            brr.iteration()
            h = brr.hyper()
            assert np.array_equal(brr.state()[1], ref["comp"][it])
            np.testing.assert_allclose(h["sigmaG"], ref["sigmaG"][it], rtol=RTOL)
            np.testing.assert_allclose(h["pi"], ref["pi"][it], rtol=RTOL)
            np.testing.assert_allclose(brr.state()[0], ref["beta"][it], rtol=RTOL, atol=1e-15)


def test_fh_errors_are_loud():
    import hydra_b200
    rng = np.random.default_rng(1)
    bed, g = random_bed(rng, 20, 200)
    with hydra_b200.GenotypeStore(200, 20, n_mix=3) as st:
        st.load_data_from_bed(bed)
        st.finalize()
        with pytest.raises(hydra_b200.HydraError, match="positive"):
            hydra_b200.BayesRRm(st, simulate_y(rng, g, 3), np.array([[0.0, 0.01, 0.1]]), fh=dict(v0L=-1.0))
        brr = hydra_b200.BayesRRm(st, simulate_y(rng, g, 3), np.array([[0.0, 0.01, 0.1]]), fh={})
        tm = oracle.TapeMaker(1, 1, 20).make(1)
        with pytest.raises(hydra_b200.HydraError, match="gnu"):
            brr.iteration(dict(zmu=tm["zmu"][0], perm=tm["perm"][0], u=tm["u"][0], z=tm["z"][0]))
