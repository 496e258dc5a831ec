"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Integer/byte work is compared bit-exact; floating point at rtol 1e-10 (north_star: "within 1e-10
relative per iteration, in double precision") with a tiny absolute floor for values that are
sums with cancellation.
"""
import numpy as np
import pytest

import oracle
from helpers import bed_from_lists, compact_lists, random_bed, reference_lists, simulate_y

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _store(N, M, **kw):
    import hydra_b200
    return hydra_b200.GenotypeStore(N, M, **kw)


def _assert_lists_equal(exported, sp):
    ref = compact_lists(sp)
    for w in range(3):
        I, S, L = exported[3 * w: 3 * w + 3]
        assert np.array_equal(L, ref[w][2]), f"class {w} lengths"
        assert np.array_equal(S, ref[w][1]), f"class {w} starts"
        assert np.array_equal(I, ref[w][0]), f"class {w} indices"


# ------------------------------------------------------------------------------------ staging
@pytest.mark.parametrize("repr_mode", ["sparse", "bed", "mixed"])
@pytest.mark.parametrize("N,M,n_na,n_slices", [(1003, 37, 0, 0), (1003, 37, 17, 3), (4, 3, 0, 0), (70001, 5, 100, 0), (257, 9, 1, 5)])
def test_staging_bit_exact(repr_mode, N, M, n_na, n_slices):
    rng = np.random.default_rng(N * 7 + M)
    bed, g = random_bed(rng, M, N, pmiss=0.03)
    na = np.sort(rng.choice(N, size=n_na, replace=False)).astype(np.uint32)
    sp = reference_lists(bed, N, na)
    Nc = N - n_na
    with _store(N, M, na_inds=na, repr_mode=repr_mode, n_slices=n_slices, threshold_fnz=0.5) as st:
        assert st.n_ind == Nc
        st.load_data_from_bed(bed[: M // 2], 0)
        st.load_data_from_bed(bed[M // 2:], M // 2)
        st.finalize()
        n1, n2, nm = st.marker_counts()
        assert np.array_equal(n1, sp.N1L) and np.array_equal(n2, sp.N2L) and np.array_equal(nm, sp.NML)
        fnz = (sp.N1L + sp.N2L + sp.NML).astype(np.float64) / Nc
        want_bed = {"sparse": np.zeros(M, bool), "bed": np.ones(M, bool), "mixed": fnz > 0.5}[repr_mode]
        assert np.array_equal(st.marker_is_bed().astype(bool), want_bed)
        _assert_lists_equal(st.export_sparse(), sp)
        ref_bed = bed_from_lists(sp, Nc)
        for m in (0, M - 1):
            assert np.array_equal(st.export_bed(m), ref_bed[m])
        mave, mstd = st.marker_stats()
        omave, omstd = oracle.marker_stats_brr(Nc, sp.N1L, sp.N2L, sp.NML)
        assert np.array_equal(mave, omave) and np.array_equal(mstd, omstd)


def test_staging_from_sparse_lists_matches_bed():
    rng = np.random.default_rng(5)
    N, M = 2051, 23
    bed, _ = random_bed(rng, M, N, pmiss=0.02)
    na = np.array([0, 5, 6, 2050], np.uint32)
    raw = oracle.sparse_fill_indices(bed, N)          # raw lists, as read from .si?/.ss?/.sl? files
    sp = reference_lists(bed, N, na)
    with _store(N, M, na_inds=na, repr_mode="mixed", threshold_fnz=0.3) as st:
        st.load_data_from_sparse(raw.I1, raw.N1S, raw.N1L, raw.I2, raw.N2S, raw.N2L, raw.IM, raw.NMS, raw.NML)
        st.finalize()
        _assert_lists_equal(st.export_sparse(), sp)


def test_synthetic_generator_matches_oracle():
    from hydra_b200 import synth
    N, M = 1501, 40
    p = synth.maf_spectrum(M, 0.01, 0.5, seed=11)
    thr = synth.thresholds(p)
    othr = np.stack([oracle.synth_thresholds(float(x)) for x in p])
    assert np.array_equal(thr, othr)
    att = np.arange(M, dtype=np.uint32) % 3
    obed = oracle.synth_bed(1234, N, thr, att)
    sp = reference_lists(obed, N)
    with _store(N, M, repr_mode="sparse") as st:
        st.load_synthetic(1234, thr, att)
        st.finalize()
        _assert_lists_equal(st.export_sparse(), sp)


# ------------------------------------------------------------------------------------ unit kernels
@pytest.mark.parametrize("repr_mode", ["sparse", "bed"])
@pytest.mark.parametrize("N,M,n_slices", [(1003, 31, 0), (1003, 31, 4), (70001, 6, 0)])
def test_dot_and_scaadd(repr_mode, N, M, n_slices):
    rng = np.random.default_rng(N + M)
    bed, _ = random_bed(rng, M, N, pmiss=0.02)
    sp = reference_lists(bed, N)
    eps = rng.normal(size=N)
    with _store(N, M, repr_mode=repr_mode, n_slices=n_slices) as st:
        st.load_data_from_bed(bed)
        st.finalize()
        mave, mstd = st.marker_stats()
        st.set_epsilon(eps)
        markers = rng.permutation(M).astype(np.uint32)
        num = st.sparse_dotprod(markers)
        want = np.array([oracle.sparse_dotprod(eps, sp, int(m), mave[m], mstd[m]) for m in markers])
        scale = np.abs(eps).sum() / N * np.sqrt(N)  # |num| is a sum with cancellation: floor ~ 1e-13*sqrt(N)
        np.testing.assert_allclose(num, want, rtol=RTOL, atol=1e-12 * scale)
        if repr_mode == "bed":
            want2 = np.array([oracle.lut_dotprod(bed[m], eps, mave[m], mstd[m]) for m in markers])
            np.testing.assert_allclose(num, want2, rtol=RTOL, atol=1e-12 * scale)
        # epsilon update: eps' = eps + sum_j deltaEps_j
        upd = markers[: min(7, M)]
        db = rng.normal(size=len(upd)) * 0.01
        st.sparse_scaadd(upd, db)
        got = st.get_epsilon()
        ref = eps.copy()
        for m, d in zip(upd, db):
            if repr_mode == "bed":
                ref += oracle.lut_scaadd(N, bed[m], d, mave[m], mstd[m])
            else:
                ref += oracle.sparse_scaadd(N, d, sp, int(m), mave[m], mstd[m])
        np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-12)   # epsilon lives on a fixed-point grid of 2^-45 .. 2^-51
        # the dot product sees the updated residual
        num2 = st.sparse_dotprod(markers[:5])
        want3 = np.array([oracle.sparse_dotprod(ref, sp, int(m), mave[m], mstd[m]) for m in markers[:5]])
        np.testing.assert_allclose(num2, want3, rtol=RTOL, atol=1e-12 * scale)


# ------------------------------------------------------------------------------------ chain replay
def _run_chain_case(N, M, T, SR, G, K, repr_mode, n_iter, seed, n_na=0, n_slices=0, replay_hyper=True, max_ctas=0,
                    threshold_fnz=0.35, n_causal=None):
    import hydra_b200
    rng = np.random.default_rng(seed)
    bed, g = random_bed(rng, M, N, pmiss=0.01)
    na = np.sort(rng.choice(N, size=n_na, replace=False)).astype(np.uint32)
    sp = reference_lists(bed, N, na)
    Nc = N - n_na
    keep = np.setdiff1d(np.arange(N), na)
    y = simulate_y(rng, g[:, keep], n_causal=max(3, M // 10) if n_causal is None else n_causal)
    groups = (np.arange(M) % G).astype(np.int32)
    mS = np.tile(np.array([0.0] + [10.0 ** (-(K - 1 - k)) for k in range(1, K)]), (G, 1))
    mS[:, 1:] *= (1.0 + 0.5 * np.arange(G))[:, None]
    sigmaG0 = rng.uniform(0.2, 0.8, size=G)
    tape = oracle.TapeMaker(seed, T, M).make(n_iter)

    fnz = (sp.N1L + sp.N2L + sp.NML).astype(np.float64) / Nc
    usebed = {"sparse": np.zeros(M, np.uint8), "bed": np.ones(M, np.uint8), "mixed": (fnz > threshold_fnz).astype(np.uint8)}[repr_mode]
    obed = bed_from_lists(sp, Nc)
    hyper_seed = (seed ^ 0x5bd1e995) & 0xFFFFFFFF
    ref = oracle.brr_chain(Nc, M, T, K, G, SR, n_iter, sp, y, groups, mS, tape, sigmaG0, usebed=usebed, bed=obed,
                           hyper_seed=hyper_seed if not replay_hyper else 0)
    # with replay_hyper the GPU is fed the oracle's hyper-parameter VALUES; otherwise both sides draw with RNG spec v1,
    # where the oracle consumes sigmaG0 from the tape and the product would draw it: pass it explicitly.
    with hydra_b200.GenotypeStore(N, M, na_inds=na, tasks=T, sync_rate=SR, n_groups=G, n_mix=K, repr_mode=repr_mode,
                                  n_slices=n_slices, max_ctas=max_ctas, threshold_fnz=threshold_fnz) as st:
        st.load_data_from_bed(bed)
        st.finalize()
        assert np.array_equal(st.marker_is_bed(), usebed)
        brr = hydra_b200.BayesRRm(st, y, mS, groups=groups, sigmaG0=sigmaG0, seed=seed)
        totals = {}
        for it in range(n_iter):
            tp = dict(zmu=tape["zmu"][it], perm=tape["perm"][it], u=tape["u"][it], z=tape["z"][it])
            if replay_hyper:
                tp.update(sigmaG=ref["sigmaG"][it], pi=ref["pi"][it], sigmaE=ref["sigmaE"][it:it + 1])
            o = brr.iteration(tp)
            beta, comp, acum = brr.state()
            h = brr.hyper()
            assert np.array_equal(comp, ref["comp"][it]), f"components differ at iteration {it}"
            np.testing.assert_allclose(beta, ref["beta"][it], rtol=RTOL, atol=1e-15, err_msg=f"beta it {it}")
            np.testing.assert_allclose(acum, ref["acum"][it], rtol=1e-9, atol=1e-300, err_msg=f"acum it {it}")
            assert np.array_equal(h["cass"], ref["cass"][it])
            np.testing.assert_allclose(h["mu"], ref["mu"][it], rtol=RTOL, atol=1e-13)
            np.testing.assert_allclose(h["bsq"], ref["bsq"][it], rtol=RTOL)
            np.testing.assert_allclose(o["e_sqn"], ref["esqn"][it], rtol=RTOL)
            assert o["n_sync"] == ref["nsync"][it]
            for k in ("windows_ahead", "draws_repeated", "n_windows", "markers_changed"):
                totals[k] = totals.get(k, 0) + o[k]
            for t in range(T):
                np.testing.assert_allclose(brr.task_epsilon(t), ref["eps"][it, t], rtol=RTOL, atol=1e-12, err_msg=f"eps it {it} task {t}")
            np.testing.assert_allclose(h["sigmaE"], ref["sigmaE"][it], rtol=RTOL)
            np.testing.assert_allclose(h["sigmaG"], ref["sigmaG"][it], rtol=RTOL)
            np.testing.assert_allclose(h["pi"], ref["pi"][it], rtol=RTOL)
    return totals


def test_chain_replay_windows_run_ahead():
    # Few causal markers: after a few iterations most steps change nothing, the reference synchronises after every step
    # (:2044-2050) and the kernel takes sync_rate such steps at a time, discarding and repeating the steps behind the first
    # change. Same per-iteration parity as every other case, and both outcomes of a window run ahead must have occurred.
    for repr_mode, seed in (("sparse", 11), ("mixed", 12)):
        t = _run_chain_case(N=1800, M=1536, T=4, SR=5, G=2, K=3, repr_mode=repr_mode, n_iter=6, seed=seed, n_causal=4, n_slices=3)
        assert t["windows_ahead"] > 20 and t["draws_repeated"] > 0, t


def test_chain_replay_config1_single_task_bed():
    # BASELINE config 1 shape (reduced): BED input, 1 task, sync every marker, 1 group x 4 mixtures
    _run_chain_case(N=1200, M=300, T=1, SR=1, G=1, K=4, repr_mode="bed", n_iter=4, seed=1222)


def test_chain_replay_config2_groups_4tasks_sr10():
    # BASELINE config 2 shape (reduced): 2 groups x 4 mixtures, 4 tasks, sync_rate 10, sparse lists
    _run_chain_case(N=1500, M=403, T=4, SR=10, G=2, K=4, repr_mode="sparse", n_iter=4, seed=7)


def test_chain_replay_mixed_na_sliced():
    _run_chain_case(N=2000, M=250, T=3, SR=5, G=3, K=3, repr_mode="mixed", n_iter=3, seed=99, n_na=13, n_slices=4)


def test_chain_replay_many_tasks_few_ctas():
    # more window positions than CTA groups: every group works through several items per window
    _run_chain_case(N=900, M=1000, T=16, SR=8, G=1, K=4, repr_mode="sparse", n_iter=2, seed=3, n_slices=2, max_ctas=8)


@pytest.mark.parametrize("K", [2, 7])
def test_chain_replay_other_mixture_counts(K):
    # K = 7: more cascade terms than one warp evaluates at once (K*K > 32), the sequential form of :1894-1921 runs
    _run_chain_case(N=1100, M=260, T=3, SR=4, G=2, K=K, repr_mode="sparse", n_iter=3, seed=40 + K)


def test_chain_replay_heavy_blocks_one_group():
    # common variants on one CTA group: slice blocks of thousands of words, cut in work units longer than one 128-word step
    _run_chain_case(N=40000, M=48, T=8, SR=4, G=1, K=4, repr_mode="sparse", n_iter=2, seed=11, max_ctas=2)


def test_chain_rng_spec_v1_matches_oracle_draws():
    # no hyper-parameter replay: both sides draw sigmaG / pi / sigmaE with RNG spec v1 from the same seed
    _run_chain_case(N=1000, M=200, T=2, SR=3, G=2, K=4, repr_mode="sparse", n_iter=4, seed=2024, replay_hyper=False)


def test_chain_device_rng_equals_tape():
    """tape=None (device Philox + host mt19937 streams) reproduces the oracle fed with TapeMaker's tape."""
    import hydra_b200
    N, M, T, SR, G, K, seed, n_iter = 800, 150, 3, 4, 1, 4, 555, 3
    rng = np.random.default_rng(1)
    bed, g = random_bed(rng, M, N)
    sp = reference_lists(bed, N)
    y = simulate_y(rng, g)
    mS = np.array([[0.0, 0.001, 0.01, 0.1]])
    sigmaG0 = np.array([0.5])
    tape = oracle.TapeMaker(seed, T, M).make(n_iter)
    ref = oracle.brr_chain(N, M, T, K, G, SR, n_iter, sp, y, np.zeros(M, np.int32), mS, tape, sigmaG0,
                           hyper_seed=(seed ^ 0x5bd1e995) & 0xFFFFFFFF)
    with hydra_b200.GenotypeStore(N, M, tasks=T, sync_rate=SR, n_groups=G, n_mix=K) as st:
        st.load_data_from_bed(bed)
        st.finalize()
        brr = hydra_b200.BayesRRm(st, y, mS, sigmaG0=sigmaG0, seed=seed)
        for it in range(n_iter):
            brr.iteration()
            beta, comp, _ = brr.state()
            assert np.array_equal(comp, ref["comp"][it])
            np.testing.assert_allclose(beta, ref["beta"][it], rtol=RTOL, atol=1e-15)
            np.testing.assert_allclose(brr.hyper()["sigmaE"], ref["sigmaE"][it], rtol=RTOL)
            for t in range(T):
                assert np.array_equal(brr.task_perm(t), tape["perm"][it][st.task_blocks()[0][t]: st.task_blocks()[0][t] + st.task_blocks()[1][t]])


def test_errors_are_loud():
    import hydra_b200
    with pytest.raises(hydra_b200.HydraError):
        hydra_b200.GenotypeStore(100, 10, tasks=20)  # more tasks than markers
    with hydra_b200.GenotypeStore(100, 10) as st:
        with pytest.raises(hydra_b200.HydraError):
            st.finalize()  # nothing staged


# ------------------------------------------------------------------------------------ BASELINE configs 1 and 2 at their real size
def _example_scale_case(T, SR, G, repr_mode, seed, n_iter=3):
    """N = 5 000, M = 10 000 (the size of example/t_M10K_N_5K, whose .bed is missing from the reference checkout):
    common-variant stand-in with the example's .mS rows (0.001,0.01,0.1), 200 causal markers, h2 ~ 0.5."""
    import hydra_b200
    N, M, K = 5000, 10000, 4
    rng = np.random.default_rng(seed)
    bed, g = random_bed(rng, M, N, maf_lo=0.01, maf_hi=0.5, pmiss=0.002)
    sp = reference_lists(bed, N)
    causal = rng.choice(M, size=200, replace=False)
    x = np.where(g[causal] < 0, 0, g[causal]).astype(np.float64)
    x = (x - x.mean(1, keepdims=True)) / x.std(1, keepdims=True)
    y = x.T @ rng.normal(0, np.sqrt(0.5 / 200), size=200) + rng.normal(0, np.sqrt(0.5), size=N)
    groups = (np.arange(M) >= M // 2).astype(np.int32) if G == 2 else np.zeros(M, np.int32)   # example/normal.group: two blocks of 5000
    mS = np.tile(np.array([0.0, 0.001, 0.01, 0.1]), (G, 1))
    sigmaG0 = rng.uniform(0.2, 0.8, size=G)
    tape = oracle.TapeMaker(seed, T, M).make(n_iter)
    usebed = np.ones(M, np.uint8) if repr_mode == "bed" else np.zeros(M, np.uint8)
    ref = oracle.brr_chain(N, M, T, K, G, SR, n_iter, sp, y, groups, mS, tape, sigmaG0, usebed=usebed,
                           bed=bed_from_lists(sp, N) if repr_mode == "bed" else None, hyper_seed=(seed ^ 0x5bd1e995) & 0xFFFFFFFF)
    with hydra_b200.GenotypeStore(N, M, tasks=T, sync_rate=SR, n_groups=G, n_mix=K, repr_mode=repr_mode) as st:
        st.load_data_from_bed(bed)
        st.finalize()
        brr = hydra_b200.BayesRRm(st, y, mS, groups=groups, sigmaG0=sigmaG0, seed=seed)
        for it in range(n_iter):
            o = brr.iteration()
            beta, comp, acum = brr.state()
            h = brr.hyper()
            assert np.array_equal(comp, ref["comp"][it]), f"components differ at iteration {it}"
            assert np.array_equal(h["cass"], ref["cass"][it]) and o["n_sync"] == ref["nsync"][it]
            np.testing.assert_allclose(beta, ref["beta"][it], rtol=RTOL, atol=1e-15, err_msg=f"beta it {it}")
            np.testing.assert_allclose(h["bsq"], ref["bsq"][it], rtol=RTOL)
            np.testing.assert_allclose(o["e_sqn"], ref["esqn"][it], rtol=RTOL)
            np.testing.assert_allclose(h["sigmaG"], ref["sigmaG"][it], rtol=RTOL)
            np.testing.assert_allclose(h["pi"], ref["pi"][it], rtol=RTOL)
            for t in range(T):
                np.testing.assert_allclose(brr.task_epsilon(t), ref["eps"][it, t], rtol=RTOL, atol=1e-12, err_msg=f"eps it {it} task {t}")
        assert (beta != 0).sum() > 10


def test_chain_replay_config1_full_size():
    # BASELINE config 1 as written: BED input, M = 10 000, N = 5 000, 1 task, sync after every marker
    _example_scale_case(T=1, SR=1, G=1, repr_mode="bed", seed=1222)


def test_chain_replay_config2_full_size():
    # BASELINE config 2 as written: grouped mixtures (2 groups), 4 tasks, sync_rate 10
    _example_scale_case(T=4, SR=10, G=2, repr_mode="sparse", seed=1223)


def test_restart_keeps_a_switched_off_group_off():
    """a group whose sigmaG was set to 0 (m0 == 0, src/BayesRRm.cpp:2534-2542) stays switched off after
    load_state (adaV re-derived from the restored sigmaG, :1592-1597): continued and uninterrupted runs are identical."""
    import hydra_b200
    N, M, T, SR, G, K = 700, 60, 2, 3, 2, 3
    rng = np.random.default_rng(8)
    bed, g = random_bed(rng, M, N)
    y = simulate_y(rng, g, n_causal=5)
    groups = np.zeros(M, np.int32)
    groups[7] = 1                              # group 1 = one null marker: it draws component 0 sooner or later
    mS = np.tile(np.array([0.0, 0.001, 0.01]), (G, 1))
    with hydra_b200.GenotypeStore(N, M, tasks=T, sync_rate=SR, n_groups=G, n_mix=K) as st:
        st.load_data_from_bed(bed)
        st.finalize()
        brr = hydra_b200.BayesRRm(st, y, mS, groups=groups, sigmaG0=np.array([0.5, 0.5]), seed=3)
        for it in range(60):
            brr.iteration()
            if brr.hyper()["sigmaG"][1] == 0.0:
                break
        assert brr.hyper()["sigmaG"][1] == 0.0, "group 1 was never switched off"
        blob = brr.save_state()
        straight = []
        for _ in range(4):
            brr.iteration()
            straight.append((brr.state()[0].copy(), brr.hyper()["cass"].copy(), brr.hyper()["sigmaG"].copy()))
        brr2 = hydra_b200.BayesRRm(st, y, mS, groups=groups, sigmaG0=np.array([0.5, 0.5]), seed=3)
        brr2.load_state(blob)
        for k in range(4):
            brr2.iteration()
            h = brr2.hyper()
            assert np.array_equal(brr2.state()[0], straight[k][0]) and np.array_equal(h["cass"], straight[k][1])
            assert np.array_equal(h["sigmaG"], straight[k][2]) and h["sigmaG"][1] == 0.0


# ------------------------------------------------------------------------------------ fixed effects (--covariates)
@pytest.mark.parametrize("replay_hyper", [True, False])
def test_chain_replay_with_covariates(replay_hyper):
    """src/BayesRRm.cpp:2648-2681: per iteration, between the group hyper-parameters and sigmaE, every covariate's gamma is
    drawn and epsilon updated. Replay: the oracle's hyper-parameter values and the covariate order / normals come from the
    tape; otherwise both sides draw them from the hyper-parameter stream of RNG spec v1 (order: sigmaG, pi per group, the
    shuffle of the covariates, their normals, sigmaE)."""
    import hydra_b200
    N, M, T, SR, G, K, F, n_iter, seed = 1300, 220, 3, 4, 2, 4, 3, 4, 61
    rng = np.random.default_rng(seed)
    bed, g = random_bed(rng, M, N, pmiss=0.01)
    sp = reference_lists(bed, N)
    X = rng.normal(size=(N, F))
    X = (X - X.mean(0)) / X.std(0, ddof=1)
    X[:, 2] += 0.3                                  # one column that is not centred: the residual of (global) task 0 matters
    y = simulate_y(rng, g, n_causal=12) + X @ np.array([0.8, -0.5, 0.0])
    groups = (np.arange(M) % G).astype(np.int32)
    mS = np.tile(np.array([0.0, 0.001, 0.01, 0.1]), (G, 1))
    sigmaG0 = np.array([0.3, 0.6])
    tape = oracle.TapeMaker(seed, T, M).make(n_iter)
    if replay_hyper:
        tape["xI"] = np.array([rng.permutation(F) for _ in range(n_iter)], np.int32)
        tape["zcov"] = rng.normal(size=(n_iter, F))
    hseed = (seed ^ 0x5bd1e995) & 0xFFFFFFFF
    ref = oracle.brr_chain(N, M, T, K, G, SR, n_iter, sp, y, groups, mS, tape, sigmaG0, hyper_seed=hseed, covariates=X)
    with hydra_b200.GenotypeStore(N, M, tasks=T, sync_rate=SR, n_groups=G, n_mix=K) as st:
        st.load_data_from_bed(bed)
        st.finalize()
        brr = hydra_b200.BayesRRm(st, y, mS, groups=groups, sigmaG0=sigmaG0, seed=seed, covariates=X)
        for it in range(n_iter):
            tp = dict(zmu=tape["zmu"][it], perm=tape["perm"][it], u=tape["u"][it], z=tape["z"][it])
            if replay_hyper:
                tp.update(sigmaG=ref["sigmaG"][it], pi=ref["pi"][it], sigmaE=ref["sigmaE"][it:it + 1], xI=tape["xI"][it], zcov=tape["zcov"][it])
            o = brr.iteration(tp)
            beta, comp, _ = brr.state()
            gam, xI = brr.gamma()
            h = brr.hyper()
            assert np.array_equal(comp, ref["comp"][it]), f"components differ at iteration {it}"
            np.testing.assert_allclose(gam, ref["gamma"][it], rtol=RTOL, atol=1e-14, err_msg=f"gamma it {it}")
            np.testing.assert_allclose(beta, ref["beta"][it], rtol=RTOL, atol=1e-15)
            np.testing.assert_allclose(o["e_sqn"], ref["esqn"][it], rtol=RTOL)
            np.testing.assert_allclose(h["sigmaE"], ref["sigmaE"][it], rtol=RTOL)
            for t in range(T):
                np.testing.assert_allclose(brr.task_epsilon(t), ref["eps"][it, t], rtol=RTOL, atol=1e-12, err_msg=f"eps it {it} task {t}")
        assert abs(gam[0]) > 0.2 and abs(gam[1]) > 0.1     # the simulated fixed effects are found
        # restart keeps the fixed effects
        blob = brr.save_state()
        brr.iteration()
        want = (brr.gamma()[0].copy(), brr.state()[0].copy())
        brr2 = hydra_b200.BayesRRm(st, y, mS, groups=groups, sigmaG0=sigmaG0, seed=seed, covariates=X)
        brr2.load_state(blob)
        brr2.iteration()
        assert np.array_equal(brr2.gamma()[0], want[0]) and np.array_equal(brr2.state()[0], want[1])


def test_async_state_readback_equals_synchronous():
    """hb_brr_get_state_async / hb_brr_state_wait: the read-back of iteration i overlaps iteration i+1 and still returns the
    state of iteration i."""
    import hydra_b200
    rng = np.random.default_rng(8)
    N, M = 1500, 600
    bed, g = random_bed(rng, M, N, pmiss=0.01)
    y = simulate_y(rng, g, n_causal=30)
    def run(asynchronous):
        res = []
        with hydra_b200.GenotypeStore(N, M, tasks=4, sync_rate=5, n_groups=1, n_mix=4, repr_mode="sparse") as st:
            st.load_data_from_bed(bed)
            st.finalize()
            brr = hydra_b200.BayesRRm(st, y, [[0.001, 0.01, 0.1]], seed=5)
            bufs = [(np.zeros(M), np.zeros(M, np.int32), np.zeros(M)) for _ in range(2)]
            for it in range(4):
                brr.iteration()
                if asynchronous:
                    done = brr.state_wait()
                    if done is not None:
                        res.append(tuple(a.copy() for a in done))
                    brr.state_async(bufs[it & 1])
                else:
                    res.append(tuple(a.copy() for a in brr.state()))
            if asynchronous:
                res.append(tuple(a.copy() for a in brr.state_wait()))
        return res
    a, b = run(True), run(False)
    assert len(a) == len(b) == 4
    for it in range(4):
        for x, yv in zip(a[it], b[it]):
            assert np.array_equal(x, yv), it
