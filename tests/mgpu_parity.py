"""Multi-GPU replay parity (launched by torchrun from tests/test_gpu_multi.py, one process per GPU):
G GPUs x T_local tasks each must reproduce the CPU oracle run with T_total = G*T_local tasks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hydra_b200  # noqa: E402
import oracle  # noqa: E402
from helpers import bed_from_lists, random_bed, reference_lists, simulate_y  # noqa: E402


def bayesw_cases(rank, world, lr):
    """BayesW on several GPUs (src/BayesW.cpp:1645, 1799-1835, 1866-1867): markers and tasks split over the GPUs, epsilon replicated,
    the window's epsilon changes summed with ncclAllReduce. G GPUs x T_local tasks = the oracle with T_total tasks."""
    if oracle.arms_ref() is None:
        if rank == 0:
            print("BayesW multi-GPU parity skipped: oracle/_ref/libarms_ref.so not built", flush=True)
        return
    EUM = 0.577215664901532
    for case, (N, M, TL, SR, G, K, repr_mode, n_iter, seed, replay_hyper) in enumerate([
        (900, 96, 2, 3, 2, 4, "sparse", 3, 17, True),
        (1100, 64, 1, 1, 1, 3, "bed", 3, 5, False),
        (1000, 203, 3, 4, 1, 4, "mixed", 2, 29, True),     # ragged task blocks: some tasks run out of markers before others
    ]):
        T, quad = TL * world, 25
        rng = np.random.default_rng(seed)
        bed, g = random_bed(rng, M, N, pmiss=0.01)
        sp = reference_lists(bed, N)
        x = np.where(g < 0, 0, g).astype(np.float64)
        x = (x - x.mean(1, keepdims=True)) / (x.std(1, keepdims=True) + 1e-12)
        causal = rng.choice(M, size=max(3, M // 10), replace=False)
        b = rng.normal(0, np.sqrt(0.3 * (np.pi ** 2 / 6) / 100.0 / len(causal)), size=len(causal))
        y = 4.1 + x[causal].T @ b + np.log(rng.exponential(size=N)) / 10.0 + EUM / 10.0
        fail = (rng.random(N) > 0.1).astype(np.float64)
        groups = (np.arange(M) % G).astype(np.int32)
        mS = np.tile(np.array([0.0] + [10.0 ** (-(K - 1 - k)) for k in range(1, K)]), (G, 1))
        tm = oracle.TapeMaker(seed, T, M).make(n_iter)
        tape = dict(perm=tm["perm"], p=tm["u"])
        ref = oracle.bw_chain(N, M, T, K, G, SR, n_iter, quad, sp, y, fail, groups, mS, tape, seed, hyper_seed=(seed ^ 0x5bd1e995) & 0xFFFFFFFF)
        st = hydra_b200.GenotypeStore(N, M, tasks=T, task_first=rank * TL, tasks_local=TL, sync_rate=SR, n_groups=G, n_mix=K,
                                      repr_mode=repr_mode, threshold_fnz=0.35, device=lr, model="bayesW")
        ms, ml = st.m_start, st.m_local
        st.load_data_from_bed(bed[ms:ms + ml])
        st.finalize()
        st.comm_init(dist)
        bw = hydra_b200.BayesW(st, y, fail, mS, groups=groups, quad_points=quad, seed=seed)
        for it in range(n_iter):
            tp = dict(perm=tape["perm"][it][ms:ms + ml], p=tape["p"][it][ms:ms + ml])
            if replay_hyper:
                tp.update(sigmaG=ref["sigmaG"][it], pi=ref["pi"][it])
            o = bw.iteration(tp)
            beta, comp = bw.state()
            h = bw.hyper()
            np.testing.assert_allclose(o["mu"], ref["mu"][it], rtol=1e-11, err_msg=f"BayesW case {case} mu it {it}")
            np.testing.assert_allclose(o["alpha"], ref["alpha"][it], rtol=1e-11, err_msg=f"BayesW case {case} alpha it {it}")
            assert np.array_equal(comp, ref["comp"][it][ms:ms + ml]), f"BayesW case {case} rank {rank}: components differ at iteration {it}"
            np.testing.assert_allclose(beta, ref["beta"][it][ms:ms + ml], rtol=1e-10, atol=1e-14)
            assert np.array_equal(h["cass"], ref["cass"][it])
            assert o["n_sync"] == ref["nsync"][it], (o["n_sync"], ref["nsync"][it])
            # epsilon: 1e-10 of its scale (|eps| ~ 0.1). An ARMS sample inverts the envelope's cumulative, which amplifies a 1-ulp
            # exp/log difference (tests/test_gpu_bayesw.py: 2e-9 on a sample of ~1e-3); through x * d(beta) that is ~1e-12 absolute
            # on epsilon (measured 1.2e-12 here), visible as a relative error only on entries that happen to be near zero
            np.testing.assert_allclose(bw.epsilon(), ref["eps"][it], rtol=1e-10, atol=1e-11)
            np.testing.assert_allclose(h["sigmaG"], ref["sigmaG"][it], rtol=1e-10)
            np.testing.assert_allclose(h["pi"], ref["pi"][it], rtol=1e-10)
        assert (ref["comp"] > 0).any()
        # the replicas of epsilon are bit-identical on all GPUs
        e = torch.from_numpy(bw.epsilon()).cuda()
        lo, hi = e.clone(), e.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), f"BayesW case {case}: epsilon replicas differ between GPUs"
        st.close()
        dist.barrier()
        if rank == 0:
            print(f"BayesW multi-GPU parity case {case} ok on {world} GPUs", flush=True)


def fh_cases(rank, world, lr):
    """bayesFHMPI on several GPUs: per-marker scales stay with the marker's GPU, the scaled sum of squares is all-reduced with the
    group statistics, tau / hypTau / c_slab are drawn on every GPU from the common stream. Against the oracle with T_total tasks."""
    for case, (N, M, TL, SR, G, K, repr_mode, n_iter, seed, replay) in enumerate([
        (1500, 600, 2, 5, 2, 4, "sparse", 4, 41, True),
        (1200, 403, 3, 4, 1, 3, "mixed", 4, 42, False),
    ]):
        T = TL * world
        rng = np.random.default_rng(seed)
        bed, g = random_bed(rng, M, N, pmiss=0.01)
        sp = reference_lists(bed, N)
        y = simulate_y(rng, g, n_causal=max(3, M // 20))
        groups = (np.arange(M) % G).astype(np.int32)
        mS = np.tile(np.array([0.0] + [10.0 ** (-(K - 1 - k)) for k in range(1, K)]), (G, 1))
        sigmaG0 = rng.uniform(0.2, 0.8, size=G)
        tape = oracle.TapeMaker(seed, T, M).make(n_iter)
        fh = dict(oracle.FH_DEFAULTS)
        state0 = None
        if replay:
            tape["gnu"] = rng.gamma(2.0, size=(n_iter, M))
            tape["glam"] = rng.gamma(2.0, size=(n_iter, M))
            tape["fh_hyper"] = np.stack([rng.uniform(0.5, 3.0, (n_iter, G)), rng.uniform(0.005, 0.05, (n_iter, G)), rng.uniform(0.1, 1.0, (n_iter, G))], axis=2)
            state0 = np.concatenate([[1.7, 0.02], rng.uniform(0.2, 0.9, G)])
        fnz = (sp.N1L + sp.N2L + sp.NML).astype(np.float64) / N
        usebed = {"sparse": np.zeros(M, np.uint8), "mixed": (fnz > 0.35).astype(np.uint8)}[repr_mode]
        ref = oracle.brr_chain(N, M, T, K, G, SR, n_iter, sp, y, groups, mS, tape, sigmaG0, usebed=usebed, bed=bed_from_lists(sp, N),
                               hyper_seed=(seed ^ 0x5bd1e995) & 0xFFFFFFFF, fh=dict(fh, state0=state0, seed=seed))
        st = hydra_b200.GenotypeStore(N, M, tasks=T, task_first=rank * TL, tasks_local=TL, sync_rate=SR, n_groups=G, n_mix=K,
                                      repr_mode=repr_mode, threshold_fnz=0.35, device=lr)
        ms, ml = st.m_start, st.m_local
        st.load_data_from_bed(bed[ms:ms + ml])
        st.finalize()
        st.comm_init(dist)
        brr = hydra_b200.BayesRRm(st, y, mS, groups=groups, sigmaG0=sigmaG0, seed=seed, fh=dict(fh, state0=state0))
        for it in range(n_iter):
            tp = None
            if replay:
                tp = dict(zmu=tape["zmu"][it][rank * TL:(rank + 1) * TL], perm=tape["perm"][it][ms:ms + ml], u=tape["u"][it][ms:ms + ml],
                          z=tape["z"][it][ms:ms + ml], gnu=tape["gnu"][it][ms:ms + ml], glam=tape["glam"][it][ms:ms + ml], fh_hyper=tape["fh_hyper"][it],
                          sigmaG=ref["sigmaG"][it], pi=ref["pi"][it], sigmaE=ref["sigmaE"][it:it + 1])
            o = brr.iteration(tp)
            beta, comp, _ = brr.state()
            h, f = brr.hyper(), brr.fh_state()
            assert np.array_equal(comp, ref["comp"][it][ms:ms + ml]), f"FH case {case} rank {rank}: components differ at iteration {it}"
            np.testing.assert_allclose(beta, ref["beta"][it][ms:ms + ml], rtol=1e-10, atol=1e-15)
            np.testing.assert_allclose(f["lambda_var"], ref["lambda"][it][ms:ms + ml], rtol=1e-10)
            np.testing.assert_allclose(f["nu_var"], ref["nu"][it][ms:ms + ml], rtol=1e-10)
            np.testing.assert_allclose([f["hypTau"], f["tau"], f["scaledBSQN"]], ref["fh"][it, :3], rtol=1e-10)
            np.testing.assert_allclose(f["c_slab"], ref["fh"][it, 3:], rtol=1e-10)
            assert np.array_equal(h["cass"], ref["cass"][it])
            np.testing.assert_allclose(h["sigmaG"], ref["sigmaG"][it], rtol=1e-10)
            np.testing.assert_allclose(h["sigmaE"], ref["sigmaE"][it], rtol=1e-10)
            assert o["n_sync"] == ref["nsync"][it]
            for t in range(TL):
                np.testing.assert_allclose(brr.task_epsilon(t), ref["eps"][it, rank * TL + t], rtol=1e-10, atol=1e-12)
        st.close()
        dist.barrier()
        if rank == 0:
            print(f"bayesFH multi-GPU parity case {case} ok on {world} GPUs", flush=True)


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    if "bayesw" in sys.argv[1:] or "fh" in sys.argv[1:]:
        if "bayesw" in sys.argv[1:]:
            bayesw_cases(rank, world, lr)
        if "fh" in sys.argv[1:]:
            fh_cases(rank, world, lr)
        dist.destroy_process_group()
        return
    for case, (N, M, TL, SR, G, K, repr_mode, n_iter, seed) in enumerate([
        (1500, 403, 2, 10, 2, 4, "sparse", 4, 7),
        (1200, 300, 1, 1, 1, 4, "bed", 3, 1222),
        (2000, 1000, 8, 8, 1, 3, "mixed", 2, 5),
        # dense trait, 64 tasks per GPU x sync rate 10: most of the 640 markers per GPU and window change, i.e. more than
        # 1024 changed markers per window from 2 GPUs on (the old merge gave up there with error 4)
        (600, 4096, 64, 10, 1, 4, "sparse", 3, 31),
        # few causal markers: most steps change nothing, the windows run ahead (sync_rate steps at a time where the reference
        # synchronises after every step) and the first changed step is taken over the lists of all GPUs
        (1800, 1536, 2, 5, 2, 3, "sparse", 6, 11),
    ]):
        T = TL * world
        rng = np.random.default_rng(seed)
        bed, g = random_bed(rng, M, N, pmiss=0.01)
        sp = reference_lists(bed, N)
        y = simulate_y(rng, g, n_causal=M if case == 3 else (4 if case == 4 else max(3, M // 10)), h2=0.9 if case == 3 else 0.5)
        groups = (np.arange(M) % G).astype(np.int32)
        mS = np.tile(np.array([0.0] + [10.0 ** (-(K - 1 - k)) for k in range(1, K)]), (G, 1))
        if case == 3:
            mS = np.array([[0.0, 1e-6, 1e-5, 1e-4]])   # tiny slab variances: about half of the markers take a non-zero effect at every draw
        sigmaG0 = rng.uniform(0.2, 0.8, size=G)
        tape = oracle.TapeMaker(seed, T, M).make(n_iter)
        fnz = (sp.N1L + sp.N2L + sp.NML).astype(np.float64) / N
        usebed = {"sparse": np.zeros(M, np.uint8), "bed": np.ones(M, np.uint8), "mixed": (fnz > 0.35).astype(np.uint8)}[repr_mode]
        ref = oracle.brr_chain(N, M, T, K, G, SR, n_iter, sp, y, groups, mS, tape, sigmaG0, usebed=usebed, bed=bed_from_lists(sp, N))
        st = hydra_b200.GenotypeStore(N, M, tasks=T, task_first=rank * TL, tasks_local=TL, sync_rate=SR, n_groups=G, n_mix=K,
                                      repr_mode=repr_mode, threshold_fnz=0.35, device=lr)
        ms, ml = st.m_start, st.m_local
        st.load_data_from_bed(bed[ms:ms + ml])
        st.finalize()
        st.comm_init(dist)
        brr = hydra_b200.BayesRRm(st, y, mS, groups=groups, sigmaG0=sigmaG0, seed=seed)
        ahead = repeated = 0
        for it in range(n_iter):
            tp = dict(zmu=tape["zmu"][it][rank * TL:(rank + 1) * TL], perm=tape["perm"][it][ms:ms + ml], u=tape["u"][it][ms:ms + ml],
                      z=tape["z"][it][ms:ms + ml], sigmaG=ref["sigmaG"][it], pi=ref["pi"][it], sigmaE=ref["sigmaE"][it:it + 1])
            o = brr.iteration(tp)
            beta, comp, acum = brr.state()
            h = brr.hyper()
            assert np.array_equal(comp, ref["comp"][it][ms:ms + ml]), f"case {case} rank {rank}: components differ at iteration {it}"
            np.testing.assert_allclose(beta, ref["beta"][it][ms:ms + ml], rtol=1e-10, atol=1e-15)
            assert np.array_equal(h["cass"], ref["cass"][it])
            np.testing.assert_allclose(h["bsq"], ref["bsq"][it], rtol=1e-10)
            np.testing.assert_allclose(o["e_sqn"], ref["esqn"][it], rtol=1e-10)
            assert o["n_sync"] == ref["nsync"][it]
            ahead += o["windows_ahead"]; repeated += o["draws_repeated"]
            if case == 3 and it == n_iter - 1 and rank == 0:
                per_window = o["markers_changed"] / max(1, o["n_windows"])
                print(f"case 3: {per_window:.0f} changed markers per window over all GPUs", flush=True)
                assert per_window > 0.4 * 640 * world, per_window   # > 1024 from 4 GPUs on
            for t in range(TL):
                np.testing.assert_allclose(brr.task_epsilon(t), ref["eps"][it, rank * TL + t], rtol=1e-10, atol=1e-12)
        if case == 4:
            assert ahead > 20 and repeated > 0, (ahead, repeated)
        st.close()
        dist.barrier()
        if rank == 0:
            print(f"multi-GPU parity case {case} ok on {world} GPUs", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
