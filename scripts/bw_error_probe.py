"""Developer tool: measured differences between the BayesW CUDA path and the CPU oracle (+ the reference's ARMS object code),
per quantity -- the numbers behind the tolerances asserted in tests/test_gpu_bayesw.py ("tolerances")."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hydra_b200  # noqa: E402
import oracle  # noqa: E402
from hydra_b200 import sampler  # noqa: E402
from helpers import random_bed, reference_lists  # noqa: E402
from test_gpu_bayesw import _weibull_data  # noqa: E402


def rel(a, b, floor=0.0):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def main():
    out = {}
    # marginal likelihoods
    worst = 0.0
    with hydra_b200.GenotypeStore(64, 4, n_groups=1, n_mix=4, model="bayesW") as st:
        for quad in (3, 7, 11, 25):
            rng = np.random.default_rng(quad)
            for _ in range(50):
                p = 0.01 + 0.49 * rng.random()
                mean, sd = 2 * p, np.sqrt(2 * p * (1 - p))
                Nn = 5000
                v1, v2, v0 = Nn * 2 * p * (1 - p) * rng.uniform(0.8, 1.2), Nn * p * p * rng.uniform(0.8, 1.2), Nn * (1 - p) ** 2 * rng.uniform(0.8, 1.2)
                pars = [rng.uniform(5, 12), rng.uniform(0.005, 0.03), rng.normal() * 30, v0 + v1 + v2, v0, v1, v2, mean, sd, mean / sd]
                prior, cVa = [0.9, 0.05, 0.03, 0.02], [0.001, 0.01, 0.1]
                worst = max(worst, rel(sampler.bw_marginal_likelihoods(st, quad, pars, prior, cVa), oracle.bw_marginal_likelihoods(quad, pars, prior, cVa), 1e-300))
        out["marginal_likelihood_max_rel"] = worst
        # ARMS
        rng = np.random.default_rng(5)
        same, worst, n = 0, 0.0, 400
        diffs = []
        for t in range(n):
            p = 0.01 + 0.49 * rng.random()
            mean, sd = 2 * p, np.sqrt(2 * p * (1 - p))
            Nn = 5000
            v1, v2, v0 = Nn * 2 * p * (1 - p), Nn * p * p, Nn * (1 - p) ** 2
            pars = [rng.uniform(5, 12), rng.uniform(0.005, 0.03), rng.normal() * 30, v0 + v1 + v2, v0, v1, v2, mean, sd, mean / sd]
            Ck, ssg, bold = float(rng.choice([0.001, 0.01, 0.1])), 0.02, float(rng.choice([0.0, 0.01, -0.02]))
            got = sampler.bw_arms_beta(st, pars, Ck, ssg, bold, 77, 3, 5, t)
            want = oracle.bw_sample_beta(pars, Ck, ssg, bold, 77, 3, 5, t)
            if got["nrand"] == want["nrand"] and got["neval"] == want["neval"]:
                same += 1
                worst = max(worst, abs(got["beta"] - want["beta"]) / max(abs(want["beta"]), 1e-15))
            else:
                diffs.append((t, got["nrand"], want["nrand"], got["neval"], want["neval"], got["beta"], want["beta"]))
        out["arms_same_control_flow"] = f"{same}/{n}"
        out["arms_beta_max_rel_when_same"] = worst
        out["arms_diverging_cases"] = diffs[:5]
    # chain
    for (T, SR, G, repr_mode, replay_hyper) in [(1, 1, 1, "sparse", True), (3, 4, 2, "sparse", True), (2, 3, 1, "bed", False)]:
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
        e = dict(mu=0.0, alpha=0.0, beta=0.0, eps=0.0, sigmaG=0.0, pi=0.0)
        with hydra_b200.GenotypeStore(N, M, tasks=T, sync_rate=SR, n_groups=G, n_mix=K, repr_mode=repr_mode, model="bayesW") as st:
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
                e["mu"] = max(e["mu"], rel(o["mu"], ref["mu"][it]))
                e["alpha"] = max(e["alpha"], rel(o["alpha"], ref["alpha"][it]))
                nz = ref["beta"][it] != 0
                e["beta"] = max(e["beta"], rel(beta[nz], ref["beta"][it][nz]))
                e["eps"] = max(e["eps"], rel(bw.epsilon(), ref["eps"][it], 1e-3))
                e["sigmaG"] = max(e["sigmaG"], rel(h["sigmaG"], ref["sigmaG"][it]))
                e["pi"] = max(e["pi"], rel(h["pi"], ref["pi"][it]))
        out[f"chain T={T} SR={SR} G={G} {repr_mode}"] = e
    for k, v in out.items():
        print(k, v)


if __name__ == "__main__":
    main()
