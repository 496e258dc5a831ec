#!/usr/bin/env python
"""bench.py -- marker updates/sec and sec/iteration of the BayesRRm marker loop (BASELINE.json metric).

One "step" = one Gibbs iteration over all M markers of the synthetic workload:
N=500K individuals, spectrum-B sparse genotypes generated on the device, T hydra tasks x sync_rate.
At --gpus N the markers (and tasks) are partitioned over the N ranks (M = m_per_gpu * N: weak scaling).

Prints ONE JSON line on rank 0 (see the field list in DESIGN.md "Measurement").
  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
`--impl reference` times the CPU oracle (the reference loop restated, OpenMP "tasks as threads";
the reference itself needs MPI+Eigen+Boost and cannot be built in this image) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "marker_updates_per_sec"
UNIT = "marker updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="hydra_b200", choices=["hydra_b200", "reference"])
    ap.add_argument("--model", default="bayesrr", choices=["bayesrr", "bayesw"],
                    help="bayesrr = the headline workload (BASELINE config 4); bayesw = the same genotypes under the Weibull model "
                         "(BASELINE config 3 at scale; 1 GPU, 131072 markers unless --m-per-gpu is given)")
    ap.add_argument("--n", type=int, default=500_000, help="individuals")
    ap.add_argument("--m-per-gpu", type=int, default=None, help="markers per GPU (M = this x gpus); default 1 000 000 (bayesw: 131 072)")
    ap.add_argument("--spectrum", default="B")
    ap.add_argument("--tasks-per-gpu", type=int, default=64)
    ap.add_argument("--sync-rate", type=int, default=10)
    ap.add_argument("--cpu-sample-markers", type=int, default=49152)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--n-slices", type=int, default=0)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --m-per-gpu markers on every GPU (the driver's default); strong: --m-total markers shared by the GPUs "
                         "(BASELINE config 5: M = 8 000 000 over 1/2/4/8 GPUs, spectrum C so that the 1-GPU point is resident)")
    ap.add_argument("--m-total", type=int, default=8_000_000, help="total markers of a --scaling strong run")
    ap.add_argument("--no-extras", action="store_true", help="skip parity_check / latency_floor / bayesw sub-runs (profiling runs)")
    a = ap.parse_args()
    if a.m_per_gpu is None:
        a.m_per_gpu = 1_000_000 if a.model == "bayesrr" else 131_072
    if a.scaling == "strong":
        world = int(os.environ.get("WORLD_SIZE", str(a.gpus)))
        if a.m_total % world:
            raise SystemExit("--m-total must be a multiple of the number of GPUs")
        a.m_per_gpu = a.m_total // world
    return a


def workload_config(a, world):
    return {
        "workload": f"BayesRRm sparse, synthetic N={a.n} M={a.m_per_gpu * world} (spectrum {a.spectrum}: log-uniform MAF, "
                    f"0.1% missing), 1 group, S=0.0001,0.001,0.01, {a.tasks_per_gpu * world} tasks x sync_rate {a.sync_rate}",
        "n_individuals": a.n, "m_markers": a.m_per_gpu * world, "spectrum": a.spectrum,
        "tasks": a.tasks_per_gpu * world, "sync_rate": a.sync_rate,
        "partition": f"markers/tasks over {world} GPU(s)" + (f" (strong scaling: M fixed at {a.m_per_gpu * world})" if a.scaling == "strong" else ""),
        "l2_policy": "inputs larger than L2 (genotype records >> 126 MB, each read once per step)",
    }


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (NVML, 20 ms period; nvidia-smi as a fallback)."""
    BITS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, index=0):
        self.sm, self.reasons, self.max_mhz, self.stop_flag, self.index = [], set(), None, False, index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self.stop_flag:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception:
            q = "clocks.sm,clocks.max.sm"
            while not self.stop_flag:
                try:
                    o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                       capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.sm.append(float(o[0]))
                    self.max_mhz = float(o[1])
                except Exception:
                    pass
                time.sleep(0.1)

    def start(self):
        self.t.start()

    def stop(self):
        self.stop_flag = True
        self.t.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


# ----------------------------------------------------------------------------- CPU oracle leg
def cpu_sample_lists(a, n_markers, m_total):
    """Reference-format lists (I1/I2/IM) of the first n_markers markers of the workload, from the oracle's own CPU copy of
    the synthetic generator (bit-identical to the device generator: tests/test_gpu_parity.py)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    from hydra_b200 import synth  # host-side recipe only (numpy): MAF spectrum and thresholds
    p = synth.maf_spectrum(m_total, *synth.SPECTRA[a.spectrum])[:n_markers]
    thr = synth.thresholds(p)
    chunk = 256
    jobs = [(o, min(chunk, n_markers - o)) for o in range(0, n_markers, chunk)]

    def work(job):
        o, k = job
        bed = oracle.synth_bed(synth.SEED_GENO, a.n, thr[o:o + k], None, j0=o, fast=True)
        return oracle.sparse_fill_indices(bed, a.n)

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        parts = list(ex.map(work, jobs))
    I, S, Ln = [[], [], []], [[], [], []], [[], [], []]
    base = [0, 0, 0]
    for sp in parts:
        for w, (ii, ss, ll) in enumerate(((sp.I1, sp.N1S, sp.N1L), (sp.I2, sp.N2S, sp.N2L), (sp.IM, sp.NMS, sp.NML))):
            I[w].append(ii); S[w].append(ss + np.uint64(base[w])); Ln[w].append(ll)
            base[w] += len(ii)
    c = np.concatenate
    return oracle.SparseLists(c(I[0]), c(S[0]), c(Ln[0]), c(I[1]), c(S[1]), c(Ln[1]), c(I[2]), c(S[2]), c(Ln[2]))


def cpu_reference_run(a, n_ind, n_markers, m_total, steps, warmup):
    """Times the CPU restatement of the reference loop on the first n_markers markers (same spectrum, same N).
    Layout: T = host cores tasks x 1 thread each (the reference's '12 tasks x 1 thread' MPI layout as OpenMP threads)."""
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the CPU leg must use the host's cores regardless
    # (libgomp reads the variable when the oracle library is loaded, so it is set before the first import)
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(ncpu)
    import oracle
    oracle.set_num_threads(ncpu)
    cores = oracle.num_threads(True)
    t_prep = time.time()
    sp = cpu_sample_lists(a, n_markers, m_total)
    t_prep = time.time() - t_prep
    T = max(1, min(cores, n_markers))
    rng = np.random.default_rng(20240902)
    y = rng.normal(size=n_ind)
    mS = np.array([[0.0, 0.0001, 0.001, 0.01]])
    n_iter = steps + warmup
    tape = oracle.TapeMaker(1222, T, n_markers).make(n_iter)
    t0 = time.time()
    out = oracle.brr_chain(n_ind, n_markers, T, 4, 1, a.sync_rate, n_iter, sp, y, np.zeros(n_markers, np.int32), mS, tape,
                           np.array([0.5]), hyper_seed=7, want_eps=False, fast=True, want_marker_out=False)
    wall = time.time() - t0
    loop = out["loop_seconds"][warmup:]
    rate = n_markers * len(loop) / float(loop.sum())
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "tasks": T, "threads_per_task": 1,
            "sample": f"{n_markers} markers of the same workload (N={n_ind}), {T} tasks x 1 OpenMP thread, sync_rate {a.sync_rate}, "
                      f"{len(loop)} timed iteration(s) after {warmup} warm-up; marker loop only; chain {wall:.1f} s + data generation {t_prep:.1f} s",
            "ms_per_step_sample": float(loop.mean() * 1e3), "compile": "gcc -Ofast -march=native -fopenmp (reference src/Makefile_G flags)"}


# ----------------------------------------------------------------------------- BayesW line (not the headline)
def weibull_phenotype(n, g, seed=3):
    """log t = mu + g + w / alpha with mu = 4.1, alpha = 10, var(g) = 0.01675 (example/Weibull.h2), 10 % censored."""
    rng = np.random.default_rng(seed)
    g = (g - g.mean()) / g.std() * np.sqrt(0.01675)
    y = 4.1 + g + np.log(rng.exponential(size=n)) / 10.0 + 0.577215664901532 / 10.0
    fail = (rng.random(n) > 0.1).astype(np.float64)
    return y, fail


def bayesw_cpu_run(a, n_markers, m_total, n_iter=2):
    """The CPU restatement of the reference's BayesW loop (oracle/hydra_oracle_bw.c + the reference's own ARMS object code),
    one thread simulating 16 tasks, on the first n_markers markers of the workload."""
    import oracle
    sp = cpu_sample_lists(a, n_markers, m_total)
    rng = np.random.default_rng(11)
    y, fail = weibull_phenotype(a.n, rng.normal(size=a.n))
    T = 16
    tm = oracle.TapeMaker(5, T, n_markers).make(n_iter)
    tape = dict(perm=tm["perm"], p=tm["u"])
    t0 = time.time()
    oracle.bw_chain(a.n, n_markers, T, 4, 1, a.sync_rate, n_iter, 25, sp, y, fail, np.zeros(n_markers, np.int32),
                    np.array([[0.0, 0.001, 0.01, 0.1]]), tape, 5, hyper_seed=7)
    wall = time.time() - t0
    return {"value": n_markers * n_iter / wall, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n_markers} markers of the same workload (N={a.n}), {T} tasks simulated by one thread, sync_rate {a.sync_rate}, "
                      f"{n_iter} iterations incl. the mu / alpha ARMS steps; {wall:.1f} s",
            "compile": "gcc -O2 (oracle) + the reference's src/BayesW_arms.cpp object code"}


def bayesw_line(a, local_rank, with_cpu=True, dist=None):
    """One GPU, or (dist = torch.distributed, one process per GPU) weak scaling: m_per_gpu markers and tasks_per_gpu tasks per GPU,
    epsilon replicated, the window's epsilon changes summed with ncclAllReduce (src/BayesW.cpp:1799-1835)."""
    import torch
    import hydra_b200
    from hydra_b200 import synth
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    M = a.m_per_gpu * world
    store = hydra_b200.GenotypeStore(a.n, M, tasks=a.tasks_per_gpu * world, task_first=rank * a.tasks_per_gpu, tasks_local=a.tasks_per_gpu,
                                     sync_rate=a.sync_rate, n_groups=1, n_mix=4, repr_mode="sparse",
                                     device=local_rank, model="bayesW", n_slices=a.n_slices)
    t0 = time.time()
    synth.stage_synthetic(store, a.spectrum)
    stage_s = time.time() - t0
    if dist is not None:
        store.comm_init(dist)
    g, _, _ = synth.simulate_phenotype(store, n_causal=max(10, M // 200), dist=dist)
    y, fail = weibull_phenotype(a.n, g)
    bw = hydra_b200.BayesW(store, y, fail, [[0.001, 0.01, 0.1]], quad_points=25, seed=5)
    n1, n2, nm = store.marker_counts()
    nnz_total = int((n1.astype(np.int64) + n2 + nm).sum())

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(*v):
        t = torch.tensor(v, dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]
    for _ in range(a.warmup):
        bw.iteration()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sync_all()
    t_wall = time.perf_counter()
    outs = [bw.iteration() for _ in range(a.steps)]
    sync_all()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    loop_ms = float(sum(o["loop_ms"] for o in outs))
    t_e2e = time.perf_counter()
    for _ in range(a.steps):
        bw.iteration()
        bw.state()
        bw.hyper()
    sync_all()
    e2e_ms = (time.perf_counter() - t_e2e) * 1e3
    wall_ms, loop_ms, e2e_ms = max_over_ranks(wall_ms, loop_ms, e2e_ms)
    clocks = sampler.stop()
    # algorithmic bytes: 12 B per stored non-zero visited (index + vi), per synchronisation 16*N (epsilon read, vi written)
    alg = [12.0 * nnz_total + 16.0 * store.n_ind * o["n_sync"] for o in outs]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else None
    peak = float(peaks["hbm_gbs"]) if peaks else 6650.0
    achieved = float(np.sum(alg)) / (loop_ms * 1e-3) / 1e9
    cfg = workload_config(a, world)
    cfg["workload"] = (f"BayesW (Weibull, 25 quadrature points) sparse, synthetic N={a.n} M={M} (spectrum {a.spectrum}), 1 group, "
                       f"S=0.001,0.01,0.1, {a.tasks_per_gpu * world} tasks x sync_rate {a.sync_rate}, 10 % censored")
    cfg["m_markers"] = M
    res = {"metric": METRIC, "model": "bayesW", "value": M * a.steps / (wall_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": wall_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic (device-generated genotypes, simulated Weibull phenotype)", "config": cfg,
           "marker_loop": {"ms_per_step": loop_ms / a.steps, "marker_updates_per_sec": M * a.steps / (loop_ms * 1e-3),
                           "windows_per_step": outs[-1]["n_windows"], "us_per_window": loop_ms * 1e3 / sum(o["n_windows"] for o in outs),
                           "markers_changed_last_step": outs[-1]["markers_changed"]},
           "e2e": {"value": M * a.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(4 * store.m_local), "d2h_bytes_per_step": int(12 * store.m_local + 256),
                   "ms_per_step": e2e_ms / a.steps, "what": "BayesW.iteration() + state() + hyper() through the C ABI"},
           "gpu_launches": int(sum(o["n_launches"] for o in outs)),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                        "kernel": "k_bw_window + k_bw_update (one launch pair per synchronisation window)",
                        "algorithmic_bytes_per_launch": float(np.mean(alg)),
                        "note": "latency-bound: the window waits for the serial ARMS draws of its few changing markers (DESIGN.md 6)"},
           "clocks": clocks,
           "layout": {"slices": store.n_slices, "genotype_bytes_per_gpu": store.genotype_bytes, "mean_nnz_per_marker": nnz_total / store.m_local,
                      "stage_seconds": stage_s}}
    if world > 1:
        res["exchange"] = "per window: ncclAllReduce of the dense epsilon change (N + 2 doubles) over NVLink / NVSwitch"
    if with_cpu and not a.no_cpu_baseline and world == 1:
        try:
            res["cpu_baseline"] = bayesw_cpu_run(a, min(2048, M), M)
        except Exception as e:
            res["cpu_baseline"] = {"error": repr(e)}
    store.close()
    return res


def bayesw_main(a, local_rank, world=1):
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    res = bayesw_line(a, local_rank, dist=dist)
    if dist is None or dist.get_rank() == 0:
        print(json.dumps(res), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------- extras of the headline line
def parity_check(rank, world, local_rank, dist):
    """A small deterministic replay against the CPU oracle ON THE RANKS OF THIS RUN, before the timed region: world GPUs x 4
    tasks each must reproduce the oracle run with 4*world tasks (mixed BED/sparse records, sync_rate 5, 3 iterations).
    The oracle is the checker here, nothing of it is timed or shipped."""
    import hydra_b200
    import oracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import bed_from_lists, random_bed, reference_lists, simulate_y
    N, M, TL, SR, G, K, n_iter, seed = 3000, 1536, 4, 5, 2, 4, 3, 77
    T = TL * world
    rng = np.random.default_rng(seed)
    bed, g = random_bed(rng, M, N, pmiss=0.01)
    sp = reference_lists(bed, N)
    y = simulate_y(rng, g, n_causal=60)
    groups = (np.arange(M) % G).astype(np.int32)
    mS = np.tile(np.array([0.0, 0.001, 0.01, 0.1]), (G, 1))
    sigmaG0 = np.array([0.4, 0.6])
    tape = oracle.TapeMaker(seed, T, M).make(n_iter)
    fnz = (sp.N1L + sp.N2L + sp.NML).astype(np.float64) / N
    usebed = (fnz > 0.35).astype(np.uint8)
    ref = oracle.brr_chain(N, M, T, K, G, SR, n_iter, sp, y, groups, mS, tape, sigmaG0, usebed=usebed, bed=bed_from_lists(sp, N))
    st = hydra_b200.GenotypeStore(N, M, tasks=T, task_first=rank * TL, tasks_local=TL, sync_rate=SR, n_groups=G, n_mix=K,
                                  repr_mode="mixed", threshold_fnz=0.35, device=local_rank)
    ms, ml = st.m_start, st.m_local
    st.load_data_from_bed(bed[ms:ms + ml])
    st.finalize()
    if world > 1:
        st.comm_init(dist)
    brr = hydra_b200.BayesRRm(st, y, mS, groups=groups, sigmaG0=sigmaG0, seed=seed)
    comp_ok, nsync_ok, beta_err, eps_err, n_changed = True, True, 0.0, 0.0, 0
    for it in range(n_iter):
        tp = dict(zmu=tape["zmu"][it][rank * TL:(rank + 1) * TL], perm=tape["perm"][it][ms:ms + ml], u=tape["u"][it][ms:ms + ml],
                  z=tape["z"][it][ms:ms + ml], sigmaG=ref["sigmaG"][it], pi=ref["pi"][it], sigmaE=ref["sigmaE"][it:it + 1])
        o = brr.iteration(tp)
        beta, comp, _ = brr.state()
        rb = ref["beta"][it][ms:ms + ml]
        comp_ok &= bool(np.array_equal(comp, ref["comp"][it][ms:ms + ml])) and bool(np.array_equal(brr.hyper()["cass"], ref["cass"][it]))
        nsync_ok &= int(o["n_sync"]) == int(ref["nsync"][it])
        beta_err = max(beta_err, float(np.max(np.abs(beta - rb) / np.maximum(np.abs(rb), 1e-5))))   # |beta| of a non-zero effect >> 1e-5
        for t in range(TL):
            re_ = ref["eps"][it, rank * TL + t]
            eps_err = max(eps_err, float(np.max(np.abs(brr.task_epsilon(t) - re_) / np.maximum(np.abs(re_), 1e-2))))
        n_changed += int(o["markers_changed"])
    st.close()
    v = np.array([beta_err, eps_err, 0.0 if comp_ok else 1.0, 0.0 if nsync_ok else 1.0])
    if world > 1:
        import torch
        t = torch.from_numpy(v).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        v = t.cpu().numpy()
    return {"n_gpus": world, "what": f"oracle replay on the ranks of this run: N={N} M={M} mixed records, {T} tasks x sync_rate {SR}, {G} groups, {n_iter} iterations",
            "max_rel_err": float(max(v[0], v[1])), "beta_max_rel_err": float(v[0]), "eps_max_rel_err": float(v[1]),
            "components_equal": bool(v[2] == 0.0), "n_sync_equal": bool(v[3] == 0.0), "markers_changed": n_changed,
            "ok": bool(v[2] == 0.0 and v[3] == 0.0 and max(v[0], v[1]) < 1e-10)}


def latency_floor(a, local_rank):
    """Per-marker latency floor: 1 task, sync after every marker (the reference's 1-rank configuration, BASELINE config 1
    at N = 500 000): every window is ONE marker, so the time per window is the serial latency of a marker update."""
    import hydra_b200
    from hydra_b200 import synth
    M = 4096
    st = hydra_b200.GenotypeStore(a.n, M, tasks=1, sync_rate=1, n_groups=1, n_mix=4, repr_mode="sparse", device=local_rank)
    synth.stage_synthetic(st, a.spectrum)
    y, _, _ = synth.simulate_phenotype(st, n_causal=20)
    brr = hydra_b200.BayesRRm(st, y, [[0.0001, 0.001, 0.01]], seed=1222)
    for _ in range(2):
        brr.iteration()
    outs = [brr.iteration() for _ in range(3)]
    st.close()
    return {"us_per_marker": float(sum(o["loop_ms"] for o in outs)) * 1e3 / (3 * M),
            "config": f"N={a.n}, {M} markers of spectrum {a.spectrum}, 1 task x sync_rate 1 (one marker per synchronisation window)"}


def bayesfh_sub(a, local_rank):
    """bayesFHMPI (horseshoe-type local scales, src/BayesRRm.cpp:1125-1163 ...) on the headline layout: the same marker kernel with
    per-marker prior operands plus k_fh_prepare / k_fh_finish either side of it. A driver-visible number, not the headline."""
    import hydra_b200
    from hydra_b200 import synth
    M = 262_144
    st = hydra_b200.GenotypeStore(a.n, M, tasks=a.tasks_per_gpu, sync_rate=a.sync_rate, n_groups=1, n_mix=4, repr_mode="sparse", device=local_rank,
                                  n_slices=a.n_slices)
    synth.stage_synthetic(st, a.spectrum)
    y, _, _ = synth.simulate_phenotype(st, n_causal=1250)
    brr = hydra_b200.BayesRRm(st, y, [[0.0001, 0.001, 0.01]], seed=1222, fh={})
    for _ in range(3):
        brr.iteration()
    outs = [brr.iteration() for _ in range(4)]
    f = brr.fh_state()
    st.close()
    it_ms = float(sum(o["iter_ms"] for o in outs)) / len(outs)
    return {"value": M / (it_ms * 1e-3), "unit": UNIT, "ms_per_step": it_ms, "marker_loop_ms": float(sum(o["loop_ms"] for o in outs)) / len(outs),
            "us_per_window": float(sum(o["loop_ms"] for o in outs)) * 1e3 / sum(o["n_windows"] for o in outs),
            "gpu_launches": int(sum(o["n_launches"] for o in outs)), "tau": float(f["tau"]), "markers_changed_last_step": int(outs[-1]["markers_changed"]),
            "workload": f"bayesFHMPI sparse, synthetic N={a.n} M={M} (spectrum {a.spectrum}), {a.tasks_per_gpu} tasks x sync_rate {a.sync_rate}, "
                        "device time of whole iterations (iter_ms)"}


# ----------------------------------------------------------------------------- main
def main():
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep NCCL's version banner off stdout (one JSON line only)
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        a.gpus = world

    if a.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(a, world)
    if a.model == "bayesw":
        return bayesw_main(a, local_rank, world)

    import torch
    import hydra_b200
    from hydra_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; hydra_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- driver-visible parity proof at THIS rank count (VERDICT r1 #1c), before anything is timed
    pc = None
    if not a.no_extras:
        try:
            pc = parity_check(rank, world, local_rank, dist)
        except Exception as e:  # the bench line must still print; a failed check is reported, not hidden
            pc = {"n_gpus": world, "ok": False, "error": repr(e)}

    T_total = a.tasks_per_gpu * world
    M_total = a.m_per_gpu * world
    store = hydra_b200.GenotypeStore(a.n, M_total, tasks=T_total, task_first=rank * a.tasks_per_gpu, tasks_local=a.tasks_per_gpu,
                                     sync_rate=a.sync_rate, n_groups=1, n_mix=4, repr_mode="sparse", device=local_rank,
                                     n_slices=a.n_slices)
    t0 = time.time()
    synth.stage_synthetic(store, a.spectrum)
    stage_s = time.time() - t0
    if world > 1:
        store.comm_init(dist)  # exchange of the epsilon updates between the GPUs (DESIGN.md "Multi-GPU")
    # phenotype: every rank needs the same y (causal markers spread over the ranks, genetic values summed)
    y, _, _ = synth.simulate_phenotype(store, n_causal=5000, dist=dist)
    brr = hydra_b200.BayesRRm(store, y, [[0.0001, 0.001, 0.01]], seed=1222)
    n1, n2, nm = store.marker_counts()
    nnz = n1.astype(np.int64) + n2 + nm
    nnz_total = int(nnz.sum())

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident leg: K iterations, inputs in HBM
    for _ in range(a.warmup):
        brr.iteration()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t_wall = time.perf_counter()
    outs = [brr.iteration() for _ in range(a.steps)]
    ev1.record()
    sync_all()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    loop_ms = float(sum(o["loop_ms"] for o in outs))
    iter_ms = float(sum(o["iter_ms"] for o in outs))
    t = torch.tensor([wall_ms, loop_ms, iter_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall_ms, loop_ms, iter_ms = [float(x) for x in t.cpu()]

    # ---- end-to-end leg through the host API: per step H2D of the marker order (+hyper tables), D2H of beta/components/acum
    # the results land in pinned host buffers (what a writer of .bet/.cpn/.acu files would hand to the library)
    # (two sets: the read-back of step i runs on a copy stream while step i+1 computes, the way a thin-every-iteration writer
    # uses hb_brr_get_state_async; every step's results are complete in host memory before the buffers are reused and before
    # the clock stops)
    pinned = [(torch.empty(store.m_local, dtype=torch.float64, pin_memory=True).numpy(),
               torch.empty(store.m_local, dtype=torch.int32, pin_memory=True).numpy(),
               torch.empty(store.m_local, dtype=torch.float64, pin_memory=True).numpy()) for _ in range(2)]
    brr.state_async(pinned[1])                 # untimed: creates the copy stream and the device snapshot buffers
    brr.state_wait()
    sync_all()
    t_e2e = time.perf_counter()
    for i in range(a.steps):
        brr.iteration()
        brr.state_wait()                       # step i-1 is in host memory (its buffer set is free again at step i+1)
        brr.state_async(pinned[i & 1])
        brr.hyper()
    brr.state_wait()
    sync_all()
    e2e_ms = (time.perf_counter() - t_e2e) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.cpu()[0])
    clocks = sampler.stop() if sampler else None

    # ---- roofline of the dominant kernel (k_brr_iteration): algorithmic bytes of SURVEY.md 8(d)
    #      12 B per stored non-zero visited by a dot product + 16 B per non-zero of a changed marker + 24*N per synchronisation
    pad_ratio = nnz_total / max(1.0, float(np.mean([o["nnz_processed"] for o in outs])))
    alg_bytes = [12.0 * nnz_total + 16.0 * o["nnz_updated"] * pad_ratio + 24.0 * store.n_ind * o["n_sync"] for o in outs]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else None
    peak = float(peaks["hbm_gbs"]) if peaks else 6650.0
    achieved = float(np.sum(alg_bytes)) / (loop_ms * 1e-3) / 1e9  # per-rank bytes / max-over-ranks kernel time
    # measured DRAM traffic of one launch (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum): only valid for the
    # 1-GPU default workload it was captured on; null elsewhere
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and world == 1 and a.m_per_gpu == 1_000_000 and a.spectrum == "B" and a.n == 500_000:
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    # second and third roofline from the kernel's own counters: what the launch really moves.
    #   DRAM: genotype words are read once (8 B per 64-bit word of four u16 indices = 2 B per stored non-zero incl. padding),
    #         the window-ordered marker data (64 B) and slice directory entries (16 B x slices) once, beta/components/Acum
    #         written (20 B), epsilon read + written once; BED records: slices*slice_len/4 B each.
    #   SMEM: every gathered word costs one LDS.64 per index lane group; the LSU moves one 128-B wavefront per clock and SM,
    #         a 64-bit gather of a full warp needs >= 2.
    Q = store.lmax * a.tasks_per_gpu
    nnz_pad = float(np.mean([o["nnz_processed"] for o in outs]))
    bed_m = float(np.mean([o["bed_markers"] for o in outs]))
    dram_bytes = (2.0 * nnz_pad + Q * (64.0 + 16.0 * store.n_slices) + 20.0 * store.m_local + 16.0 * store.n_slices * store.slice_len
                  + bed_m * store.n_slices * store.slice_len / 4.0)
    upd_pad = float(np.mean([o["nnz_updated"] for o in outs])) * store.n_cta_groups   # every CTA group applies every change
    sm_clock_hz = 1e6 * (clocks["sm_mhz"] if clocks and clocks.get("sm_mhz") else 1965.0)
    n_sms = store.n_slices * store.n_cta_groups
    lsu_wavefronts = 2.0 * (nnz_pad + 2.0 * upd_pad) / 32.0       # ideal: 2 wavefronts per 32-lane 64-bit access (update = load + store)
    t_loop = loop_ms / a.steps * 1e-3
    roof_dram = {"bytes_per_launch": dram_bytes, "gbps": dram_bytes / t_loop / 1e9, "frac_of_hbm": dram_bytes / t_loop / 1e9 / peak,
                 "source": "kernel counters (2 B x stored non-zeros + window tables + state)"}
    roof_smem = {"wavefronts_per_launch": lsu_wavefronts, "frac_of_lsu_peak": lsu_wavefronts / (t_loop * sm_clock_hz * n_sms),
                 "source": "kernel counters, ideal 2 wavefronts per 64-bit warp gather; peak = 1 wavefront/clock/SM on the CTAs' SMs"}

    if rank == 0:
        res = {
            "metric": METRIC, "value": M_total * a.steps / (wall_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": wall_ms / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic (device-generated genotypes, simulated phenotype)", "config": workload_config(a, world),
            "sec_per_iteration": wall_ms / a.steps * 1e-3, "parity_check": pc,
            "marker_loop": {"ms_per_step": loop_ms / a.steps, "marker_updates_per_sec": M_total * a.steps / (loop_ms * 1e-3),
                            "windows_per_step": outs[-1]["n_windows"], "syncs_per_step": outs[-1]["n_sync"],
                            "markers_changed_last_step": outs[-1]["markers_changed"], "us_per_window": loop_ms * 1e3 / sum(o["n_windows"] for o in outs)},
            "e2e": {"value": M_total * a.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(4 * store.m_local + 8 * 4 * 4 + 1), "d2h_bytes_per_step": int(20 * store.m_local + 8 * (3 + 2 * store.n_slices) + 16 + 128),
                    "ms_per_step": e2e_ms / a.steps, "what": "BayesRRm.iteration() + state_async()/state_wait() + hyper() through the C ABI: marker order H2D, beta/components/acum D2H into pinned host memory every step (thin=1), the read-back of a step overlapping the next step's marker loop"},
            "gpu_launches": int(sum(o["n_launches"] for o in outs)),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "dram": roof_dram, "smem": roof_smem,
                         "real_bound": "smem-lsu" if roof_smem["frac_of_lsu_peak"] > roof_dram["frac_of_hbm"] else "hbm",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "kernel": "k_brr_iteration (one cooperative launch per step)",
                         "algorithmic_bytes_per_launch": float(np.mean(alg_bytes)),
                         "note": "`frac` uses SURVEY 8(d)'s algorithmic bytes (12/16 B per non-zero + 24*N per sync), a model of the REFERENCE's "
                                 "traffic; the kernel keeps epsilon in shared memory and indices in 16 bits, so its real DRAM traffic is "
                                 "`dram` (~2 B per stored non-zero) and its real ceiling is the shared-memory gather rate (`smem`) plus the "
                                 "per-window latency chain (marker_loop.us_per_window)"},
            "clocks": clocks,
            "layout": {"slices": store.n_slices, "slice_len": store.slice_len, "cta_groups": store.n_cta_groups,
                       "genotype_bytes_per_gpu": store.genotype_bytes, "mean_nnz_per_marker": nnz_total / store.m_local, "stage_seconds": stage_s},
        }
        store.close()
        if world == 1 and not a.no_extras:
            try:
                res["latency_floor"] = latency_floor(a, local_rank)
            except Exception as e:
                res["latency_floor"] = {"error": repr(e)}
            try:
                res["bayesfh"] = bayesfh_sub(a, local_rank)
            except Exception as e:
                res["bayesfh"] = {"error": repr(e)}
            try:  # BayesW (BASELINE config 3 at scale): a driver-visible number next to the headline
                import copy
                b = copy.copy(a)
                b.m_per_gpu, b.steps, b.warmup = 131_072, 4, 2
                bl = bayesw_line(b, local_rank, with_cpu=False)
                res["bayesw"] = {k: bl[k] for k in ("value", "unit", "ms_per_step", "marker_loop", "roofline", "gpu_launches")}
                res["bayesw"]["workload"] = bl["config"]["workload"]
            except Exception as e:
                res["bayesw"] = {"error": repr(e)}
        if not a.no_cpu_baseline and world == 1:
            try:
                res["cpu_baseline"] = cpu_reference_run(a, a.n, min(a.cpu_sample_markers, M_total), M_total, 1, 1)
            except Exception as e:  # the bench line must still print
                res["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(res), flush=True)
    store.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def reference_arm(a, world):
    """CPU leg as its own arm: same metric/config, bounded sample per step, all host threads."""
    if a.model == "bayesw":
        n_markers = min(2048, a.m_per_gpu)
        r = bayesw_cpu_run(a, n_markers, a.m_per_gpu, n_iter=max(2, min(a.steps, 4)))
        cfg = workload_config(a, 1)
        cfg["workload"] = f"BayesW (Weibull, 25 quadrature points) sparse, synthetic N={a.n} M={a.m_per_gpu} (spectrum {a.spectrum}), 64 tasks x sync_rate {a.sync_rate}"
        cfg["m_markers"] = a.m_per_gpu
        print(json.dumps({"impl": "reference", "metric": METRIC, "model": "bayesW", "value": r["value"], "unit": UNIT, "n_gpus": 1, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": a.m_per_gpu / r["value"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic", "config": cfg, "cpu_baseline": r,
                          "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "note": "ms_per_step extrapolated linearly from the sample; CPU restatement of the reference's BayesW loop with the reference's own ARMS object code"}), flush=True)
        return 0
    n_markers = min(a.cpu_sample_markers, a.m_per_gpu * world)
    r = cpu_reference_run(a, a.n, n_markers, a.m_per_gpu * world, a.steps, a.warmup)
    res = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": r["ms_per_step_sample"] * (a.m_per_gpu * world / n_markers), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(a, world),
           "cpu_baseline": r, "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "reference_layout": {"tasks_run": r["tasks"], "threads_per_task": 1, "host_cores": r["cores"],
                                "why": "the CPU leg runs one hydra task per host core (the reference's '12 tasks x 1 thread per node' MPI layout); "
                                       "config.tasks is the GPU arm's task count"},
           "note": "ms_per_step extrapolated linearly from the sample to the full M; the reference binary needs MPI+Eigen+Boost (absent), "
                   "so this is the CPU restatement of its loop (oracle/hydra_oracle.c)"}
    print(json.dumps(res), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
