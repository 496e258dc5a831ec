"""Quick performance probe of the marker loop (developer tool, not the bench)."""
import argparse
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import hydra_b200
from hydra_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=500000)
ap.add_argument("--m", type=int, default=65536)
ap.add_argument("--spectrum", default="B")
ap.add_argument("--tasks", type=int, nargs="+", default=[64])
ap.add_argument("--sr", type=int, nargs="+", default=[10])
ap.add_argument("--slices", type=int, nargs="+", default=[0])
ap.add_argument("--repr", default="sparse")
ap.add_argument("--iters", type=int, default=4)
a = ap.parse_args()

for ns in a.slices:
    for T in a.tasks:
        for SR in a.sr:
            t0 = time.time()
            st = hydra_b200.GenotypeStore(a.n, a.m, tasks=T, sync_rate=SR, n_groups=1, n_mix=4, repr_mode=a.repr, n_slices=ns)
            synth.stage_synthetic(st, a.spectrum)
            t1 = time.time()
            y, causal, beta = synth.simulate_phenotype(st, n_causal=max(10, a.m // 200))
            brr = hydra_b200.BayesRRm(st, y, [[0.0001, 0.001, 0.01]], seed=1222)
            n1, n2, nm = st.marker_counts()
            nnz = float((n1.astype(np.int64) + n2 + nm).mean())
            print(f"# N={a.n} M={a.m} T={T} SR={SR} S={st.n_slices} L={st.slice_len} R={st.n_cta_groups} nnz/marker={nnz:.0f} "
                  f"geno={st.genotype_bytes/1e9:.2f} GB stage={t1-t0:.1f}s", flush=True)
            for it in range(a.iters):
                tw = time.time()
                o = brr.iteration()
                tw = time.time() - tw
                print(f"  it {it}: loop {o['loop_ms']:.3f} ms iter {o['iter_ms']:.3f} ms wall {tw*1e3:.1f} ms  {a.m/o['loop_ms']/1e3:.2f} M markers/s  "
                      f"windows {o['n_windows']} syncs {o['n_sync']} changed {o['markers_changed']} us/window {o['loop_ms']*1e3/max(1,o['n_windows']):.2f} "
                      f"sigmaE {o['sigmaE']:.4f} sigmaG {brr.hyper()['sigmaG'][0]:.4f}\n      cyc/window: " + " ".join(f"{n}={v/max(1,o['n_windows']):.0f}" for n, v in zip(["tab","dot","pub","bar","upd0","sum","updL","updA"], o["phase_cycles"])), flush=True)
            st.close()
