"""Quick performance probe of the BayesW marker loop (developer tool, not the bench)."""
import argparse
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import hydra_b200
from hydra_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=500000)
ap.add_argument("--m", type=int, default=32768)
ap.add_argument("--spectrum", default="B")
ap.add_argument("--tasks", type=int, default=64)
ap.add_argument("--sr", type=int, default=10)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--quad", type=int, default=25)
a = ap.parse_args()

st = hydra_b200.GenotypeStore(a.n, a.m, tasks=a.tasks, sync_rate=a.sr, n_groups=1, n_mix=4, repr_mode="sparse", model="bayesW")
synth.stage_synthetic(st, a.spectrum)
g, causal, beta = synth.simulate_phenotype(st, n_causal=max(10, a.m // 200))  # genetic values + noise; only the genetic part is used
rng = np.random.default_rng(3)
g = (g - g.mean()) / g.std() * np.sqrt(0.01675)
y = 4.1 + g + np.log(rng.exponential(size=a.n)) / 10.0 + 0.577215664901532 / 10.0  # log t = mu + g + w / alpha (example/Weibull.h2)
fail = (rng.random(a.n) > 0.1).astype(np.float64)                                  # 10 % censored
bw = hydra_b200.BayesW(st, y, fail, [[0.001, 0.01, 0.1]], quad_points=a.quad, seed=5)
print(f"# BayesW N={a.n} M={a.m} T={a.tasks} SR={a.sr} S={st.n_slices} quad_points={a.quad}", flush=True)
for it in range(a.iters):
    t0 = time.time()
    o = bw.iteration()
    w = time.time() - t0
    print(f"  it {it}: loop {o['loop_ms']:.2f} ms iter {o['iter_ms']:.2f} ms wall {w*1e3:.1f} ms  {a.m/o['loop_ms']/1e3:.3f} M markers/s  windows {o['n_windows']} "
          f"launches {o['n_launches']} changed {o['markers_changed']} us/window {o['loop_ms']*1e3/max(1,o['n_windows']):.1f}  mu {o['mu']:.4f} alpha {o['alpha']:.3f} "
          f"density evals {o['density_evals']}", flush=True)
st.close()
