import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import hydra_b200
from hydra_b200 import synth
N, M = 500000, 1000000
st = hydra_b200.GenotypeStore(N, M, tasks=64, sync_rate=10, n_groups=1, n_mix=4, repr_mode="sparse")
synth.stage_synthetic(st, "B")
y, causal, beta = synth.simulate_phenotype(st, n_causal=5000)
brr = hydra_b200.BayesRRm(st, y, [[0.0001, 0.001, 0.01]], seed=1222)
for _ in range(6): brr.iteration()
pin = [(torch.empty(M, dtype=torch.float64, pin_memory=True).numpy(), torch.empty(M, dtype=torch.int32, pin_memory=True).numpy(), torch.empty(M, dtype=torch.float64, pin_memory=True).numpy()) for _ in range(2)]
def timed(name, fn, n=6):
    torch.cuda.synchronize(); t = time.perf_counter()
    tt = {}
    for i in range(n): fn(i, tt)
    brr.state_wait(); torch.cuda.synchronize()
    print(name, round((time.perf_counter() - t) * 1e3 / n, 3), "ms/step", {k: round(v * 1e3 / n, 3) for k, v in tt.items()}, flush=True)
def f_iter(i, tt):
    t = time.perf_counter(); o = brr.iteration(); tt["iter"] = tt.get("iter", 0) + time.perf_counter() - t; tt["dev_iter_ms"] = tt.get("dev_iter_ms", 0) + o["iter_ms"] * 1e-3; tt["loop_ms"] = tt.get("loop_ms", 0) + o["loop_ms"] * 1e-3
def f_async(i, tt):
    f_iter(i, tt)
    t = time.perf_counter(); brr.state_wait(); tt["wait"] = tt.get("wait", 0) + time.perf_counter() - t
    t = time.perf_counter(); brr.state_async(pin[i & 1]); tt["async"] = tt.get("async", 0) + time.perf_counter() - t
def f_async_h(i, tt):
    f_async(i, tt)
    t = time.perf_counter(); brr.hyper(); tt["hyper"] = tt.get("hyper", 0) + time.perf_counter() - t
def f_sync(i, tt):
    f_iter(i, tt)
    t = time.perf_counter(); brr.state(out=pin[0]); tt["state"] = tt.get("state", 0) + time.perf_counter() - t
timed("iteration only      ", f_iter)
timed("+ state_async       ", f_async)
timed("+ state_async+hyper ", f_async_h)
timed("+ state (sync)      ", f_sync)
timed("iteration only      ", f_iter)
