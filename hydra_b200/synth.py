"""Synthetic genotype / phenotype recipe of SURVEY.md 8(d) (bench + tests scaffold).

Genotype cells come from a counter-based generator evaluated ON THE DEVICE (hb_stage_synth): cell (i,j) is
word i%4 of Philox4x32-10(ctr=(i/4, j, attempt, 'GENO'), key=(seed,0)) compared with three integer
thresholds per marker, so the same matrix can be regenerated bit-for-bit anywhere (the CPU oracle has the
same generator for parity tests).
"""
from __future__ import annotations

import numpy as np

SEED_GENO = 20240901
SEED_PHEN = 20240902
SPECTRA = {  # log-uniform MAF ranges
    "A": (2e-4, 0.5),    # "ukb-like"
    "B": (2e-4, 0.05),   # rare
    "C": (2e-4, 0.02),   # very rare
    "E": (0.01, 0.5),    # example-scale common variants
}


def maf_spectrum(m_total: int, lo: float, hi: float, seed: int = SEED_GENO) -> np.ndarray:
    rng = np.random.Generator(np.random.Philox(key=seed))
    return np.exp(rng.uniform(np.log(lo), np.log(hi), size=m_total))


def thresholds(p: np.ndarray, pmiss: float = 0.001) -> np.ndarray:
    """(M,3) uint32: w < t0 missing, < t1 two copies, < t2 one copy, else zero copies."""
    p = np.asarray(p, dtype=np.float64)
    tm = np.floor(pmiss * 4294967296.0)
    rest = 4294967296.0 - tm
    t2 = tm + np.floor(p * p * rest)
    t1 = t2 + np.floor(2.0 * p * (1.0 - p) * rest)
    return np.stack([np.full_like(p, tm), t2, t1], axis=1).astype(np.uint32)


def stage_synthetic(store, spectrum="B", pmiss=0.001, seed=SEED_GENO, chunk=1 << 16):
    """Stage the local markers of `store` from the synthetic recipe; monomorphic markers are re-drawn."""
    lo, hi = SPECTRA[spectrum] if isinstance(spectrum, str) else spectrum
    p = maf_spectrum(store.m_total, lo, hi, seed)[store.m_start: store.m_start + store.m_local]
    thr = thresholds(p, pmiss)
    for o in range(0, store.m_local, chunk):
        store.load_synthetic(seed, thr[o: o + chunk], None, m_first=o)
    attempts = np.zeros(store.m_local, np.uint32)
    for _ in range(8):
        n1, n2, nm = store.marker_counts()
        bad = np.flatnonzero((n1.astype(np.int64) + n2) == 0)
        if len(bad) == 0:
            break
        attempts[bad] += 1
        for m in bad:
            store.load_synthetic(seed, thr[m: m + 1], attempts[m: m + 1], m_first=int(m))
    store.finalize()
    return p, thr, attempts


def genetic_values(store, n_causal, h2, seed):
    """g = X beta for n_causal local markers with beta ~ N(0, h2/n_causal_total-ish), on standardised columns, computed
    on the device with the sampler's own update kernel."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    n_causal = min(n_causal, store.m_local)
    causal = np.sort(rng.choice(store.m_local, size=n_causal, replace=False)).astype(np.uint32)
    beta = rng.normal(0.0, 1.0, size=n_causal)
    store.set_epsilon(np.zeros(store.n_ind))
    store.sparse_scaadd(causal, beta)
    return store.get_epsilon(), causal, beta


def simulate_phenotype(store, n_causal=5000, h2=0.5, seed=SEED_PHEN, dist=None):
    """y = X beta + e with var(X beta) ~ h2. With `dist` (torch.distributed) the causal markers are spread over the ranks
    and the genetic values are summed, so that every rank ends up with the same phenotype."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    g, causal, beta = genetic_values(store, max(1, n_causal // world), h2, seed + 1 + rank)
    if dist is not None:
        import torch
        t = torch.from_numpy(g)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t)
        g = t.cpu().numpy()
    g = g * np.sqrt(h2 / max(g.var(), 1e-300))
    e = np.random.Generator(np.random.Philox(key=seed)).normal(0.0, np.sqrt(1.0 - h2), size=store.n_ind)
    return g + e, causal, beta
