"""Shared test scaffolding: small random genotype sets in the reference's formats."""
import numpy as np

import oracle


def random_bed(rng, M, N, maf_lo=0.02, maf_hi=0.5, pmiss=0.01):
    """(M, ceil(N/4)) uint8 PLINK columns (pad bits 00) + the dense genotype matrix (M,N) with -1 = missing."""
    p = rng.uniform(maf_lo, maf_hi, size=M)
    g = rng.binomial(2, p[:, None], size=(M, N)).astype(np.int8)
    g[rng.random((M, N)) < pmiss] = -1
    for j in range(M):  # no monomorphic markers
        if (g[j] > 0).sum() == 0:
            g[j, rng.integers(N)] = 1
    code = np.select([g == 2, g == 1, g == 0], [0, 2, 3], default=1).astype(np.uint8)
    nb = (N + 3) // 4
    pad = np.zeros((M, nb * 4), np.uint8)
    pad[:, :N] = code
    pad = pad.reshape(M, nb, 4)
    bed = (pad[:, :, 0] | (pad[:, :, 1] << 2) | (pad[:, :, 2] << 4) | (pad[:, :, 3] << 6)).astype(np.uint8)
    return bed, g


def reference_lists(bed, Nraw, na_inds=None):
    """Reference representation: sparse_data_fill_indices on the raw columns, then NA compaction."""
    sp = oracle.sparse_fill_indices(bed, Nraw)
    if na_inds is not None and len(na_inds):
        oracle.correct_for_missing_phenotype(sp, np.asarray(na_inds, np.uint32))
    return sp


def compact_lists(sp):
    """Lists with the NA holes squeezed out (starts recomputed) -- for comparisons with exports."""
    out = []
    for I, S, L in ((sp.I1, sp.N1S, sp.N1L), (sp.I2, sp.N2S, sp.N2L), (sp.IM, sp.NMS, sp.NML)):
        parts = [I[int(s): int(s) + int(l)] for s, l in zip(S, L)]
        out.append((np.concatenate(parts) if parts else np.zeros(0, np.uint32), np.concatenate([[0], np.cumsum(L)[:-1]]).astype(np.uint64), L.copy()))
    return out


def bed_from_lists(sp, N):
    """NA-compacted BED bytes per marker via the oracle's get_bed_marker_from_sparse."""
    M = len(sp.N1S)
    nb = (N + 3) // 4
    out = np.zeros((M, nb), np.uint8)
    for m in range(M):
        out[m] = oracle.bed_marker_from_sparse(nb, sp.I1[int(sp.N1S[m]): int(sp.N1S[m] + sp.N1L[m])],
                                               sp.I2[int(sp.N2S[m]): int(sp.N2S[m] + sp.N2L[m])],
                                               sp.IM[int(sp.NMS[m]): int(sp.NMS[m] + sp.NML[m])])
        if N % 4:  # the reference leaves the pad bits of the last byte at 11; PLINK files carry 00
            out[m, -1] &= (1 << (2 * (N % 4))) - 1
    return out


def simulate_y(rng, g, n_causal=20, h2=0.5):
    M, N = g.shape
    x = np.where(g < 0, 0, g).astype(np.float64)
    x = (x - x.mean(1, keepdims=True)) / (x.std(1, keepdims=True) + 1e-12)
    causal = rng.choice(M, size=min(n_causal, M), replace=False)
    b = rng.normal(0, np.sqrt(h2 / len(causal)), size=len(causal))
    return x[causal].T @ b + rng.normal(0, np.sqrt(1 - h2), size=N) + 3.0
