"""hydra_b200: B200-native (sm_100a) per-marker Gibbs hot path of hydra (BayesRRm / BayesW).

Public API: `GenotypeStore` (genotype staging + epsilon, the role of the reference's Data class)
and `BayesRRm` (the sampler), both thin ctypes front-ends of libhydra_b200.so (include/hydra_b200.h).
No CPU fallback: importing works anywhere, computing needs the CUDA library and a B200.
"""
from . import capi, synth
from .capi import HydraError
from .sampler import BayesRRm, BayesW, GenotypeStore

__all__ = ["GenotypeStore", "BayesRRm", "BayesW", "HydraError", "capi", "synth"]
