"""Pins of the CPU oracle to the REFERENCE'S OWN OBJECT CODE.

oracle/build_ref.sh cuts the bodies of the reference's unit kernels and staging functions out of
/root/reference/src/BayesRRm.cpp (:60-413 helpers, sparse_dotprod, sparse_scaadd, center_and_scale, marker blocks; :1757-1849 and
:1976-2019, the LUT statements of the marker loop) and src/data.cpp (:826-865, :1112-1290) and compiles them, from where they
lie, behind stub class declarations into oracle/_ref/libref_kernels.so.  Here every restated function of oracle/hydra_oracle.c
is compared with that object code BIT FOR BIT (both strict IEEE, sums in source order).  What stays unpinned after this file:
the random streams (Boost.Random is absent) and the skeleton of the marker loop / mixture cascade (Eigen expressions)."""
import numpy as np
import pytest

import oracle
from helpers import random_bed

pytestmark = pytest.mark.skipif(oracle.ref_kernels() is None, reason="oracle/_ref/libref_kernels.so not built (needs /root/reference; run oracle/build_ref.sh)")


def _lists_equal(a, b):
    for f in ("I1", "N1S", "N1L", "I2", "N2S", "N2L", "IM", "NMS", "NML"):
        x, y = getattr(a, f), getattr(b, f)
        assert x.dtype == y.dtype and np.array_equal(x, y), f


@pytest.mark.parametrize("N,M", [(1, 3), (4, 5), (6, 1), (1003, 29), (4099, 7)])
def test_sparse_fill_indices_equals_reference_object_code(N, M):
    rng = np.random.default_rng(N * 31 + M)
    bed, _ = random_bed(rng, M, N, pmiss=0.07)
    _lists_equal(oracle.sparse_fill_indices(bed, N), oracle.ref_sparse_fill_indices(bed, N))


@pytest.mark.parametrize("n_na", [0, 1, 9, 200])
def test_na_compaction_equals_reference_object_code(n_na):
    rng = np.random.default_rng(n_na)
    N, M = 777, 21
    bed, _ = random_bed(rng, M, N, pmiss=0.05)
    na = np.sort(rng.choice(N, size=n_na, replace=False)).astype(np.uint32)
    a, b = oracle.sparse_fill_indices(bed, N), oracle.ref_sparse_fill_indices(bed, N)
    oracle.correct_for_missing_phenotype(a, na)
    oracle.ref_correct_for_missing_phenotype(b, na)
    for f in ("N1S", "N1L", "N2S", "N2L", "NMS", "NML"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    for I, S, L, J in ((a.I1, a.N1S, a.N1L, b.I1), (a.I2, a.N2S, a.N2L, b.I2), (a.IM, a.NMS, a.NML, b.IM)):
        for s, l in zip(S, L):  # holes are left behind the shrunk lists (:1155): only the live part is defined
            assert np.array_equal(I[int(s): int(s + l)], J[int(s): int(s + l)])


def test_bed_marker_from_sparse_equals_reference_object_code():
    rng = np.random.default_rng(3)
    for N in (5, 64, 1001):
        bed, _ = random_bed(rng, 6, N, pmiss=0.1)
        sp = oracle.sparse_fill_indices(bed, N)
        for m in range(6):
            lists = [I[int(S[m]): int(S[m] + L[m])] for I, S, L in ((sp.I1, sp.N1S, sp.N1L), (sp.I2, sp.N2S, sp.N2L), (sp.IM, sp.NMS, sp.NML))]
            assert np.array_equal(oracle.bed_marker_from_sparse(oracle.snp_len_byt(N), *lists), oracle.ref_bed_marker_from_sparse(N, *lists))


@pytest.mark.parametrize("N", [3, 257, 5000])
def test_dot_and_update_kernels_equal_reference_object_code(N):
    rng = np.random.default_rng(N)
    M = 12
    bed, _ = random_bed(rng, M, N, pmiss=0.03)
    sp = oracle.sparse_fill_indices(bed, N)
    mave, mstd = oracle.marker_stats_brr(N, sp.N1L, sp.N2L, sp.NML)
    eps = rng.normal(size=N)
    for m in range(M):
        # sparse_dotprod (:316-342) and the LUT dot of the marker loop (:1757-1840)
        assert oracle.sparse_dotprod(eps, sp, m, mave[m], mstd[m]) == oracle.ref_sparse_dotprod(eps, sp, m, mave[m], mstd[m])
        assert oracle.lut_dotprod(bed[m], eps, mave[m], mstd[m]) == oracle.ref_lut_dotprod(bed[m], eps, mave[m], mstd[m])
        # sparse_scaadd (:250-281) and the LUT deltaEps (:1976-2010); dMULT == 0 takes the other branch of :262
        for db in (0.0, -0.37):
            assert np.array_equal(oracle.sparse_scaadd(N, db, sp, m, mave[m], mstd[m]), oracle.ref_sparse_scaadd(N, db, sp, m, mave[m], mstd[m]))
        assert np.array_equal(oracle.lut_scaadd(N, bed[m], 0.61, mave[m], mstd[m]), oracle.ref_lut_scaadd(N, bed[m], 0.61, mave[m], mstd[m]))


def test_vector_helpers_blocks_and_scaling_equal_reference_object_code():
    import ctypes as C
    R = oracle.ref_kernels()
    rng = np.random.default_rng(0)
    y = rng.normal(3.0, 2.0, size=1237)
    a = y.copy()
    R.rk_center_and_scale(a.ctypes.data_as(C.c_void_p), C.c_int(len(a)))           # :371-388
    assert np.array_equal(a, oracle.center_and_scale(y))
    for Mtot, T in ((10, 3), (10000, 7), (5, 5), (1 << 20, 64)):                   # :396-413
        s, l = np.zeros(T, np.int32), np.zeros(T, np.int32)
        R.rk_define_blocks(C.c_int(Mtot), s.ctypes.data_as(C.c_void_p), l.ctypes.data_as(C.c_void_p), C.c_uint(T))
        os_, ol = oracle.define_blocks(Mtot, T)
        assert np.array_equal(s, os_) and np.array_equal(l, ol)
