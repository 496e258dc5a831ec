// Test infrastructure (oracle/): C entry points around function bodies of the REFERENCE that oracle/build_ref.sh cuts out of
// /root/reference/src/{BayesRRm.cpp,data.cpp} by line range at build time (into oracle/_ref/*.inc, git-ignored, never
// committed).  This file holds no reference code: only the stub class declarations the bodies need (the real headers pull
// in MPI / Eigen / Boost, none installed) and extern "C" wrappers.  tests/test_oracle.py compares oracle/hydra_oracle.c with
// the resulting object code bit for bit.
//
//   brr_kernels.inc   src/BayesRRm.cpp:60-388    offset/set/copy/sum vector helpers, sparse_set/add, sparse_scaadd,
//                                                partial_sparse_dotprod, sparse_dotprod, center_and_scale
//   brr_blocks.inc    src/BayesRRm.cpp:396-413   mpi_define_blocks_of_markers
//   brr_num.inc       src/BayesRRm.cpp:1757-1849 the `if (USEBED[marker]) {LUT dot} else {sparse_dotprod}` statement of the loop
//   brr_deps.inc      src/BayesRRm.cpp:1976-2019 the `if (USEBED[marker]) {LUT deltaEps} else {sparse_scaadd}` statement
//   data_bed.inc      src/data.cpp:826-865       Data::get_bed_marker_from_sparse
//   data_sparse.inc   src/data.cpp:1112-1290     NA-phenotype compaction, sizes from raw, sparse_data_fill_indices
#include <assert.h>
#include <immintrin.h>
#include <math.h>
#include <mm_malloc.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <iostream>
#include <vector>
using namespace std;
typedef unsigned int uint;
#define USE_MPI 1

static inline void check_malloc(const void *ptr, const int linenumber, const char *filename) {  // src/mpi_utils.hpp:19-25 without MPI_Abort
    if (ptr == NULL) { fprintf(stderr, "#FATAL#: malloc failed on line %d of %s\n", linenumber, filename); abort(); }
}

class BayesRRm {  // the members whose bodies lie in src/BayesRRm.cpp:60-413 (declared in src/BayesRRm.h:116-150)
public:
    void offset_vector_f64(double *__restrict__ vec, const double offset, const int N);
    void set_vector_f64(double *__restrict__ vec, const double val, const int N);
    void copy_vector_f64(double *__restrict__ dest, const double *__restrict__ source, const int N);
    double sum_vector_elements_f64_base(const double *__restrict__ vec, const int N);
    double sum_vector_elements_f64(const double *__restrict__ vec, const int N);
    void sum_vectors_f64(double *__restrict__ out, const double *__restrict__ in1, const double *__restrict__ in2, const int N);
    void sum_vectors_f64(double *__restrict__ out, const double *__restrict__ in1, const int N);
    void sparse_add(double *__restrict__ vec, const double val, const uint *__restrict__ IX, const size_t NXS, const size_t NXL);
    void sparse_scaadd(double *__restrict__ vout, const double dMULT, const uint *__restrict__ I1, const size_t N1S, const size_t N1L,
                       const uint *__restrict__ I2, const size_t N2S, const size_t N2L, const uint *__restrict__ IM, const size_t NMS,
                       const size_t NML, const double mu, const double sig_inv, const int N);
    double sparse_dotprod(const double *__restrict__ vin1, const uint *__restrict__ I1, const size_t N1S, const size_t N1L,
                          const uint *__restrict__ I2, const size_t N2S, const size_t N2L, const uint *__restrict__ IM, const size_t NMS,
                          const size_t NML, const double mu, const double sig_inv, const int N, const int marker);
    void mpi_define_blocks_of_markers(const int Mtot, int *MrankS, int *MrankL, const uint nblocks);
    // wrappers (this file) around the two statements cut out of the marker loop
    double loop_num(const bool *USEBED, int marker, uint *I1, const size_t *N1S, const size_t *N1L, uint *I2, const size_t *N2S,
                    const size_t *N2L, uint *IM, const size_t *NMS, const size_t *NML, double *epsilon, int Ntot, size_t snpLenByt,
                    const double *mave, const double *mstd);
    void loop_deltaeps(const bool *USEBED, int marker, uint *I1, const size_t *N1S, const size_t *N1L, uint *I2, const size_t *N2S,
                       const size_t *N2L, uint *IM, const size_t *NMS, const size_t *NML, double *deltaEps, double deltaBeta, int Ntot,
                       const double *mave, const double *mstd);
};

#include "dotp_lut.h"          // -I /root/reference/src
#include "_ref/brr_kernels.inc"
#include "_ref/brr_blocks.inc"

double BayesRRm::loop_num(const bool *USEBED, int marker, uint *I1, const size_t *N1S, const size_t *N1L, uint *I2, const size_t *N2S,
                          const size_t *N2L, uint *IM, const size_t *NMS, const size_t *NML, double *epsilon, int Ntot, size_t snpLenByt,
                          const double *mave, const double *mstd) {
    double num = 0.0;
    (void)snpLenByt;
#include "_ref/brr_num.inc"
    return num;
}

void BayesRRm::loop_deltaeps(const bool *USEBED, int marker, uint *I1, const size_t *N1S, const size_t *N1L, uint *I2, const size_t *N2S,
                             const size_t *N2L, uint *IM, const size_t *NMS, const size_t *NML, double *deltaEps, double deltaBeta, int Ntot,
                             const double *mave, const double *mstd) {
#include "_ref/brr_deps.inc"
}

class Data {  // the members used by the bodies of src/data.cpp:826-865, 1112-1290 (src/data.hpp:127-130)
public:
    unsigned numInds = 0;
    unsigned numNAs = 0;
    vector<uint> NAsInds;
    void get_bed_marker_from_sparse(char *bdat, const int Ntot, const size_t S1, const size_t L1, const uint *I1, const size_t S2,
                                    const size_t L2, const uint *I2, const size_t SM, const size_t LM, const uint *IM);
    void sparse_data_correct_for_missing_phenotype(const size_t *NS, size_t *NL, uint *I, const int M, const bool *USEBED);
    void sparse_data_get_sizes_from_raw(const char *rawdata, const uint NC, const uint NB, const uint NA, size_t &N1, size_t &N2, size_t &NM);
    void sparse_data_fill_indices(const char *rawdata, const uint NC, const uint NB, const uint NA, size_t *N1S, size_t *N1L, uint *I1,
                                  size_t *N2S, size_t *N2L, uint *I2, size_t *NMS, size_t *NML, uint *IM);
};
#include "_ref/data_bed.inc"
#include "_ref/data_sparse.inc"

extern "C" {
#define RK __attribute__((visibility("default")))
RK double rk_sum_vector_elements_f64(const double *v, int N) { BayesRRm b; return b.sum_vector_elements_f64(v, N); }
RK void rk_sum_vectors_f64(double *out, const double *in1, int N) { BayesRRm b; b.sum_vectors_f64(out, in1, N); }
RK void rk_sum_vectors3_f64(double *out, const double *in1, const double *in2, int N) { BayesRRm b; b.sum_vectors_f64(out, in1, in2, N); }
RK void rk_offset_vector_f64(double *v, double off, int N) { BayesRRm b; b.offset_vector_f64(v, off, N); }
RK void rk_center_and_scale(double *v, int N) { center_and_scale(v, N); }
RK void rk_define_blocks(int Mtot, int *S, int *L, unsigned nblocks) { BayesRRm b; b.mpi_define_blocks_of_markers(Mtot, S, L, nblocks); }
RK double rk_sparse_dotprod(const double *eps, const uint *I1, size_t N1S, size_t N1L, const uint *I2, size_t N2S, size_t N2L, const uint *IM,
                            size_t NMS, size_t NML, double mu, double sig_inv, int N) {
    BayesRRm b;
    return b.sparse_dotprod(eps, I1, N1S, N1L, I2, N2S, N2L, IM, NMS, NML, mu, sig_inv, N, 0);
}
RK void rk_sparse_scaadd(double *out, double dMULT, const uint *I1, size_t N1S, size_t N1L, const uint *I2, size_t N2S, size_t N2L,
                         const uint *IM, size_t NMS, size_t NML, double mu, double sig_inv, int N) {
    BayesRRm b;
    b.sparse_scaadd(out, dMULT, I1, N1S, N1L, I2, N2S, N2L, IM, NMS, NML, mu, sig_inv, N);
}
// marker loop statements: one marker stored as BED bytes inside the I1 array (USEBED = 1; src/data.cpp:1043)
RK double rk_loop_num_bed(const uint8_t *raw, double *eps, int Ntot, double mave, double mstd) {
    BayesRRm b;
    const size_t nb = (size_t)(Ntot + 3) / 4;
    vector<uint> I1((nb + 3) / 4 + 1, 0u);
    memcpy(I1.data(), raw, nb);
    const bool ub = true;
    const size_t z = 0;
    return b.loop_num(&ub, 0, I1.data(), &z, &z, nullptr, &z, &z, nullptr, &z, &z, eps, Ntot, nb, &mave, &mstd);
}
RK void rk_loop_deltaeps_bed(const uint8_t *raw, double *deltaEps, double deltaBeta, int Ntot, double mave, double mstd) {
    BayesRRm b;
    const size_t nb = (size_t)(Ntot + 3) / 4;
    vector<uint> I1((nb + 3) / 4 + 1, 0u);
    memcpy(I1.data(), raw, nb);
    const bool ub = true;
    const size_t z = 0;
    b.loop_deltaeps(&ub, 0, I1.data(), &z, &z, nullptr, &z, &z, nullptr, &z, &z, deltaEps, deltaBeta, Ntot, &mave, &mstd);
}
RK void rk_get_bed_marker_from_sparse(char *bdat, int Ntot, const uint *I1, size_t L1, const uint *I2, size_t L2, const uint *IM, size_t LM) {
    Data d;
    d.get_bed_marker_from_sparse(bdat, Ntot, 0, L1, I1, 0, L2, I2, 0, LM, IM);
}
RK void rk_sparse_get_sizes_from_raw(const char *raw, unsigned NC, unsigned NB, unsigned numInds, size_t *N1, size_t *N2, size_t *NM) {
    Data d;
    d.numInds = numInds;
    d.sparse_data_get_sizes_from_raw(raw, NC, NB, 0, *N1, *N2, *NM);
}
RK void rk_sparse_fill_indices(const char *raw, unsigned NC, unsigned NB, unsigned numInds, size_t *N1S, size_t *N1L, uint *I1, size_t *N2S,
                               size_t *N2L, uint *I2, size_t *NMS, size_t *NML, uint *IM) {
    Data d;
    d.numInds = numInds;
    d.sparse_data_fill_indices(raw, NC, NB, 0, N1S, N1L, I1, N2S, N2L, I2, NMS, NML, IM);
}
RK void rk_sparse_correct_for_missing_phenotype(const size_t *NS, size_t *NL, uint *I, int M, const uint8_t *usebed, const uint *na, int n_na) {
    Data d;
    d.numNAs = (unsigned)n_na;
    d.NAsInds.assign(na, na + n_na);
    vector<char> ub(M > 0 ? M : 1, 0);
    static_assert(sizeof(bool) == 1, "bool");
    for (int i = 0; i < M; i++) ub[i] = usebed ? (usebed[i] != 0) : 0;
    d.sparse_data_correct_for_missing_phenotype(NS, NL, I, M, reinterpret_cast<const bool *>(ub.data()));
}
}
