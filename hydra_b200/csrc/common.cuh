// Common helpers for the hydra_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hydra_b200.h"

namespace hb {

// ---------------------------------------------------------------- errors
void set_error(const char *fmt, ...);

#define HB_CUDA(x)                                                                                   \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess) {                                                                     \
            hb::set_error("%s:%d CUDA error in %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            return HB_ERR_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

#define HB_CHECK(cond, code, ...)        \
    do {                                 \
        if (!(cond)) {                   \
            hb::set_error(__VA_ARGS__);  \
            return (code);               \
        }                                \
    } while (0)

#define HB_TRY(x)               \
    do {                        \
        int rc_ = (x);          \
        if (rc_ != HB_OK) return rc_; \
    } while (0)

// ---------------------------------------------------------------- record layout
// Individuals (after NA-phenotype compaction) are cut in S slices of L individuals
// (L multiple of 64, L <= 65535); slice c is owned by the CTAs with blockIdx % S == c
// and lives in their shared memory.
//
// SPARSE record of a marker (16-byte aligned):
//   dir[S] : 3 x u32 per slice { word offset of the slice block from the payload base,
//                                n1 | n2 << 16, nm }
//   payload (starts at align16(12*S)): per slice one block of u64 words, each word =
//   4 x u16 slice-local indices; first ceil(n1/4) words hold the "ones", then
//   ceil(n2/4) words the "twos", then ceil(nm/4) words the missing. Staging fills every
//   class in ascending index order (the order of src/data.cpp:1262-1280) and k_bank_order
//   then permutes the indices inside a class for conflict-free gathers; unused lanes of a
//   word hold PAD = L, the index of a dummy slot, anywhere inside the class.
// BED record: S*L/4 bytes, PLINK 2-bit codes at compacted individual positions,
//   pad positions (>= N) hold 11 (genotype 0); slice c starts at byte c*L/4.
// rec[m] = device address | 1 if BED.
constexpr uint32_t kDirWordsPerSlice = 3;
__host__ __device__ inline uint32_t dir_bytes(uint32_t S) { return (12u * S + 15u) & ~15u; }

constexpr int kMaxMix = 16;      // K <= 16 mixture components (incl. zero)
constexpr int kThreads = 512;    // threads per CTA of the sampler kernel
constexpr int kTabCap = 128;     // window items staged per table chunk
constexpr int kUnitCap = 256;    // dot-phase work units per table chunk
constexpr int kChgCap = 64;      // changed markers staged per epsilon-update round

#ifdef __CUDACC__
// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(uint32_t *p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t atom_add_acq_rel_u32(uint32_t *p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
// streaming read of genotype words: read-only path, do not pollute L1
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Grid-wide barrier for a cooperative launch (all CTAs co-resident). `bar` is a
// monotonically increasing arrival counter, `target` the per-CTA running target.
__device__ __forceinline__ void grid_barrier(uint32_t *bar, uint32_t &target, uint32_t nctas) {
    __syncthreads();
    target += nctas;
    if (threadIdx.x == 0) {
        // release: makes the CTA's writes (ordered before by bar.sync) visible; acquire: the other CTAs'
        red_release_add_u32(bar, 1u);
        while (ld_acquire_u32(bar) < target) {
        }
    }
    __syncthreads();
}

// Philox4x32-10 (RNG spec v1; DESIGN.md "Draw tape")
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                           uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
#endif

}  // namespace hb
