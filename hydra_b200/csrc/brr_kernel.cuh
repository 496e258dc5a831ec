// BayesRRm marker-loop kernel (one cooperative, persistent launch per Gibbs
// iteration).  Replaces the `for (j<lmax)` body of BayesRRm::runMpiGibbs
// (reference src/BayesRRm.cpp:1709-2490): sparse_dotprod (:316-342) / LUT dot
// (:1757-1809), the mixture draw (:1744-1921), sparse_scaadd (:250-281) / LUT
// deltaEps (:1976-2010) and the epsilon synchronisation (:2044-2488).
//
// Design: between two synchronisations hydra's tasks all read a
// STALE epsilon, so every marker of a sync window (sync_rate steps x T tasks) is
// independent of the others.  The kernel therefore processes a whole window in
// parallel:
//   * individuals are cut in S slices; CTA (r,c) keeps slice c of the residual in
//     shared memory (r = replica group, R = gridDim/S groups share the window);
//   * item table: k_window_order lays the iteration out in window order (64-byte marker
//     records + per-slice directory entries); cp.async copies the next window's part
//     into the second table during the dot phase, the upper warps cut the slice blocks
//     into work units of <= 128 words during the draw phase;
//   * dot: warp w streams units w, w+16, ... (u16 local indices or 2-bit BED words,
//     coalesced 64-bit loads, two units of words in flight) and gathers from shared memory;
//   * publish: every slice partial goes out as one 16-byte store carrying the window
//     tag next to the data, so the consumer needs neither a fence nor a counter;
//   * draw: marker k of a group is drawn by slice-CTA k % S, one warp per marker
//     (lanes poll the S slots; all terms of the mixture cascade are evaluated at once);
//     changed markers are appended to a list;
//   * ONE grid barrier per window; with several GPUs the changed markers (list
//     entries + genotype records) are then written into every peer's inbox over
//     NVLink as tagged 16-byte units (LL form: no fence, no arrival counter);
//   * update: every CTA applies the changed markers of all GPUs to its slice. The slice is held as 64-bit FIXED POINT
//     (56-bit biased value, grid 2^-sh chosen per launch from sum|eps|; top byte free): integer adds are associative, so
//     every sum of the kernel -- dot products, slice sums, updates -- is exact and ORDER-INDEPENDENT. The changed markers
//     of a window are therefore applied without any ordering (no per-marker barrier, no merge sort by window position,
//     peers' markers as they arrive, any number of them): the words of all changed markers of a chunk are loaded
//     together and applied marker by marker (one CTA barrier per marker keeps two threads from updating the same
//     individual at once); all epsilon replicas on all GPUs stay bit-identical and a run is bit-reproducible.
#pragma once
#include "common.cuh"

namespace hb {

enum : int { MODE_CHAIN = 0, MODE_DOT = 1, MODE_SCAADD = 2 };

// epsilon in shared memory: llrint(eps * 2^sh) as a 64-bit two's complement integer
typedef unsigned long long u64;
// The scalar base term of a launch is summed on its own, finer grid: an error in it is an error of every individual
// (N times in the sum of the residual), so it gets the resolution of a double at magnitude 1; |sum| stays far below 2^10.
constexpr int kOffShift = 52;

struct ChgEnt {  // 32 bytes
    uint32_t p;      // GLOBAL window position: step * T_total + global task (the order hydra's ranks are summed in)
    uint32_t m;      // local marker
    double dbs, mave;
    uint64_t rec;    // record address on the owning GPU / payload offset inside an inbox region
};

// Per iteration, in window order (q = step * T + task): everything the item table needs about the marker of a position,
// gathered once by k_window_order so that the table build inside the marker loop is one streaming read without
// dependent loads (marker -> record -> slice directory used to be three round trips per window).
struct __align__(16) WinMeta {  // 64 bytes
    uint64_t rec;
    double mave, mstd, beta, u, z;
    int32_t m, grp;
    uint32_t rec_bytes;  // size of the record (what a peer GPU needs of it)
    uint32_t pad;
};

constexpr uint32_t kMaxRanks = 8;
constexpr uint32_t kMaxMerged = 4096;        // changed markers of one window that one GPU can ship to its peers (inbox entry region)
// Inbox region of one (window parity, source GPU): 16-byte header {count | window tag << 32}, kMaxMerged list entries and
// the genotype records behind them. Entries and records travel in the LL form: every 16-byte unit is
// {data lo, tag, data hi, tag} with tag = the window's sequence number, so that the receiver can tell, unit by unit,
// whether the data of THIS window have arrived -- no fence and no arrival counter at system scope (a
// red.release.sys per CTA and window cost 4-5 us on NVLink).
constexpr size_t kLLEntry = 2 * sizeof(ChgEnt);
constexpr size_t kDirBase = 16 + (size_t)kMaxMerged * kLLEntry;      // slice directory entries of a list's first kDirCap markers
constexpr uint32_t kDirCap = 64, kDirSlices = 148;                    //   [kDirCap][S] x 2 LL units: {first unit of the block | n1 | n2 << 16, nm}
constexpr size_t kInboxHeader = kDirBase + (size_t)kDirCap * kDirSlices * 32;

// Exchange of the changed markers between the GPUs of one node (replaces MPI_Allreduce of deltaEps,
// src/BayesRRm.cpp:2051, 2456): every GPU pushes (position, deltaBeta*mstd, mave, genotype record) of its
// changed markers into every peer's inbox over NVLink as tagged units; the receiver validates count, entries and records by their tags.
struct PeerComm {
    uint32_t nranks, rank;
    uint32_t T_total, t_first;
    unsigned char *inbox_local;               // [2 parities][nranks sources] regions of inbox_stride bytes
    unsigned char *inbox_peer[kMaxRanks];     // peer-mapped inbox of rank h (unused for self)
    size_t inbox_stride;
    unsigned long long seq_base;              // sequence number of this launch's window 0, minus 1
    const uint32_t *rec_bytes;                // [M] bytes of each local record
    uint32_t *err;                            // != 0: exchange failed (capacity / timeout)
    long long timeout_cycles;
};

struct BrrParams {
    // layout
    uint32_t N, S, L, R, M;
    const uint64_t *rec;   // [M] record address | BED flag
    const double *mave;    // [M]
    const double *mstd;    // [M] inverse sd (BayesRRm)
    const int32_t *grp;    // [M]
    // epsilon
    const double *E_in;    // [S*L] common residual without base terms
    double *E_out;         // [S*L]
    double shift_in;       // folded into E at load (base terms of the previous launch)
    double q_scale, q_inv; // 2^sh and 2^-sh of the launch's fixed-point grid
    double *off_out;       // base-term accumulator of this launch (scalar)
    double *slice_sum_out; // [S]  sum of E over the slice (i < N)
    double *slice_sq_out;  // [S]  sum of E^2
    double *slice_abs_out; // [S]  sum of |E|  -> the next launch's grid
    double *slice_max_out; // [S]  max |E|
    // marker state
    double *beta;          // [M]
    int32_t *comp;         // [M]
    double *acum;          // [M]
    int32_t *cass;         // [G*K]
    // per-iteration inputs, window-ordered: q = j*T + task
    const int32_t *order;  // [lmax*T] local marker or -1
    const double *u;       // [lmax*T]
    const double *z;       // [lmax*T]
    const WinMeta *meta;   // [lmax*T] chain mode: window-ordered marker data (NULL in the unit modes)
    const uint4 *dirw;     // [S][lmax*T] chain mode: slice directory entries {word offset, n1 | n2 << 16, nm, -} of the positions
    uint32_t T, SR, lmax, K, G;
    // hyper-parameter tables [G*K]
    const double *logPi, *chalf, *denom, *sdk;
    const uint8_t *grp_active; // [G]
    double i_2sigE, dNm1;
    // scratch
    uint4 *slots;          // [Wmax*S] slice partials as {lo, tag, hi, tag}: data and flag travel together
    unsigned long long *chg_cnt;  // [3] per window (triple buffered): changed markers | 16-byte units of their records << 32
    ChgEnt *chg_list;      // [3*Wmax] their (position, deltaBeta*mstd, mave, record), in arrival order
    uint32_t *chg_off;     // [3*Wmax] multi-GPU: place of the record in the peers' inboxes (LL units), same order
    uint4 *chg_dir;        // [3*Wmax][S] the changed markers' slice directory entries, same order (every CTA finds its own next to the entry)
    double *dB;            // [2*Wmax] deltaBeta*mstd per window position, double buffered
    double *dMave;         // [2*Wmax] mave of the changed marker
    uint64_t *dRec;        // [2*Wmax] its record
    uint32_t Wmax;
    uint32_t *bar;         // grid barrier counter
    unsigned long long *stats; // [16]: 0 nsync 1 nwindows 2 nnz dot 3 nnz upd 4 bed markers 5 changed, 8.. phase cycles
    // unit modes
    int mode;
    double *num_out;       // MODE_DOT: [W]
    PeerComm pc;
    uint32_t flags;        // bit 0: no L2 prefetch of the next window (developer knob)
    uint32_t n_ahead;      // steps of a window run ahead (>= sync_rate, <= kSpecMax; 0 = sync_rate)
    const double *fh;      // bayesFHMPI: [3*M] per-marker denom, 0.5*log(..), sd of this iteration (k_fh_prepare), else NULL
    unsigned long long *cta_cycles;  // optional [gridDim*8] per-CTA phase cycles (HB_DEBUG_CYCLES=1)
};

struct ItemTab {  // the window positions this CTA group works on, with everything the draw needs
    WinMeta meta[kTabCap];        // chain mode: copied in by cp.async one dot phase ahead (64 bytes per position)
    uint4 dir[kTabCap];           // chain mode: directory entry of this CTA's slice, same copy
    const uint64_t *ptr[kTabCap];
    uint32_t nw[kTabCap];         // u64 words of the slice block
    uint32_t b1[kTabCap];         // first word of class 2   (0xFFFFFFFF = BED block)
    uint32_t b2[kTabCap];         // first word of class "missing"
    // dot-phase work units: the slice blocks cut in pieces of (128 << ushift) words (BED: a quarter of that, a BED word
    // costs 32 gathers), so that the warps of the CTA share the chunk's words evenly whatever the marker sizes are
    uint16_t ucum[kTabCap + 1];   // exclusive prefix of the units per item
    uint32_t nunits, ushift;
    uint32_t tag_base, tag_W, tag_k0, tag_valid;  // which window chunk the table describes
};

struct Blk {
    const uint64_t *ptr;
    uint32_t nw, b1, b2;
    uint32_t n1, n2, nm;        // genotype counts of the slice block (sparse records; BED: 0)
};

// slice block c of a marker record (common.cuh "record layout")
template <bool kCoherent = false>
__device__ __forceinline__ Blk decode_block(uint64_t rr, uint32_t c, uint32_t S, uint32_t L) {
    Blk b;
    if (rr & 1ull) {
        b.ptr = reinterpret_cast<const uint64_t *>(rr & ~15ull) + (size_t)c * (L / 32);
        b.nw = L / 32;
        b.b1 = 0xFFFFFFFFu;
        b.b2 = 0xFFFFFFFFu;
        b.n1 = 0; b.n2 = 0; b.nm = 0;
    } else {
        const uint8_t *bp = reinterpret_cast<const uint8_t *>(rr);
        const uint32_t *dir = reinterpret_cast<const uint32_t *>(bp) + c * 3;
        // kCoherent: the record may sit in an inbox that a peer GPU rewrites during the launch -> read through L2
        const uint32_t st = kCoherent ? __ldcg(dir) : __ldg(dir), n12 = kCoherent ? __ldcg(dir + 1) : __ldg(dir + 1);
        const uint32_t nm = kCoherent ? __ldcg(dir + 2) : __ldg(dir + 2);
        const uint32_t w1 = ((n12 & 0xFFFFu) + 3) / 4, w2 = ((n12 >> 16) + 3) / 4, wm = (nm + 3) / 4;
        b.ptr = reinterpret_cast<const uint64_t *>(bp + dir_bytes(S)) + st;
        b.b1 = w1;
        b.b2 = w1 + w2;
        b.nw = w1 + w2 + wm;
        b.n1 = n12 & 0xFFFFu; b.n2 = n12 >> 16; b.nm = nm;
    }
    return b;
}

// the same from a directory entry that is already at hand ({word offset, n1 | n2 << 16, nm, -})
__device__ __forceinline__ Blk block_from_dir(uint64_t rr, const uint4 dv, uint32_t c, uint32_t S, uint32_t L) {
    Blk b;
    if (rr & 1ull) {
        b.ptr = reinterpret_cast<const uint64_t *>(rr & ~15ull) + (size_t)c * (L / 32);
        b.nw = L / 32; b.b1 = 0xFFFFFFFFu; b.b2 = 0xFFFFFFFFu; b.n1 = 0; b.n2 = 0; b.nm = 0;
    } else {
        const uint32_t w1 = ((dv.y & 0xFFFFu) + 3) / 4, w2 = ((dv.y >> 16) + 3) / 4, wm = (dv.z + 3) / 4;
        b.ptr = reinterpret_cast<const uint64_t *>(reinterpret_cast<const uint8_t *>(rr) + dir_bytes(S)) + dv.x;
        b.b1 = w1; b.b2 = w1 + w2; b.nw = w1 + w2 + wm;
        b.n1 = dv.y & 0xFFFFu; b.n2 = dv.y >> 16; b.nm = dv.z;
    }
    return b;
}
__device__ __forceinline__ void st_ll(uint4 *dst, uint64_t v, uint32_t tag) {
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"((uint32_t)v), "r"(tag), "r"((uint32_t)(v >> 32)), "r"(tag) : "memory");
}
// waits for unit `src` of this window; *err is raised if it does not come (dead peer)
__device__ __forceinline__ uint64_t ld_ll(const uint4 *src, uint32_t tag, uint32_t *err) {
    uint4 v;
    for (uint32_t it = 0;; it++) {
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src) : "memory");
        if (v.y == tag && v.w == tag) break;
        if (it > (1u << 26)) { atomicExch(err, 2u); break; }
    }
    return ((uint64_t)v.z << 32) | v.x;
}
// three consecutive units, requested together
__device__ __forceinline__ void ld_ll3(const uint4 *src, uint32_t tag, uint32_t *err, uint64_t &d0, uint64_t &d1, uint64_t &d2) {
    uint4 a, b, c;
    for (uint32_t it = 0;; it++) {
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(src) : "memory");
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(src + 1) : "memory");
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(src + 2) : "memory");
        if (a.y == tag && a.w == tag && b.y == tag && b.w == tag && c.y == tag && c.w == tag) break;
        if (it > (1u << 26)) { atomicExch(err, 2u); break; }
    }
    d0 = ((uint64_t)a.z << 32) | a.x; d1 = ((uint64_t)b.z << 32) | b.x; d2 = ((uint64_t)c.z << 32) | c.x;
}
// four units at different places, requested together
__device__ __forceinline__ void ld_ll4(const uint4 *p0, const uint4 *p1, const uint4 *p2, const uint4 *p3, uint32_t tag, uint32_t *err,
                                       uint64_t &d0, uint64_t &d1, uint64_t &d2, uint64_t &d3) {
    uint4 a, b, c, d;
    for (uint32_t it = 0;; it++) {
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p0) : "memory");
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p1) : "memory");
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(p2) : "memory");
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "l"(p3) : "memory");
        if (a.y == tag && a.w == tag && b.y == tag && b.w == tag && c.y == tag && c.w == tag && d.y == tag && d.w == tag) break;
        if (it > (1u << 26)) { atomicExch(err, 2u); break; }
    }
    d0 = ((uint64_t)a.z << 32) | a.x; d1 = ((uint64_t)b.z << 32) | b.x; d2 = ((uint64_t)c.z << 32) | c.x; d3 = ((uint64_t)d.z << 32) | d.x;
}
// slice block of a record in an inbox from the directory units its sender shipped with the list
__device__ __forceinline__ Blk block_from_ll_dir(const uint4 *payload, uint64_t da, uint64_t db, uint32_t L) {
    Blk b;
    const uint32_t n12 = (uint32_t)(da >> 32);
    b.ptr = reinterpret_cast<const uint64_t *>(payload + (uint32_t)da);
    if (n12 == 0xFFFFFFFFu) {
        b.nw = L / 32; b.b1 = 0xFFFFFFFFu; b.b2 = 0xFFFFFFFFu; b.n1 = 0; b.n2 = 0; b.nm = 0;
    } else {
        const uint32_t nm = (uint32_t)db;
        const uint32_t w1 = ((n12 & 0xFFFFu) + 3) / 4, w2 = ((n12 >> 16) + 3) / 4, wm = (nm + 3) / 4;
        b.b1 = w1; b.b2 = w1 + w2; b.nw = w1 + w2 + wm;
        b.n1 = n12 & 0xFFFFu; b.n2 = n12 >> 16; b.nm = nm;
    }
    return b;
}
__device__ __forceinline__ uint32_t ld_ll_u32(const uint4 *rec, uint32_t byte_off, uint32_t tag, uint32_t *err) {
    const uint64_t v = ld_ll(rec + byte_off / 8u, tag, err);
    return (byte_off & 4u) ? (uint32_t)(v >> 32) : (uint32_t)v;
}
__device__ __forceinline__ uint64_t ld_l2_u64(const uint64_t *p) {  // coherent at L2 (peer-written inbox data)
    return __ldcg(reinterpret_cast<const unsigned long long *>(p));
}
__device__ __forceinline__ uint4 ld_nc_v4(const uint4 *p) {  // read-only for the lifetime of the kernel
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// bulk form: [p, p + bytes) into L2, both multiples of 16
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---- slice block dot product on the fixed-point slice. Integer sums are exact: the lanes', warps' and units' shares can
//      be added in any order. Two accumulators per marker: a12 = sum_1 + 2 sum_2, am = sum_missing.
__device__ __forceinline__ u64 gather4(uint64_t x, const u64 *__restrict__ Eq) {
    return (Eq[x & 0xFFFFu] + Eq[(x >> 16) & 0xFFFFu]) + (Eq[(x >> 32) & 0xFFFFu] + Eq[x >> 48]);
}
// BED word: the lane owns 32 consecutive individuals (word w of the slice block).
// Shared-memory banks: individual k of ANY word sits in 8-byte bank k mod 16, so lanes that walk their words from the same end
// hit the same bank (16-way conflicts on a dense marker). Every lane therefore walks its word from its own starting point:
// the masks are rotated by 2*(lane & 15) bits, find-first-set runs on the rotated masks, and the position is rotated back.
// The sums are integer: the order of the additions does not matter.
__constant__ uint32_t g_bed_norot;   // developer knob (HB_NO_BED_ROT=1): walk every word from bit 0
__device__ __forceinline__ uint32_t rotr32(uint32_t x, uint32_t r) { return __funnelshift_r(x, x, r); }
__device__ __forceinline__ void dot_bed_word(uint64_t bits, uint32_t w, const u64 *__restrict__ Eq, u64 &a12, u64 &am) {
    // PLINK codes (b1 b0): 00 -> 2, 10 -> 1, 11 -> 0, 01 -> missing. Only the non-zero genotypes are
    // visited: A = individuals with b0 == 0 (genotype 1 or 2), of which those with b1 == 0 as well count twice; the
    // missing ones (b0 == 1, b1 == 0) are rare. The trip count is the lane's number of non-zeros, not 32.
    if (bits == ~0ull) return;  // 32 x genotype 0
    const u64 *e = Eq + 32u * w;
    const uint32_t r2 = g_bed_norot ? 0u : (threadIdx.x & 15u) * 2u;
#pragma unroll
    for (uint32_t h = 0; h < 2; h++) {  // 16 individuals per 32-bit half
        const uint32_t v = (uint32_t)(bits >> (32u * h));
        const uint32_t b0 = v & 0x55555555u, b1 = (v >> 1) & 0x55555555u;
        uint32_t a = rotr32(~b0 & 0x55555555u, r2);      // genotype 1 or 2
        const uint32_t two = rotr32(~b0 & ~b1 & 0x55555555u, r2);   // genotype 2
        uint32_t miss = rotr32(b0 & ~b1, r2);
        const u64 *eh = e + 16u * h;
        while (a) {
            const uint32_t pos = __ffs((int)a) - 1u;  // even bit position = 2 x individual (rotated)
            a &= a - 1u;
            const u64 x = eh[((pos + r2) & 31u) >> 1];
            a12 += x << ((two >> pos) & 1u);          // genotype 2 counts twice
        }
        while (miss) {
            const uint32_t pos = __ffs((int)miss) - 1u;
            miss &= miss - 1u;
            am += eh[((pos + r2) & 31u) >> 1];
        }
    }
}

// Work unit of the dot phase, one 16-byte descriptor in shared memory:
//   x, y  address of the unit's first 64-bit word
//   z     number of words | (first word of class "twos", relative, clipped) << 16      (0xFFFF: BED unit)
//   w     first word of class "missing" (relative, clipped; BED: the unit's first word inside the slice block) | item << 16
__device__ __forceinline__ uint4 make_unit(const uint64_t *ptr, uint32_t nwords, uint32_t b1rel, uint32_t b2rel, uint32_t k) {
    const uint64_t a = (uint64_t)(uintptr_t)ptr;
    return make_uint4((uint32_t)a, (uint32_t)(a >> 32), nwords | (b1rel << 16), b2rel | (k << 16));
}
// first (normally only) 128 words of a sparse unit / 32 words of a BED unit; lanes past the end get the pad word,
// whose four indices point at the dummy slot (always 0 during the dot phase)
__device__ __forceinline__ void load_unit(const uint4 d, uint64_t (&x)[4], uint32_t lane, uint64_t padw) {
    const uint64_t *ptr = reinterpret_cast<const uint64_t *>(((uint64_t)d.y << 32) | d.x);
    const uint32_t nwords = d.z & 0xFFFFu;
#pragma unroll
    for (uint32_t t = 0; t < 4; t++) {
        const uint32_t w = lane + 32u * t;
        x[t] = padw;
        if (w < nwords) x[t] = ld_stream_u64(ptr + w);
    }
}
__device__ __forceinline__ void dot_words(const uint64_t (&x)[4], uint32_t w0, uint32_t nwords, uint32_t b1, uint32_t b2,
                                          const u64 *__restrict__ Eq, uint32_t lane, u64 &a12, u64 &am) {
#pragma unroll
    for (uint32_t t = 0; t < 4; t++) {
        // most slice blocks are short (rare variants): a group of 32 words that lies entirely past the end is skipped
        // (warp-uniform test) instead of gathering the pad slot 4 x 32 times
        const uint32_t wb = w0 + 32u * t;
        if (wb < nwords) {
            const uint32_t w = wb + lane;
            const u64 g = gather4(x[t], Eq);      // pad lanes: 4 x the dummy slot = 0
            // class weights as small integer factors (two multiply-adds per accumulator, no branches, no selects of 64-bit
            // values): genotype 1 -> a12 += g, genotype 2 -> a12 += 2g, missing -> am += g
            const uint32_t wt = (w < b1) ? 1u : ((w < b2) ? 2u : 0u), wm = (w < b2) ? 0u : 1u;
            a12 += g * (u64)wt;
            am += g * (u64)wm;
        }
    }
}
// all of a unit: the lane's share of (a12, am)
__device__ __forceinline__ void dot_unit(const uint4 d, const uint64_t (&x)[4], const u64 *__restrict__ Eq,
                                         uint32_t lane, uint64_t padw, u64 &a12, u64 &am) {
    const uint32_t nwords = d.z & 0xFFFFu, b1 = d.z >> 16, b2 = d.w & 0xFFFFu;
    const uint64_t *ptr = reinterpret_cast<const uint64_t *>(((uint64_t)d.y << 32) | d.x);
    if (b1 == 0xFFFFu) {  // BED: b2 = first word of the unit inside the slice block
#pragma unroll
        for (uint32_t t = 0; t < 4; t++)  // load_unit has fetched up to four words per lane
            if (lane + 32u * t < nwords) dot_bed_word(x[t], b2 + lane + 32u * t, Eq, a12, am);
        // the rest of a long unit: four words per lane requested together (one L2 / HBM round trip per 128 words, not per 32)
        for (uint32_t w0 = 128u; w0 < nwords; w0 += 128u) {
            uint64_t y[4];
#pragma unroll
            for (uint32_t t = 0; t < 4; t++) {
                const uint32_t w = w0 + lane + 32u * t;
                y[t] = ~0ull;                              // 32 x genotype 0: skipped
                if (w < nwords) y[t] = ld_stream_u64(ptr + w);
            }
#pragma unroll
            for (uint32_t t = 0; t < 4; t++) dot_bed_word(y[t], b2 + w0 + lane + 32u * t, Eq, a12, am);
        }
        return;
    }
    dot_words(x, 0u, nwords, b1, b2, Eq, lane, a12, am);
    if (nwords > 128u) {  // units longer than one step (very heavy chunks only)
        for (uint32_t w0 = 128u; w0 < nwords; w0 += 128u) {
            uint64_t y[4];
#pragma unroll
            for (uint32_t t = 0; t < 4; t++) {
                const uint32_t w = w0 + lane + 32u * t;
                y[t] = padw;
                if (w < nwords) y[t] = ld_stream_u64(ptr + w);
            }
            dot_words(y, w0, nwords, b1, b2, Eq, lane, a12, am);
        }
    }
}

__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double block_sum(double v, double *red /*[32]*/) {
    v = warp_sum(v);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    const uint32_t nw = blockDim.x >> 5;
    for (uint32_t w = 0; w < nw; w++) s += red[w];  // fixed order: identical in every CTA
    return s;
}
__device__ __forceinline__ u64 block_sum_u64(u64 v, double *red /*[32]*/) {
    v = warp_sum_u64(v);
    u64 *r = reinterpret_cast<u64 *>(red);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) r[warp] = v;
    __syncthreads();
    u64 s = 0;
    const uint32_t nw = blockDim.x >> 5;
    for (uint32_t w = 0; w < nw; w++) s += r[w];
    return s;
}
__device__ __forceinline__ double block_max(double v, double *red /*[32]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    const uint32_t nw = blockDim.x >> 5;
    for (uint32_t w = 0; w < nw; w++) s = fmax(s, red[w]);
    return s;
}

// slice block c of a record that sits in an inbox in the LL form; the returned ptr counts 16-byte units
__device__ __forceinline__ Blk decode_block_ll(const uint4 *rec, bool bed, uint32_t c, uint32_t S, uint32_t L, uint32_t tag, uint32_t *err) {
    Blk b;
    if (bed) {
        b.ptr = reinterpret_cast<const uint64_t *>(rec + (size_t)c * (L / 32));
        b.nw = L / 32; b.b1 = 0xFFFFFFFFu; b.b2 = 0xFFFFFFFFu;
        b.n1 = 0; b.n2 = 0; b.nm = 0;
    } else {
        // the 12-byte directory entry lies in one or two 8-byte words: both units are requested together
        const uint32_t bo = c * 12u, u0 = bo / 8u, u1 = (bo + 8u) / 8u;
        uint4 a, d;
        for (uint32_t it = 0;; it++) {
            asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(rec + u0) : "memory");
            asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "l"(rec + u1) : "memory");
            if (a.y == tag && a.w == tag && d.y == tag && d.w == tag) break;
            if (it > (1u << 26)) { atomicExch(err, 2u); break; }
        }
        uint32_t st, n12, nm;
        if (bo & 4u) { st = a.z; n12 = d.x; nm = d.z; }   // entry starts in the upper half of word u0, rest = word u1 = u0 + 1
        else { st = a.x; n12 = a.z; nm = d.x; }           // u1 = u0 + 1: third field in its lower half
        const uint32_t w1 = ((n12 & 0xFFFFu) + 3) / 4, w2 = ((n12 >> 16) + 3) / 4, wm = (nm + 3) / 4;
        b.ptr = reinterpret_cast<const uint64_t *>(rec + dir_bytes(S) / 8u + st);
        b.b1 = w1; b.b2 = w1 + w2; b.nw = w1 + w2 + wm;
        b.n1 = n12 & 0xFFFFu; b.n2 = n12 >> 16; b.nm = nm;
    }
    return b;
}

struct HypTabs {
    const double *logPi, *chalf, *denom, *sdk;
};

__device__ __forceinline__ ChgEnt ld_chg_ent(const ChgEnt *e) {
    const uint4 *src = reinterpret_cast<const uint4 *>(e);
    const uint4 a = __ldcg(src), b = __ldcg(src + 1);
    ChgEnt en;
    en.p = a.x; en.m = a.y;
    en.dbs = __longlong_as_double((long long)(((unsigned long long)a.w << 32) | a.z));
    en.mave = __longlong_as_double((long long)(((unsigned long long)b.y << 32) | b.x));
    en.rec = ((unsigned long long)b.w << 32) | b.z;
    return en;
}
__device__ __forceinline__ uint4 ld_slot(const uint4 *p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_slot(uint4 *p, double val, uint32_t tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(val);
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"((uint32_t)b), "r"(tag), "r"((uint32_t)(b >> 32)), "r"(tag) : "memory");
}

// ---- mixture draw of one marker (src/BayesRRm.cpp:1721-1933), by one warp ---------
// The warp first collects the S slice partials of table entry k (window position p): every slot carries
// the window tag next to the data, so no fence or arrival counter is needed. Lane kk < K then evaluates
// mixture component kk; sums over components are taken in component order (same order as the reference).
__device__ __forceinline__ void draw_marker_warp(const BrrParams &P, const ItemTab *tab, uint32_t k, const HypTabs &H, uint32_t p, uint32_t q,
                                 uint32_t dbuf, uint32_t buf3, uint32_t tag, uint32_t lane, long long *tp, unsigned long long *gs = nullptr) {
    long long t0_ = tp ? clock64() : 0ll;
#define HB_GS(i) do { if (gs && lane == 0) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); gs[i] = g_; } } while (0)
    HB_GS(0);
    const uint4 *sl = P.slots + (size_t)p * P.S;
    // A marker with a non-zero effect will change: its place in the window's list of changed markers is reserved now, so
    // that the round trip of the atomic overlaps the wait for the partials and the draw.
    // The same atomic reserves the place of its record in the peers' inboxes (multi-GPU).
    const unsigned long long resv = 1ull | ((P.pc.nranks > 1) ? ((unsigned long long)(tab->meta[k].rec_bytes >> 3) << 32) : 0ull);   // (read again below with the rest)  // LL units = 8-byte words
    unsigned long long idx_early = ~0ull;   // (meaningful in lane 0)
    if (lane == 0 && P.mode == MODE_CHAIN && tab->meta[k].beta != 0.0) idx_early = atomicAdd(P.chg_cnt + buf3, resv);
    // The marker's slice directory entries (one per lane) are requested now, for every marker: if it changes, they are
    // stored next to its list entry so that the update needs no dependent read (the load is hidden by the wait below).
    uint4 mydir = make_uint4(0u, 0u, 0u, 0u);
    const size_t Qn = (size_t)P.lmax * P.T;
    if (P.mode == MODE_CHAIN && P.dirw && lane < P.S) mydir = ld_nc_v4(P.dirw + (size_t)lane * Qn + q);
    // everything the draw needs from the table is read before the wait (the poll's memory clobber would hold the loads back)
    const WinMeta wm = tab->meta[k];
    const bool g_active = (P.mode == MODE_DOT) ? true : (P.grp_active[wm.grp] != 0);
    // bayesFHMPI: the marker's own prior (src/BayesRRm.cpp:1730, 1748, 1871) replaces the per-(group, component) tables
    double fh_den = 0.0, fh_ch = 0.0, fh_sd = 0.0;
    if (P.fh) { const double *f = P.fh + 3 * (size_t)wm.m; fh_den = __ldg(f); fh_ch = __ldg(f + 1); fh_sd = __ldg(f + 2); }
    double acc = 0.0;
    for (uint32_t c0 = 0; c0 < P.S; c0 += 32) {
        const uint32_t cc = c0 + lane;
        if (cc < P.S) {
            uint4 v;
            do { v = ld_slot(sl + cc); } while (v.y != tag || v.w != tag);
            acc += __longlong_as_double((long long)(((unsigned long long)v.z << 32) | v.x));
        }
    }
    const double sum = warp_sum(acc);  // xor tree: fixed order
    if (tp && lane == 0) { const long long t_ = clock64(); tp[0] += t_ - t0_; t0_ = t_; }
    HB_GS(1);
    const double mstd = wm.mstd;
    (void)dbuf;
    if (P.mode == MODE_DOT) {
        if (lane == 0) P.num_out[q] = __dmul_rn(mstd, sum);
        return;
    }
    const int32_t m = wm.m;
    const int g = wm.grp;
    const uint32_t K = P.K;
    const double beta_old = wm.beta;
    double beta_new = 0.0, acum0 = 1.0;
    int comp = -1;
    if (g_active) {
        // num = mstd*(...) ; num += beta*(N-1)            (:1809/:316-342, :1855)
        const double num = __dadd_rn(__dmul_rn(mstd, sum), __dmul_rn(beta_old, P.dNm1));
        const uint32_t kk = (lane < K) ? lane : 0;
        double muk = 0.0, logL = H.logPi[g * K + kk];
        if (kk > 0) {
            muk = num / (P.fh ? fh_den : H.denom[g * K + kk]);                        // :1859
            // log(pi) - 0.5*log(...) + muk*num*i_2sigE, evaluated left to right        (:1874-1876; FH :1869-1872)
            logL = __dadd_rn(__dadd_rn(logL, -(P.fh ? fh_ch : H.chalf[g * K + kk])), __dmul_rn(__dmul_rn(muk, num), P.i_2sigE));
        }
        const double prob = wm.u;                                                         // :1880
        if (K * K <= 32u) {
            // All K terms of the cascade at once: lane j*K + i evaluates exp(logL[i] - logL[j]); term[j] = 1 / sum_i (in
            // component order), or 0 where the reference's 700-test fires (:1884-1890 for j = 0, :1915-1919 for j > 0).
            const uint32_t j = lane / K, i = lane % K;
            const double Li = __shfl_sync(0xffffffffu, logL, i), Lj = __shfl_sync(0xffffffffu, logL, j % K);
            const double e = exp(Li - Lj);
            const bool bigp = (j < K) && (i >= max(j, 1u)) && (fabs(Li - Lj) > 700.0);
            const uint32_t bigm = __ballot_sync(0xffffffffu, bigp);
            double s = 0.0;
            for (uint32_t i2 = 0; i2 < K; i2++) s += __shfl_sync(0xffffffffu, e, (j * K + i2) & 31u);
            const bool big = ((bigm >> ((j * K) & 31u)) & ((1u << K) - 1u)) != 0u;
            const double term = big ? 0.0 : 1.0 / s;
            double acum = __shfl_sync(0xffffffffu, term, 0);
            acum0 = acum;                                                            // Acum(marker), :1892
            for (uint32_t c = 0; c < K; c++) {                                       // :1894-1921
                if (prob <= acum || c == K - 1) { comp = (int)c; break; }
                acum += __shfl_sync(0xffffffffu, term, ((c + 1u) * K) & 31u);
            }
        } else {
        double ref = __shfl_sync(0xffffffffu, logL, 0);
        bool big = __ballot_sync(0xffffffffu, lane > 0 && lane < K && fabs(logL - ref) > 700.0) != 0u;  // :1884
        double e = exp(logL - ref), s = 0.0;
        for (uint32_t i = 0; i < K; i++) s += __shfl_sync(0xffffffffu, e, i);
        double acum = big ? 0.0 : 1.0 / s;
        acum0 = acum;                                                                // Acum(marker), :1892
        for (uint32_t c = 0; c < K; c++) {                                           // :1894-1921 (uniform across the warp)
            if (prob <= acum || c == K - 1) { comp = (int)c; break; }
            ref = __shfl_sync(0xffffffffu, logL, c + 1);
            big = __ballot_sync(0xffffffffu, lane > c && lane < K && fabs(logL - ref) > 700.0) != 0u;   // :1915
            e = exp(logL - ref);
            s = 0.0;
            for (uint32_t i = 0; i < K; i++) s += __shfl_sync(0xffffffffu, e, i);
            if (!big) acum += 1.0 / s;
        }
        }
        const double muc = __shfl_sync(0xffffffffu, muk, comp);
        if (comp > 0) beta_new = __dadd_rn(muc, __dmul_rn(P.fh ? fh_sd : H.sdk[g * K + comp], wm.z));  // :1901
    }                                                                                // else :1924-1925
    if (tp && lane == 0) { const long long t_ = clock64(); tp[1] += t_ - t0_; t0_ = t_; }
    HB_GS(2);
    {   // changed marker: the directory entries go next to the list entry (all lanes), then lane 0 writes the entry
        const double dbeta_w = beta_old - beta_new;
        const bool chg_w = (dbeta_w != 0.0 || idx_early != ~0ull);   // lane 0's view decides
        if (__shfl_sync(0xffffffffu, (int)chg_w, 0)) {
            unsigned long long got_w = 0ull;
            if (lane == 0) got_w = (idx_early != ~0ull) ? idx_early : atomicAdd(P.chg_cnt + buf3, resv);
            got_w = __shfl_sync(0xffffffffu, got_w, 0);
            idx_early = got_w;   // lane 0 uses it below
            uint4 *dd = P.chg_dir + ((size_t)buf3 * P.Wmax + (uint32_t)got_w) * P.S;
            if (lane < P.S) dd[lane] = mydir;
            for (uint32_t cc = lane + 32u; cc < P.S; cc += 32u) dd[cc] = ld_nc_v4(P.dirw + (size_t)cc * Qn + q);
            __syncwarp();
        }
    }
    if (lane == 0) {
        // (cass, :1904, is counted from comp[] after the marker loop, k_beta_sqnorm: a draw can be discarded and repeated, see
        // "windows run ahead" in the kernel)
        if (comp >= 0) P.comp[m] = comp;
        const double dbeta = beta_old - beta_new;                                    // :1933
        if (dbeta != 0.0 || idx_early != ~0ull) {
            // (a reserved entry of a marker that drew its old value again carries 0 and does not count as a change)
            const double dbs = (dbeta != 0.0) ? __dmul_rn(dbeta, mstd) : 0.0;
            const unsigned long long got = (idx_early != ~0ull) ? idx_early : atomicAdd(P.chg_cnt + buf3, resv);
            const uint32_t idx = (uint32_t)got, off_units = (uint32_t)(got >> 32);
            ChgEnt en;
            en.p = (P.pc.nranks > 1) ? (p / P.T) * P.pc.T_total + P.pc.t_first + (p % P.T) : p;
            en.m = (uint32_t)m; en.dbs = dbs; en.mave = wm.mave; en.rec = wm.rec;
            P.chg_list[(size_t)buf3 * P.Wmax + idx] = en;
            if (P.pc.nranks > 1) P.chg_off[(size_t)buf3 * P.Wmax + idx] = off_units;  // pushed to the peers after the grid barrier
        }
        P.beta[m] = beta_new;
        P.acum[m] = acum0;
        if (tp) { const long long t_ = clock64(); tp[2] += t_ - t0_; }
    }
}

__device__ __forceinline__ void named_barrier(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Item table of one window chunk, step 1 (chain mode): asynchronous copy of the window-ordered marker data and of this
// slice's directory entries into the table; one thread per window position of this CTA group, no registers held.
__device__ __forceinline__ void stage_items(ItemTab *tab, const BrrParams &P, uint32_t r, uint32_t c, uint32_t base,
                                            uint32_t W, uint32_t k0, uint32_t tl, uint32_t nt) {
    const uint32_t n_items = (W > r) ? (W - r + P.R - 1) / P.R : 0;
    const uint32_t nk = (n_items > k0) ? min((uint32_t)kTabCap, n_items - k0) : 0;
    const uint32_t Q = P.lmax * P.T;
    for (uint32_t k = tl; k < nk; k += nt) {
        const uint32_t q = base + r + P.R * (k0 + k);
        const uint4 *mp = reinterpret_cast<const uint4 *>(P.meta + q);
        uint4 *dst = reinterpret_cast<uint4 *>(&tab->meta[k]);
        cp_async16(dst, mp); cp_async16(dst + 1, mp + 1); cp_async16(dst + 2, mp + 2); cp_async16(dst + 3, mp + 3);
        cp_async16(&tab->dir[k], P.dirw + (size_t)c * Q + q);
    }
}
// Unit modes (no window-ordered data): gather the marker data of the positions directly.
__device__ __forceinline__ void fill_items_sync(ItemTab *tab, const BrrParams &P, uint32_t r, uint32_t c, uint32_t base,
                                                uint32_t W, uint32_t k0, uint32_t tl, uint32_t nt) {
    const uint32_t n_items = (W > r) ? (W - r + P.R - 1) / P.R : 0;
    const uint32_t nk = (n_items > k0) ? min((uint32_t)kTabCap, n_items - k0) : 0;
    for (uint32_t k = tl; k < nk; k += nt) {
        const uint32_t p = r + P.R * (k0 + k);
        const int32_t m = P.order[base + p];
        WinMeta wm;
        wm.m = m; wm.rec = 0; wm.mave = 0.0; wm.mstd = 0.0; wm.beta = 0.0; wm.u = 0.0; wm.z = 0.0; wm.grp = 0; wm.rec_bytes = 0; wm.pad = 0;
        tab->nw[k] = 0;
        if (m >= 0) {
            wm.rec = P.rec[m]; wm.mave = P.mave[m]; wm.mstd = P.mstd[m]; wm.beta = P.beta[m]; wm.grp = P.grp[m];
            if (P.pc.rec_bytes) wm.rec_bytes = P.pc.rec_bytes[m];
            wm.u = P.u[base + p]; wm.z = P.z[base + p];
            const Blk b = decode_block(wm.rec, c, P.S, P.L);
            tab->ptr[k] = b.ptr; tab->nw[k] = b.nw; tab->b1[k] = b.b1; tab->b2[k] = b.b2;
            tab->dir[k] = make_uint4(0u, b.n1 | (b.n2 << 16), b.nm, 0u);   // genotype counts of the block: bias of the dot product
        }
        tab->meta[k] = wm;
    }
}

// Step 2: slice block of every item from the staged data (prefetched into L2), work units of the dot phase, tags.
// Threads [t0, t0+nt) take part.
__device__ __forceinline__ void finish_table(ItemTab *tab, uint4 *udesc, const BrrParams &P, uint32_t r, uint32_t c, uint32_t base,
                                             uint32_t W, uint32_t k0, uint32_t t0, uint32_t nt, bool prefetch) {
    const uint32_t n_items = (W > r) ? (W - r + P.R - 1) / P.R : 0;
    const uint32_t nk = (n_items > k0) ? min((uint32_t)kTabCap, n_items - k0) : 0;
    const uint32_t tl = threadIdx.x - t0;
    if (P.meta) {
        for (uint32_t k = tl; k < nk; k += nt) {
            tab->nw[k] = 0;
            if (tab->meta[k].m >= 0) {
                const uint64_t rr = tab->meta[k].rec;
                Blk b;
                if (rr & 1ull) {
                    b = decode_block(rr, c, P.S, P.L);
                } else {
                    const uint4 dv = tab->dir[k];
                    const uint32_t w1 = ((dv.y & 0xFFFFu) + 3) / 4, w2 = ((dv.y >> 16) + 3) / 4, wm = (dv.z + 3) / 4;
                    b.ptr = reinterpret_cast<const uint64_t *>(reinterpret_cast<const uint8_t *>(rr) + dir_bytes(P.S)) + dv.x;
                    b.b1 = w1; b.b2 = w1 + w2; b.nw = w1 + w2 + wm;
                }
                tab->ptr[k] = b.ptr; tab->nw[k] = b.nw; tab->b1[k] = b.b1; tab->b2[k] = b.b2;
                if (prefetch && b.nw) {   // one bulk request per block (the copy engine's path, not the load/store unit's queue)
                    const uintptr_t pa = reinterpret_cast<uintptr_t>(b.ptr) & ~(uintptr_t)15;
                    const uintptr_t pe = (reinterpret_cast<uintptr_t>(b.ptr) + (size_t)b.nw * 8 + 15) & ~(uintptr_t)15;
                    prefetch_l2_bulk(reinterpret_cast<const void *>(pa), (uint32_t)(pe - pa));
                }
            }
        }
    }
    // ---- work units of the dot phase: prefix of the units per item, unit -> (item, piece)
    named_barrier(1, nt);
    if (tl < 32) {
        uint32_t cw[4];  // cost of the item's block in sparse-word equivalents
#pragma unroll
        for (uint32_t i = 0; i < 4; i++) {
            const uint32_t k = tl * 4 + i;
            cw[i] = (k < nk) ? ((tab->b1[k] == 0xFFFFFFFFu) ? tab->nw[k] * 4u : tab->nw[k]) : 0u;
        }
        uint32_t sh = 0, un[4], tot;
        for (;; sh++) {
            const uint32_t span = 128u << sh;
            tot = 0;
#pragma unroll
            for (uint32_t i = 0; i < 4; i++) { un[i] = (cw[i] + span - 1) / span; tot += un[i]; }
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tl >= (uint32_t)o) incl += t;
            }
            const uint32_t all = __shfl_sync(0xffffffffu, incl, 31);
            if (all <= (uint32_t)kUnitCap) {
                uint32_t a = incl - tot;
#pragma unroll
                for (uint32_t i = 0; i < 4; i++) { tab->ucum[tl * 4 + i] = (uint16_t)a; a += un[i]; }
                if (tl == 31) { tab->ucum[kTabCap] = (uint16_t)all; tab->nunits = all; tab->ushift = sh; }
                break;
            }
        }
    }
    named_barrier(1, nt);
    if (tl < nk) {  // descriptors of the item's units (the unit table is free between two dot phases)
        const uint32_t u0 = tab->ucum[tl], u1 = tab->ucum[tl + 1], nw = tab->nw[tl], b1 = tab->b1[tl], b2 = tab->b2[tl];
        const bool bed = (b1 == 0xFFFFFFFFu);
        const uint32_t span = bed ? (32u << tab->ushift) : (128u << tab->ushift);
        for (uint32_t u = u0; u < u1; u++) {
            const uint32_t w0 = (u - u0) * span, n = min(span, nw - w0);
            const uint32_t b1r = bed ? 0xFFFFu : ((b1 > w0) ? min(b1 - w0, n) : 0u);
            const uint32_t b2r = bed ? w0 : ((b2 > w0) ? min(b2 - w0, n) : 0u);
            udesc[u] = make_unit(tab->ptr[tl] + w0, n, b1r, b2r, tl);
        }
    }
    if (tl == 0) { tab->tag_base = base; tab->tag_W = W; tab->tag_k0 = k0; tab->tag_valid = 1; }
}

// last index x with cum[x] <= f (cum ascending, cum[0] = 0, f < cum[n])
__device__ __forceinline__ uint32_t find_entry(const uint32_t *cum, uint32_t n, uint32_t f) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (cum[mid] <= f) lo = mid; else hi = mid;
    }
    return lo;
}

constexpr uint32_t kApplyQ = 4;          // 64-bit words staged in registers per thread and round
constexpr uint32_t kHypSmem = 64;        // G*K up to this: hyper-parameter tables live in shared memory
constexpr uint32_t kWarps = kThreads / 32;
constexpr uint32_t kSpecMax = 64;        // windows run ahead: up to this many steps
constexpr uint32_t kDrawWarps = 8;       // warps that collect partials and draw; the others prebuild the next table

// clock read that cannot be scheduled before `dep` is available (developer timing of phases that end with loads in flight)
__device__ __forceinline__ long long clock_after(uint32_t dep) {
    long long t;
    asm volatile("mov.u64 %0, %%clock64; // %1" : "=l"(t) : "r"(dep) : "memory");
    return t;
}
// ---- epsilon update -------------------------------------------------------------------------------------------------
struct ChgTab {  // changed markers of a window, staged for the epsilon update. Packed in 16-byte fields (three stores per entry)
    uint4 a[kChgCap];           // x,y: address of the slice block; z,w: q1 = fixed-point deltaBeta*mstd (genotype 1; genotype 2: twice that)
    uint4 b[kChgCap];           // x,y: qm = fixed-point mave*deltaBeta*mstd (missing genotype); z: words (0: nothing to apply); w: first word of class 2 (0xFFFFFFFF: BED)
    uint4 c[kChgCap];           // x: first word of class "missing"; y: != 0: block in an inbox in the LL form with this tag (address counts 16-byte units);
                                // z: exclusive prefix of the words of the sparse entries (BED entries count 0)
    uint32_t total, n_sparse, n_bed, any;   // (16-byte aligned: written with one store)
    long long delta_sum;        // what the chunk's sparse entries add to the slice sum (from the genotype counts of their directory entries)
    long long bed_delta;        // same for its BED entries (counted while they are applied)
    long long qm_sum;           // sum of mave*mstd*deltaBeta on the grid 2^-kOffShift: the base term -mave*mstd*deltaBeta of every
                                // individual (:265-267) is kept as ONE scalar per launch
    unsigned long long live_mask;   // bit x: entry x is a sparse block with something to apply
    uint32_t n_changed;             // entries with deltaBeta != 0 (whatever their slice blocks hold): "the window changed something"
    uint32_t pad0_, pad1_, pad2_;
};
static_assert(sizeof(ChgTab) % 16 == 0, "ChgTab");

// Fixed-point deltas of one changed marker: q1 (genotype 1), qm (missing), its share dl of the slice sum, its base term qo
// on the finer grid of the launch's scalar.
__device__ __forceinline__ void entry_deltas(const Blk &b, double dbs, double mave, double q_scale, long long &q1, long long &qm,
                                             long long &dl, long long &qo) {
    q1 = __double2ll_rn(dbs * q_scale);
    qm = __double2ll_rn(mave * dbs * q_scale);
    const double base = mave * dbs;
    // what the entry adds to the slice sum (exact): n1*q1 + n2*2*q1 + nm*qm; BED blocks are counted while they are applied
    dl = (b.b1 == 0xFFFFFFFFu) ? 0ll : ((long long)b.n1 + 2ll * (long long)b.n2) * q1 + (long long)b.nm * qm;
    qo = (fabs(base) < 512.0) ? __double2ll_rn(base * 4503599627370496.0 /* 2^kOffShift */) : (long long)(1ull << 62);   // absurd value: caught by the caller
}
__device__ __forceinline__ void store_entry(ChgTab *chg, uint32_t x, const Blk &b, uint32_t ll, long long q1, long long qm, uint32_t nw_eff, uint32_t cum) {
    const uint64_t pa = (uint64_t)(uintptr_t)b.ptr;
    chg->a[x] = make_uint4((uint32_t)pa, (uint32_t)(pa >> 32), (uint32_t)(u64)q1, (uint32_t)((u64)q1 >> 32));
    chg->b[x] = make_uint4((uint32_t)(u64)qm, (uint32_t)((u64)qm >> 32), nw_eff, b.b1);
    chg->c[x] = make_uint4(b.b2, ll, cum, 0u);
}
// Warps 0 and 1 stage a chunk of <= 64 entries (one entry per thread): deltas, prefix of the block lengths, counters.
// Warp 1's prefix needs warp 0's total, which it computes itself from the lengths (both warps see all 64 lengths through
// one shared array written before a named barrier of the two warps).
__device__ __forceinline__ void stage_chunk(ChgTab *chg, uint32_t *scr /*[64]*/, uint32_t nx, const Blk &b, double dbs, double mave, uint32_t ll,
                                            double q_scale, uint32_t tid /* < 64 */) {
    const uint32_t lane = tid & 31u, wrp = tid >> 5;
    const bool two = nx > 32u;   // short chunks (the usual case) are staged by the first warp alone: no exchange, no barrier
    if (!two && wrp) return;
    const bool valid = tid < nx;
    long long q1 = 0, qm = 0, dl = 0, qo = 0;
    if (valid) entry_deltas(b, dbs, mave, q_scale, q1, qm, dl, qo);
    const uint32_t nwe = (valid && dbs != 0.0) ? b.nw : 0u;
    const bool live = nwe != 0u, bed = live && b.b1 == 0xFFFFFFFFu;
    const uint32_t a = (live && !bed) ? nwe : 0u;
    uint32_t s0 = a;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t0 = __shfl_up_sync(0xffffffffu, s0, o);
        if (lane >= (uint32_t)o) s0 += t0;
    }
    const uint32_t tot = __shfl_sync(0xffffffffu, s0, 31);
    dl = (long long)warp_sum_u64((u64)dl);
    qo = (long long)warp_sum_u64((u64)qo);
    const uint32_t msp = __ballot_sync(0xffffffffu, live && !bed);
    const uint32_t nsp = __popc(msp), nbd = __popc(__ballot_sync(0xffffffffu, bed));
    const uint32_t nch = __popc(__ballot_sync(0xffffffffu, valid && dbs != 0.0));
    if (!two) {
        if (valid) store_entry(chg, tid, b, ll, q1, qm, nwe, s0 - a);
        if (tid == 0) {
            *reinterpret_cast<uint4 *>(&chg->total) = make_uint4(tot, nsp, nbd, nch ? 1u : 0u);
            chg->delta_sum = dl; chg->qm_sum = qo; chg->bed_delta = 0ll; chg->live_mask = (unsigned long long)msp;
            chg->n_changed = nch;
        }
        return;
    }
    if (lane == 0) {   // this warp's totals for the other one
        scr[16 + wrp] = nch;
        scr[wrp * 8 + 0] = tot; scr[wrp * 8 + 1] = nsp; scr[wrp * 8 + 2] = nbd; scr[wrp * 8 + 3] = msp;
        *reinterpret_cast<long long *>(scr + wrp * 8 + 4) = dl; *reinterpret_cast<long long *>(scr + wrp * 8 + 6) = qo;
    }
    named_barrier(2, 64);
    const uint32_t base = wrp ? scr[0] : 0u;
    if (valid) store_entry(chg, tid, b, ll, q1, qm, nwe, base + s0 - a);
    if (tid == 0) {
        *reinterpret_cast<uint4 *>(&chg->total) = make_uint4(scr[0] + scr[8], scr[1] + scr[9], scr[2] + scr[10], (scr[16] + scr[17]) ? 1u : 0u);
        chg->n_changed = scr[16] + scr[17];
        chg->delta_sum = *reinterpret_cast<long long *>(scr + 4) + *reinterpret_cast<long long *>(scr + 12);
        chg->qm_sum = *reinterpret_cast<long long *>(scr + 6) + *reinterpret_cast<long long *>(scr + 14);
        chg->bed_delta = 0ll;
        chg->live_mask = (unsigned long long)scr[3] | ((unsigned long long)scr[11] << 32);
    }
}
// Unit mode (one entry per thread, any thread): the entry, its sums through shared-memory atomics; finish_chunk adds the prefix.
__device__ __forceinline__ void stage_entry_any(ChgTab *chg, uint32_t x, const Blk &b, double dbs, double mave, double q_scale) {
    long long q1, qm, dl, qo;
    entry_deltas(b, dbs, mave, q_scale, q1, qm, dl, qo);
    store_entry(chg, x, b, 0u, q1, qm, (dbs != 0.0) ? b.nw : 0u, 0u);
    atomicAdd(reinterpret_cast<unsigned long long *>(&chg->delta_sum), (unsigned long long)dl);
    atomicAdd(reinterpret_cast<unsigned long long *>(&chg->qm_sum), (unsigned long long)qo);
}
__device__ __forceinline__ void finish_chunk(ChgTab *chg, uint32_t nx, uint32_t lane) {   // warp 0, after a CTA barrier
    const uint4 b0 = (lane < nx) ? chg->b[lane] : make_uint4(0u, 0u, 0u, 0u), b1 = (lane + 32u < nx) ? chg->b[lane + 32u] : make_uint4(0u, 0u, 0u, 0u);
    const bool v0 = b0.z != 0u, v1 = b1.z != 0u, bed0 = v0 && b0.w == 0xFFFFFFFFu, bed1 = v1 && b1.w == 0xFFFFFFFFu;
    const uint32_t a0 = (v0 && !bed0) ? b0.z : 0u, a1 = (v1 && !bed1) ? b1.z : 0u;
    uint32_t s0 = a0, s1 = a1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t0 = __shfl_up_sync(0xffffffffu, s0, o), t1 = __shfl_up_sync(0xffffffffu, s1, o);
        if (lane >= (uint32_t)o) { s0 += t0; s1 += t1; }
    }
    const uint32_t tot0 = __shfl_sync(0xffffffffu, s0, 31), tot1 = __shfl_sync(0xffffffffu, s1, 31);
    if (lane < nx) chg->c[lane].z = s0 - a0;
    if (lane + 32u < nx) chg->c[lane + 32u].z = tot0 + s1 - a1;
    const uint32_t m0 = __ballot_sync(0xffffffffu, v0 && !bed0), m1 = __ballot_sync(0xffffffffu, v1 && !bed1);
    const uint32_t nsp = __popc(m0) + __popc(m1);
    const uint32_t nbd = __popc(__ballot_sync(0xffffffffu, bed0)) + __popc(__ballot_sync(0xffffffffu, bed1));
    if (lane == 0) {
        chg->bed_delta = 0ll; chg->total = tot0 + tot1; chg->n_sparse = nsp; chg->n_bed = nbd; chg->any = (nsp + nbd) ? 1u : 0u;
        chg->n_changed = nsp + nbd;
        chg->live_mask = (unsigned long long)m0 | ((unsigned long long)m1 << 32);
    }
}
// last entry x < nx whose prefix is <= f
__device__ __forceinline__ uint32_t find_chunk_entry(const ChgTab *chg, uint32_t nx, uint32_t f) {
    uint32_t lo = 0, hi = nx;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (chg->c[mid].z <= f) lo = mid; else hi = mid;
    }
    return lo;
}
// sparse word: four u16 indices, all different inside one marker; unused lanes of a word point at the dummy slot L (skipped)
__device__ __forceinline__ void apply_word(uint64_t x, long long q, u64 *__restrict__ Eq, uint32_t L) {
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    const uint32_t i0 = lo & 0xFFFFu, i1 = lo >> 16, i2 = hi & 0xFFFFu, i3 = hi >> 16;
    const u64 v0 = Eq[i0], v1 = Eq[i1], v2 = Eq[i2], v3 = Eq[i3];   // independent loads (the indices differ, or are the dummy slot)
    if (i0 != L) Eq[i0] = v0 + (u64)q;
    if (i1 != L) Eq[i1] = v1 + (u64)q;
    if (i2 != L) Eq[i2] = v2 + (u64)q;
    if (i3 != L) Eq[i3] = v3 + (u64)q;
}
// BED word w: individuals 32w .. 32w+31; returns what was added (a BED block has no genotype counts in a directory)
__device__ __forceinline__ long long apply_bed_word(uint64_t bits, uint32_t w, long long q1, long long q2, long long qm, u64 *__restrict__ Eq) {
    long long added = 0;
    if (bits == ~0ull) return added;
    const uint32_t r2 = g_bed_norot ? 0u : (threadIdx.x & 15u) * 2u;   // every lane starts at its own bank (see dot_bed_word)
#pragma unroll
    for (uint32_t h = 0; h < 2; h++) {  // only the non-zero genotypes are visited (codes as in dot_bed_word)
        const uint32_t v = (uint32_t)(bits >> (32u * h));
        const uint32_t b0 = v & 0x55555555u, b1 = (v >> 1) & 0x55555555u;
        uint32_t a = rotr32((~b0 & 0x55555555u) | (b0 & ~b1), r2);  // genotype 1, 2 or missing
        const uint32_t two = rotr32(~b0 & ~b1 & 0x55555555u, r2), miss = rotr32(b0 & ~b1, r2);
        u64 *eh = Eq + 32u * w + 16u * h;
        while (a) {
            const uint32_t pos = __ffs((int)a) - 1u;
            a &= a - 1u;
            const long long d = ((miss >> pos) & 1u) ? qm : (((two >> pos) & 1u) ? q2 : q1);
            eh[((pos + r2) & 31u) >> 1] += (u64)d;
            added += d;
        }
    }
    return added;
}

// Epsilon update with the nx entries staged in chg (CTA barrier passed). All threads of the CTA call it.
// The slice is fixed point: the order in which the entries are added does not matter, and the slice sum follows from the
// genotype counts (no reduction). The sparse words of all entries are flattened and loaded together (`ll` is the same for
// the whole chunk: two separate load paths, because a value that is selected between two loads is moved behind the join
// and every load would be waited for before the next one is issued); they are then applied entry by entry, one CTA
// barrier per entry, so that two threads never update the same individual at once. BED blocks follow, one at a time.
// Returns (in every thread) what was added to the slice sum. Ends with a CTA barrier.
__device__ __forceinline__ long long apply_chunk(ChgTab *chg, uint32_t nx, bool ll_chunk, u64 *__restrict__ Eq, uint32_t L,
                                                 unsigned long long *nnz_upd, uint32_t *err) {
    const uint32_t tid = threadIdx.x;
    const uint32_t total = chg->total, nbed = chg->n_bed;
    const unsigned long long live = chg->live_mask;
    long long added = chg->delta_sum;
    bool synced = true;   // nothing written since the last barrier
    for (uint32_t f0 = 0; f0 < total; f0 += kApplyQ * kThreads) {
        uint64_t wd[kApplyQ];
        long long qv[kApplyQ];
        uint32_t we[kApplyQ];
        uint32_t xlo = 0xFFFFFFFFu, xhi = 0u;
#pragma unroll
        for (uint32_t i = 0; i < kApplyQ; i++) {
            const uint32_t f = f0 + tid + i * kThreads;
            we[i] = 0xFFFFFFFFu; wd[i] = 0ull; qv[i] = 0;
            if (f < total) {
                const uint32_t x = find_chunk_entry(chg, nx, f);
                const uint4 ea = chg->a[x], eb = chg->b[x], ec = chg->c[x];
                const uint32_t w = f - ec.z;
                const uint64_t pa = ((uint64_t)ea.y << 32) | ea.x;
                we[i] = x; xlo = min(xlo, x); xhi = max(xhi, x);
                const long long q1 = (long long)(((u64)ea.w << 32) | ea.z);
                qv[i] = (w < eb.w) ? q1 : ((w < ec.x) ? 2ll * q1 : (long long)(((u64)eb.y << 32) | eb.x));
                if (!ll_chunk) wd[i] = __ldcg(reinterpret_cast<const unsigned long long *>(pa) + w);
                else wd[i] = ld_ll(reinterpret_cast<const uint4 *>(pa) + w, ec.y, err);
            }
        }
        // entries touched by this round: [x_first, x_last]; trailing entries without flattened words are skipped
        const uint32_t fend = min(total, f0 + kApplyQ * kThreads);
        const uint32_t x_first = find_chunk_entry(chg, nx, f0), x_last = find_chunk_entry(chg, nx, fend - 1u);
        // the sparse entries of [x_first, x_last] that have something to apply (the others, and BED blocks, are skipped)
        unsigned long long todo = live & (~0ull << x_first) & (~0ull >> (63u - x_last));
        for (; todo; todo &= todo - 1ull) {
            const uint32_t x = (uint32_t)__ffsll((long long)todo) - 1u;
            if (!synced) __syncthreads();
            if (x >= xlo && x <= xhi) {
#pragma unroll
                for (uint32_t i = 0; i < kApplyQ; i++)
                    if (we[i] == x) apply_word(wd[i], qv[i], Eq, L);
            }
            synced = false;
        }
    }
    if (nbed) {
        long long dl = 0;
        for (uint32_t x = 0; x < nx; x++) {
            const uint4 eb = chg->b[x];
            if (eb.w != 0xFFFFFFFFu || eb.z == 0u) continue;
            if (!synced) __syncthreads();   // the previous entry (or the sparse pass) is complete
            synced = false;
            const uint4 ea = chg->a[x], ec = chg->c[x];
            const uint32_t nwb = eb.z, ll = ec.y;
            const uint64_t pa = ((uint64_t)ea.y << 32) | ea.x;
            const long long q1 = (long long)(((u64)ea.w << 32) | ea.z), q2 = 2ll * q1, qm = (long long)(((u64)eb.y << 32) | eb.x);
            if (!ll_chunk) {
                for (uint32_t w = tid; w < nwb; w += blockDim.x) dl += apply_bed_word(__ldcg(reinterpret_cast<const unsigned long long *>(pa) + w), w, q1, q2, qm, Eq);
            } else {
                for (uint32_t w = tid; w < nwb; w += blockDim.x) dl += apply_bed_word(ld_ll(reinterpret_cast<const uint4 *>(pa) + w, ll, err), w, q1, q2, qm, Eq);
            }
        }
        dl = (long long)warp_sum_u64((u64)dl);
        if ((tid & 31u) == 0u && dl != 0) atomicAdd(reinterpret_cast<unsigned long long *>(&chg->bed_delta), (unsigned long long)dl);
        __syncthreads();
        added += chg->bed_delta;
        synced = false;   // (chg was read after the barrier: one more before it is restaged)
    }
    if (tid == 0 && nnz_upd) *nnz_upd += 4ull * total;   // (total = words of the chunk's sparse entries)
    __syncthreads();   // the slice is complete; chg may be restaged
    return added;
}

// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) k_brr_iteration(const BrrParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *Eq = reinterpret_cast<u64 *>(smem_raw);                        // [L+1] fixed point, slot L = dummy for PAD (always 0)
    ItemTab *tabs = reinterpret_cast<ItemTab *>(smem_raw + (((size_t)P.L + 2) * 8 + 15) / 16 * 16);
    ChgTab *chg = reinterpret_cast<ChgTab *>(tabs + 2);
    __shared__ double red[32];
    __shared__ double hyp_s[4 * kHypSmem];
    __shared__ uint4 udesc[kUnitCap];                   // work units of the dot phase
    __shared__ unsigned long long cnt_s[10];             // traffic counters: per warp 0..3 {non-zeros read by the dot, BED blocks}, [8] update
    __shared__ uint32_t chg_n;
    __shared__ unsigned long long spec_cnt[2];
    __shared__ uint32_t cnt_step[2 * kSpecMax];         // windows run ahead: non-zeros read / BED blocks by step of the window
    __shared__ uint32_t chg_base[33];
    __shared__ __align__(16) uint32_t psort[4 * kUnitCap];   // scratch of the exchange (unit offsets of the entries)
    u64 *upart = reinterpret_cast<u64 *>(psort);        // unit partials of the dot phase: {a12, am} per unit
    __shared__ uint32_t pcnt[kMaxRanks];
    __shared__ uint32_t abort_s;                        // != 0: the exchange failed somewhere, leave the window loop
    __shared__ __align__(8) uint32_t stg_scr[20];       // the two staging warps' totals; [18]: first changed step of a window run ahead

    const uint32_t S = P.S, L = P.L, R = P.R;
    const uint32_t c = blockIdx.x % S, r = blockIdx.x / S;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nctas = gridDim.x;

    if (tid == 0) { tabs[0].tag_valid = 0; tabs[1].tag_valid = 0; abort_s = 0u; }
    HypTabs H{P.logPi, P.chalf, P.denom, P.sdk};
    {
        const uint32_t gk = P.G * P.K;
        if (gk <= kHypSmem) {
            for (uint32_t i = tid; i < gk; i += blockDim.x) {
                hyp_s[i] = P.logPi[i]; hyp_s[kHypSmem + i] = P.chalf[i];
                hyp_s[2 * kHypSmem + i] = P.denom[i]; hyp_s[3 * kHypSmem + i] = P.sdk[i];
            }
            H = HypTabs{hyp_s, hyp_s + kHypSmem, hyp_s + 2 * kHypSmem, hyp_s + 3 * kHypSmem};
        }
    }
    // ---- load the slice: fixed point on the launch's grid (the base terms of the previous launch are folded in) ----
    long long slice_sum_q;   // sum of the slice's values (exact), the same in every thread
    {
        u64 v = 0;
        bool bad = false;
        for (uint32_t i = tid; i < L; i += blockDim.x) {
            const uint32_t gi = c * L + i;
            const long long q = (gi < P.N) ? __double2ll_rn((P.E_in[gi] + P.shift_in) * P.q_scale) : 0ll;
            bad |= (q > (1ll << 61)) || (q < -(1ll << 61));
            Eq[i] = (u64)q;
            v += (u64)q;
        }
        if (tid == 0) Eq[L] = 0ull;
        if (bad) atomicExch(P.pc.err, 5u);
        slice_sum_q = (long long)block_sum_u64(v, red);
    }

    uint32_t bar_target = 0;
    __shared__ long long tph[16];            // phase cycle counters of thread 0 (shared memory: keeps 32 registers free); 8..15: developer detail
    __shared__ unsigned long long gts[16];
    if (tid < 16) tph[tid] = 0;
    if (tid < 16) gts[tid] = 0;
    long long tclk = clock64(), tsub = tclk;
#define HB_PHASE(i) do { if (tid == 0) { long long t_ = clock64(); tph[i] += t_ - tclk; tclk = t_; \
        if (P.cta_cycles && win == 10) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); gts[i] = g_; } } } while (0)
#define HB_SUB(i) do { if (tid == 0) { long long t_ = clock64(); tph[i] += t_ - tsub; tsub = t_; } } while (0)
#define HB_STAMP(i, cond) do { if (P.cta_cycles && win == 10 && (cond)) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); gts[i] = g_; } } while (0)
    long long off_q = 0;     // base terms of this launch, fixed point (every thread keeps the same value)
    uint32_t j0 = 0, since = 0, win = 0;
    uint32_t n_sync = 0;
    if (tid < 10) cnt_s[tid] = 0;
    if (tid < 2) spec_cnt[tid] = 0;
    if (tid < 2 * kSpecMax) cnt_step[tid] = 0;
    const uint64_t padw = (uint64_t)L * 0x0001000100010001ull;
    const uint32_t SR = (P.SR == 0) ? 1u : P.SR;
    // A window run ahead takes NA steps: sync_rate of them, or more where a step holds few markers (one task with sync rate 1:
    // every marker is a step of its own, and a window of one marker pays the whole latency chain)
    const uint32_t NA = (SR > kSpecMax) ? 1u : min(kSpecMax, max(SR, P.n_ahead));

    while (j0 < P.lmax) {
        // Windows run ahead. After `sync_rate` steps without a change the reference synchronises after every single step
        // (:2044-2050), i.e. until the next change the steps read an epsilon that does not move. Such steps are taken `sync_rate`
        // at a time here: if none of them changes a marker the result is the same; otherwise the steps up to and including
        // the first one with a change are exactly what the reference computes -- their changed markers are applied -- and
        // the later steps are discarded and repeated by the next window (effects, components and Acum are simply written
        // again; the component counts are taken from comp[] after the loop; the draws are counter-based). With several GPUs
        // the first changed step is taken over the lists of all GPUs (every GPU sees every list), before anything is applied.
        const bool spec_win = (P.mode == MODE_CHAIN && NA > 1u && since >= SR && !(P.flags & 2u));
        const uint32_t Tdiv = (P.pc.nranks > 1) ? P.pc.T_total : P.T;   // list entries carry the global window position
        const uint32_t n = (P.mode != MODE_CHAIN) ? (P.lmax - j0)
                                                  : ((since >= SR) ? (spec_win ? min(NA, P.lmax - j0) : 1u) : min(SR - since, P.lmax - j0));
        uint32_t s_star = 0xFFFFFFFFu;   // first step of a window run ahead that changed a marker
        const uint32_t W = n * P.T, base = j0 * P.T;
        const uint32_t dbuf = win & 1u;
        const uint32_t n_items = (W > r) ? (W - r + R - 1) / R : 0;

        if (P.mode != MODE_SCAADD) {
            for (uint32_t k0 = 0; k0 < n_items; k0 += kTabCap) {
                const uint32_t nk = min((uint32_t)kTabCap, n_items - k0);
                ItemTab *tab = &tabs[win & 1u];
                // ---- 1. item table (normally prebuilt during the previous window) -------
                if (!(tab->tag_valid && tab->tag_base == base && tab->tag_W == W && tab->tag_k0 == k0)) {
                    __syncthreads();
                    if (P.meta) { stage_items(tab, P, r, c, base, W, k0, tid, blockDim.x); cp_async_wait_all(); }
                    else fill_items_sync(tab, P, r, c, base, W, k0, tid, blockDim.x);
                    __syncthreads();
                    finish_table(tab, udesc, P, r, c, base, W, k0, 0, blockDim.x, false);
                }
                __syncthreads();
                HB_PHASE(0);
                // the next window's marker data start their way into the other table now (speculating that this window
                // ends with a synchronisation: next window = SR steps); they land during the dot phase
                const bool stage_next = (P.mode == MODE_CHAIN && k0 + kTabCap >= n_items && j0 + n < P.lmax);
                const uint32_t n_next = min(spec_win ? NA : SR, P.lmax - (j0 + n));   // the likely next window: no change -> the regime stays
                if (stage_next) stage_items(&tabs[(win + 1u) & 1u], P, r, c, (j0 + n) * P.T, n_next * P.T, 0, tid, blockDim.x);
                // ---- 2. dot: warp w takes the units w, w + 16, ... two at a time (their gathers and the two warp reductions
                //         interleave: twice the independent work per warp), while the words of the next two units are in flight
                //         (four register sets in rotation, no copies)
                {
                    const uint32_t nun = tab->nunits;
                    const uint4 none = make_uint4(0u, 0u, 0u, 0u);
                    uint32_t u = warp;
                    uint4 da, db, dc, dd;
                    uint64_t xa[4], xb[4], xc[4], xd[4];
#define HB_FETCH(D, X, UU) do { D = ((UU) < nun) ? udesc[UU] : none; load_unit(D, X, lane, padw); } while (0)
#define HB_COMPUTE2(D0, X0, D1, X1) do { \
                        u64 a0_ = 0, a1_ = 0, m0_ = 0, m1_ = 0; \
                        if ((D0.z >> 16) != 0xFFFFu && (D1.z >> 16) != 0xFFFFu && (D0.z & 0xFFFFu) <= 128u && (D1.z & 0xFFFFu) <= 128u) { \
                            dot_words(X0, 0u, D0.z & 0xFFFFu, D0.z >> 16, D0.w & 0xFFFFu, Eq, lane, a0_, m0_); \
                            dot_words(X1, 0u, D1.z & 0xFFFFu, D1.z >> 16, D1.w & 0xFFFFu, Eq, lane, a1_, m1_); \
                        } else { \
                            dot_unit(D0, X0, Eq, lane, padw, a0_, m0_); \
                            dot_unit(D1, X1, Eq, lane, padw, a1_, m1_); \
                        } \
                        _Pragma("unroll") for (int o_ = 16; o_ > 0; o_ >>= 1) { \
                            const u64 t0_ = __shfl_xor_sync(0xffffffffu, a0_, o_), t1_ = __shfl_xor_sync(0xffffffffu, a1_, o_); \
                            a0_ += t0_; a1_ += t1_; \
                        } \
                        /* missing genotypes are rare: most units have no word of that class (warp-uniform test) */ \
                        const bool hm0_ = (D0.z >> 16) == 0xFFFFu || (D0.w & 0xFFFFu) < (D0.z & 0xFFFFu); \
                        const bool hm1_ = (D1.z >> 16) == 0xFFFFu || (D1.w & 0xFFFFu) < (D1.z & 0xFFFFu); \
                        if (hm0_) m0_ = warp_sum_u64(m0_); \
                        if (hm1_) m1_ = warp_sum_u64(m1_); \
                        if (lane == 0) { \
                            upart[2u * u] = a0_; upart[2u * u + 1u] = hm0_ ? m0_ : 0ull; \
                            if (u + kWarps < nun) { upart[2u * (u + kWarps)] = a1_; upart[2u * (u + kWarps) + 1u] = hm1_ ? m1_ : 0ull; } \
                        } \
                        u += 2 * kWarps; } while (0)
                    HB_FETCH(da, xa, u);
                    HB_FETCH(db, xb, u + kWarps);
                    while (u < nun) {
                        HB_FETCH(dc, xc, u + 2 * kWarps);
                        HB_FETCH(dd, xd, u + 3 * kWarps);
                        HB_COMPUTE2(da, xa, db, xb);
                        if (u >= nun) break;
                        HB_FETCH(da, xa, u + 2 * kWarps);
                        HB_FETCH(db, xb, u + 3 * kWarps);
                        HB_COMPUTE2(dc, xc, dd, xd);
                    }
#undef HB_FETCH
#undef HB_COMPUTE2
                }
                cp_async_wait_all();
                __syncthreads();
                HB_PHASE(1);
                // ---- 3. publish the slice partials (data + tag in one 16-byte store, no fence) ----
                const uint32_t tag = win + 1u;
                if (tid < nk) {
                    const uint32_t k = tid, p = r + R * (k0 + k);
                    if (tab->meta[k].m >= 0) {   // (a padded task step contributes nothing, :2029-2034)
                        // partial of num/mstd: sum_1 + 2 sum_2 + mave*sum_M - mave*sum_slice   (:327-339)
                        u64 s12 = 0, sm = 0;
#pragma unroll 1
                        for (uint32_t u = tab->ucum[k], u1 = tab->ucum[k + 1]; u < u1; u++) { s12 += upart[2u * u]; sm += upart[2u * u + 1u]; }
                        const double d12 = (double)(long long)s12 * P.q_inv, dm = (double)(long long)(sm - (u64)slice_sum_q) * P.q_inv;
                        st_slot(P.slots + (size_t)p * S + c, fma(tab->meta[k].mave, dm, d12), tag);
                    }
                }
                HB_STAMP(11, tid == 0);
                // ---- 4. draws: item kk of the group is drawn by slice-CTA kk % S, one warp per item;
                //         meanwhile the upper warps prepare the next window's table
                if (warp < kDrawWarps) {
                    const uint32_t kf = (k0 == 0u) ? c : (c + S - (k0 % S)) % S;  // first table entry owned by this CTA
                    for (uint32_t k = kf + warp * S; k < nk; k += kDrawWarps * S)
                        if (tab->meta[k].m >= 0) draw_marker_warp(P, tab, k, H, r + R * (k0 + k), base + r + R * (k0 + k), dbuf, win % 3u, tag, lane, (warp == 0) ? &tph[13] : nullptr,
                                                                     (P.cta_cycles && win == 10 && warp == 0) ? &gts[12] : nullptr);
                    HB_STAMP(8 + (warp & 1u), lane == 0 && warp < 2);
                } else if (tid < kDrawWarps * 32 + kTabCap) {  // traffic counters, by the table warps (off the draws' path)
                    const uint32_t t = tid - kDrawWarps * 32, w = t >> 5;
                    uint32_t v = 0, vb = 0;
                    if (t < nk && tab->meta[t].m >= 0) {
                        if (tab->b1[t] == 0xFFFFFFFFu) vb = 1u; else v = 4u * tab->nw[t];
                    }
                    if (spec_win) {   // by step: only the steps that are kept count (window end)
                        const uint32_t st = (r + R * (k0 + t)) / P.T;
                        if (v) atomicAdd(&cnt_step[st], v);
                        if (vb) atomicAdd(&cnt_step[kSpecMax + st], vb);
                    } else {
                        v = __reduce_add_sync(0xffffffffu, v);
                        vb = __reduce_add_sync(0xffffffffu, vb);
                        if (lane == 0) { cnt_s[2 * w] += v; cnt_s[2 * w + 1] += vb; }
                    }
                }
                if (warp >= kDrawWarps && stage_next) {
                    const uint32_t j1 = j0 + n, n1 = n_next;
                    finish_table(&tabs[(win + 1u) & 1u], udesc, P, r, c, j1 * P.T, n1 * P.T, 0, kDrawWarps * 32, blockDim.x - kDrawWarps * 32, !(P.flags & 1u));
                    HB_STAMP(10, tid == kDrawWarps * 32);
                }
                __syncthreads();  // table reuse
                HB_PHASE(2);
            }
            if (P.mode == MODE_DOT) break;  // single window, nothing to apply
            // ---- 4. the one grid barrier of the window --------------------------------
            __syncthreads();
            bar_target += nctas;
            if (tid == 0) red_release_add_u32(P.bar, 1u);  // arrive (release: the CTA's results, ordered before by bar.sync)
            if (tid == 0) {
                // A CTA that gave up on the exchange (dead peer, capacity) never arrives here again: everybody else finds
                // the sticky error word while spinning and leaves the window loop as well (no hang on a partial failure).
                uint32_t spin = 0;
                while (ld_acquire_u32(P.bar) < bar_target) {
                    if (P.pc.nranks > 1 && (++spin & 0x3FFu) == 0u && *reinterpret_cast<volatile uint32_t *>(P.pc.err) != 0u) { abort_s = 1u; break; }
                }
            }
            __syncthreads();
            if (abort_s) break;
            HB_PHASE(3);
        }

        // ---- 5. apply the window's non-zero deltaBetas (in any order: the slice is fixed point) ---------------
        bool any = false;
        const size_t dslot = (size_t)dbuf * P.Wmax;
        if (P.mode == MODE_CHAIN) {
            const uint32_t buf3 = win % 3u, par = win & 1u, NR = P.pc.nranks, me = P.pc.rank;
            const unsigned long long seq_w = P.pc.seq_base + win + 1ull;  // tag of this window's data in the inboxes
            const ChgEnt *llist = P.chg_list + (size_t)buf3 * P.Wmax;
            const uint4 *ldir = P.chg_dir + (size_t)buf3 * P.Wmax * S;   // [entry][slice] directory entries of the changed markers
            if (tid == 0) tsub = clock64();
            // the first 64 entries and their directory entries are requested together with the count (one L2 round trip)
            ChgEnt spec;
            spec.p = 0xFFFFFFFFu; spec.m = 0; spec.dbs = 0.0; spec.mave = 0.0; spec.rec = 0;
            uint4 sdir = make_uint4(0u, 0u, 0u, 0u);
            uint32_t soff = 0;                        // multi-GPU: first unit of the marker's record in the peers' inboxes
            // (multi-GPU: warps 2 and 3 read the same for the exchange, so that warps 0 and 1 go straight to the local update)
            const uint32_t ts = (NR > 1 && tid >= kChgCap) ? tid - kChgCap : tid;
            if (tid < kChgCap || (NR > 1 && tid < 2 * kChgCap)) {
                spec = ld_chg_ent(llist + ts); sdir = __ldcg(ldir + (size_t)ts * S + c);
                if (NR > 1) soff = __ldcg(P.chg_off + (size_t)buf3 * P.Wmax + ts);
            }
            const unsigned long long cntv = __ldcg(P.chg_cnt + buf3);
            const uint32_t nloc = (uint32_t)cntv;
            if (blockIdx.x == 0 && tid == 0) P.chg_cnt[(win + 2u) % 3u] = 0ull;  // free since the previous grid barrier
            bool ok = true, fast_push = false;
            uint32_t push_f = 0;
            unsigned long long push_v = 0ull;
            if (NR > 1) {
                // ---- 5a. this GPU's changed markers for every peer, over NVLink in the LL form (tagged 16-byte stores on
                // peer-mapped pointers, no fence, no arrival counter): the count (one 8-byte store, tagged), the list
                // entries (block 0) and the genotype records, whose units all CTAs share evenly -- one SM alone pushes
                // only ~6 GB/s. The records lie in the inbox in list order: entry i starts at unit off[i] (reserved with
                // the list slot by the same atomic), so the unit space is contiguous.
                const unsigned long long seq = seq_w;
                const size_t region = ((size_t)par * NR + me) * P.pc.inbox_stride;
                const uint32_t units = (uint32_t)(cntv >> 32);
                ok = nloc <= kMaxMerged && kInboxHeader + (size_t)units * 16 <= P.pc.inbox_stride;
                if (!ok && blockIdx.x == 0 && tid == 0) atomicExch(P.pc.err, 1u);
                if (blockIdx.x == 0 && tid < NR && tid != me)
                    *reinterpret_cast<volatile unsigned long long *>(P.pc.inbox_peer[tid] + region) =
                        (unsigned long long)(ok ? nloc : 0xFFFFFFFFu) | (seq << 32);
                // The slice directory entries of the list's first kDirCap markers travel with it: the CTA of slice c in replica
                // group 0 ships those of its slice (no dependent read of the record's directory on the other side).
                auto push_dirs = [&](uint32_t t) {
                    if (ok && r == 0 && t < min(nloc, kDirCap)) {
                        uint64_t da, db;
                        if (spec.rec & 1ull) { da = (uint64_t)(soff + c * (L / 32)) | (0xFFFFFFFFull << 32); db = 0ull; }
                        else { da = (uint64_t)(soff + dir_bytes(S) / 8u + sdir.x) | ((uint64_t)sdir.y << 32); db = (uint64_t)sdir.z; }
                        for (uint32_t h = 0; h < NR; h++) {
                            if (h == me) continue;
                            uint4 *d = reinterpret_cast<uint4 *>(P.pc.inbox_peer[h] + region + kDirBase) + ((size_t)t * S + c) * 2;
                            st_ll(d, da, (uint32_t)seq); st_ll(d + 1, db, (uint32_t)seq);
                        }
                    }
                };
                constexpr uint32_t kPushThreads = kThreads - 2 * 32;   // warps 2..15 move the records
                fast_push = ok && nloc > 0 && nloc <= kChgCap && units <= nctas * kPushThreads;
                if (fast_push) {
                    // The usual case: warps 0 and 1 go straight to the local update; warps 2 and 3 (which hold the same
                    // speculative read) build the scratch, ship entries and directory entries; warps 2..15 move the
                    // records, at most one unit per thread, whose load is in flight during the staging of the local chunk
                    // (stores: see 5b).
                    if (warp >= 2) {
                        uint32_t *sp = psort;
                        unsigned long long *srec = reinterpret_cast<unsigned long long *>(psort + ((nloc + 2u) & ~1u));
                        const uint32_t t2 = tid - 2 * 32;
                        if (t2 < nloc) { sp[t2] = soff; srec[t2] = spec.rec; }
                        if (t2 == 0) sp[nloc] = units;
                        named_barrier(3, kPushThreads);
                        push_f = blockIdx.x * kPushThreads + t2;
                        if (push_f < units) {
                            const uint32_t e = find_entry(sp, nloc, push_f);
                            push_v = __ldg(reinterpret_cast<const unsigned long long *>(srec[e] & ~15ull) + (push_f - sp[e]));
                        }
                        if (blockIdx.x == 0 && t2 < nloc) {
                            const uint64_t e0 = (uint64_t)spec.p | ((uint64_t)(uint32_t)seq << 32);
                            const uint64_t e3 = (spec.rec & 1ull) | ((unsigned long long)soff << 4);
                            for (uint32_t h = 0; h < NR; h++) {
                                if (h == me) continue;
                                uint4 *d = reinterpret_cast<uint4 *>(P.pc.inbox_peer[h] + region + 16 + (size_t)t2 * kLLEntry);
                                st_ll(d, e0, (uint32_t)seq); st_ll(d + 1, (uint64_t)__double_as_longlong(spec.dbs), (uint32_t)seq);
                                st_ll(d + 2, (uint64_t)__double_as_longlong(spec.mave), (uint32_t)seq); st_ll(d + 3, e3, (uint32_t)seq);
                            }
                        }
                        if (warp < 4) push_dirs(t2);
                    }
                } else
                if (ok && nloc > 0) {   // long lists / heavy records: all threads
                    if (tid < kChgCap) push_dirs(tid);
                    // scratch: first unit of every entry [nloc + 1], then the record addresses [nloc]; very long lists
                    // are searched in global memory instead
                    constexpr uint32_t kScr = 4 * kUnitCap;
                    const bool sp_fit = nloc + 2u <= kScr;
                    uint32_t *sp = psort;
                    unsigned long long *srec = reinterpret_cast<unsigned long long *>(psort + ((nloc + 2u) & ~1u));
                    const bool recs_fit = ((nloc + 2u) & ~1u) + 2u * nloc <= kScr;
                    const uint32_t *goff = P.chg_off + (size_t)buf3 * P.Wmax;
                    for (uint32_t i = tid; i < nloc; i += blockDim.x) {
                        const ChgEnt en = ld_chg_ent(llist + i);
                        const uint32_t o = __ldcg(goff + i);
                        if (sp_fit) sp[i] = o;
                        if (recs_fit) srec[i] = en.rec;
                        if (blockIdx.x == 0) {  // the list entries first: the peers stage them while the records travel
                            const uint64_t e0 = (uint64_t)en.p | ((uint64_t)(uint32_t)seq << 32);
                            const uint64_t e3 = (en.rec & 1ull) | ((unsigned long long)o << 4);  // first LL unit in the payload | BED flag
                            for (uint32_t h = 0; h < NR; h++) {
                                if (h == me) continue;
                                uint4 *d = reinterpret_cast<uint4 *>(P.pc.inbox_peer[h] + region + 16 + (size_t)i * kLLEntry);
                                st_ll(d, e0, (uint32_t)seq); st_ll(d + 1, (uint64_t)__double_as_longlong(en.dbs), (uint32_t)seq);
                                st_ll(d + 2, (uint64_t)__double_as_longlong(en.mave), (uint32_t)seq); st_ll(d + 3, e3, (uint32_t)seq);
                            }
                        }
                    }
                    if (tid == 0 && sp_fit) sp[nloc] = units;
                    __syncthreads();
                    for (uint32_t f = blockIdx.x * blockDim.x + tid; f < units; f += nctas * blockDim.x) {
                        uint32_t e, e_first;
                        if (sp_fit) { e = find_entry(sp, nloc, f); e_first = sp[e]; }
                        else {  // offsets ascend with the list index (one atomic reserves both)
                            uint32_t lo = 0, hi = nloc;
                            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldcg(goff + mid) <= f) lo = mid; else hi = mid; }
                            e = lo; e_first = __ldcg(goff + lo);
                        }
                        const unsigned long long ra = recs_fit ? srec[e] : __ldcg(&llist[e].rec);
                        const uint64_t v = __ldg(reinterpret_cast<const unsigned long long *>(ra & ~15ull) + (f - e_first));
                        for (uint32_t h = 0; h < NR; h++)
                            if (h != me) st_ll(reinterpret_cast<uint4 *>(P.pc.inbox_peer[h] + region + kInboxHeader) + f, v, (uint32_t)seq);
                    }
                    __syncthreads();  // the scratch is free again
                }
            }
            // 5c as a function: it runs before the local update in a window run ahead, after it otherwise
            const unsigned long long seq = seq_w;
            auto wait_counts = [&]() -> bool {
                // ---- 5c. wait for every peer's count of this window (it follows the peer's own grid barrier); entries and
                //          records are validated unit by unit when they are read
                if (tid < NR) {
                    uint32_t nh = ok ? nloc : 0xFFFFFFFFu;
                    if (tid != me) {
                        const long long t0 = clock64();
                        bool late = false;
                        const volatile unsigned long long *hdr =
                            reinterpret_cast<const volatile unsigned long long *>(P.pc.inbox_local + ((size_t)par * NR + tid) * P.pc.inbox_stride);
                        unsigned long long hv = 0ull;
                        uint32_t spin = 0;
                        while (!late && (uint32_t)(hv >> 32) != (uint32_t)seq) {
                            hv = *hdr;
                            if ((uint32_t)(hv >> 32) == (uint32_t)seq) break;
                            if (clock64() - t0 > P.pc.timeout_cycles) late = true;
                            if ((++spin & 0xFFu) == 0u && *reinterpret_cast<volatile uint32_t *>(P.pc.err) != 0u) late = true;  // somebody already gave up
                        }
                        nh = late ? 0xFFFFFFFFu : (uint32_t)hv;
                        if (late) atomicExch(P.pc.err, 2u);
                    }
                    pcnt[tid] = nh;
                }
                __syncthreads();
                HB_SUB(12);
                bool bad = false;
                for (uint32_t h = 0; h < NR; h++) bad |= (pcnt[h] == 0xFFFFFFFFu);
                if (bad && blockIdx.x == 0 && tid == 0) atomicExch(P.pc.err, 3u);  // give up: the host reports the failure after the launch
                return bad;
            };
            bool counts_done = false;
            if (spec_win && NR > 1) {
                // window run ahead, several GPUs: the first step with a change over ALL lists, before anything is applied
                if (wait_counts()) break;
                counts_done = true;
                uint32_t v = 0xFFFFFFFFu;
                for (uint32_t i = tid; i < nloc; i += blockDim.x) {
                    const ChgEnt en = ld_chg_ent(llist + i);
                    if (en.dbs != 0.0) v = min(v, en.p / Tdiv);
                }
                uint32_t n_peers = 0;
                for (uint32_t hh = 1; hh < NR; hh++) n_peers += pcnt[(me + hh) % NR];
                for (uint32_t y = tid; y < n_peers; y += blockDim.x) {
                    uint32_t x = y, h = me;
                    for (uint32_t hh = 1; hh < NR; hh++) {
                        h = (me + hh) % NR;
                        const uint32_t n_h = pcnt[h];
                        if (x < n_h) break;
                        x -= n_h;
                    }
                    const uint4 *le = reinterpret_cast<const uint4 *>(P.pc.inbox_local + ((size_t)par * NR + h) * P.pc.inbox_stride + 16 + (size_t)x * kLLEntry);
                    const uint64_t e0 = ld_ll(le, (uint32_t)seq, P.pc.err), e1 = ld_ll(le + 1, (uint32_t)seq, P.pc.err);
                    if (__longlong_as_double((long long)e1) != 0.0) v = min(v, (uint32_t)e0 / Tdiv);
                }
                v = __reduce_min_sync(0xffffffffu, v);
                if (lane == 0) chg_base[warp] = v;
                __syncthreads();
                v = 0xFFFFFFFFu;
                for (uint32_t w = 0; w < (blockDim.x >> 5); w++) v = min(v, chg_base[w]);
                __syncthreads();
                s_star = v;
            }
            HB_SUB(11);
            // ---- 5b. this GPU's own changed markers (the peers' are still on their way). Warps 0 and 1 stage a chunk (one entry
            //          per thread); the slice directory entry travels with the list entry (written by the drawing warp), and
            //          the first 64 entries were requested together with the count: one L2 round trip up to here.
            if (ok) {
                if (spec_win && NR == 1 && nloc > 0) {   // first step with a change (entries of later steps are dropped below)
                    if (nloc <= kChgCap) {
                        if (tid < kChgCap) {
                            uint32_t v = (tid < nloc && spec.dbs != 0.0) ? spec.p / Tdiv : 0xFFFFFFFFu;
                            v = __reduce_min_sync(0xffffffffu, v);
                            if (lane == 0) stg_scr[18 + warp] = v;   // (read by everybody after the staging barrier)
                            if (nloc > 32u) { named_barrier(2, 64); v = min(stg_scr[18], stg_scr[19]); }
                            s_star = v;
                        }
                    } else {
                        uint32_t v = 0xFFFFFFFFu;
                        for (uint32_t i = tid; i < nloc; i += blockDim.x) {
                            const ChgEnt en = ld_chg_ent(llist + i);
                            if (en.dbs != 0.0) v = min(v, en.p / Tdiv);
                        }
                        v = __reduce_min_sync(0xffffffffu, v);
                        if (lane == 0) chg_base[warp] = v;
                        __syncthreads();
                        v = 0xFFFFFFFFu;
                        for (uint32_t w = 0; w < (blockDim.x >> 5); w++) v = min(v, chg_base[w]);
                        __syncthreads();
                        s_star = v;
                    }
                }
                for (uint32_t x0 = 0; x0 < nloc; x0 += kChgCap) {
                    const uint32_t nx = min((uint32_t)kChgCap, nloc - x0);
                    if (tid < kChgCap) {
                        Blk bk;
                        bk.ptr = nullptr; bk.nw = 0; bk.b1 = 0; bk.b2 = 0; bk.n1 = 0; bk.n2 = 0; bk.nm = 0;
                        double dbs = 0.0, mv = 0.0;
                        if (tid < nx) {
                            const ChgEnt en = (x0 == 0) ? spec : ld_chg_ent(llist + x0 + tid);
                            const uint4 dv = (x0 == 0) ? sdir : __ldcg(ldir + (size_t)(x0 + tid) * S + c);
                            bk = block_from_dir(en.rec, dv, c, S, L);
                            dbs = en.dbs; mv = en.mave;
                            if (spec_win && en.p / Tdiv != s_star) dbs = 0.0;   // a later step of a window run ahead: repeated
                        }
                        stage_chunk(chg, stg_scr, nx, bk, dbs, mv, 0u, P.q_scale, tid);
                    }
                    if (x0 == 0 && fast_push && warp >= 2 && push_f < (uint32_t)(cntv >> 32)) {
                        for (uint32_t h = 0; h < NR; h++)
                            if (h != me)
                                st_ll(reinterpret_cast<uint4 *>(P.pc.inbox_peer[h] + ((size_t)par * NR + me) * P.pc.inbox_stride + kInboxHeader) + push_f, push_v,
                                      (uint32_t)seq_w);
                    }
                    __syncthreads();
                    HB_SUB(9);
                    any |= chg->any != 0u;
                    if (blockIdx.x == 0 && tid == 0) cnt_s[9] += chg->n_changed;
                    off_q -= chg->qm_sum;
                    slice_sum_q += apply_chunk(chg, nx, false, Eq, L, (r == 0) ? &cnt_s[8] : nullptr, P.pc.err);
                    HB_SUB(10);
                }
                if (spec_win && NR == 1 && nloc > 0 && nloc <= kChgCap) s_star = min(stg_scr[18], stg_scr[19]);
            }
            HB_PHASE(4);
            if (NR > 1) {
                // ---- 5c. the peers' counts (unless the window ran ahead: done above)
                if (!counts_done && wait_counts()) break;
                // ---- 5d. the peers' changed markers, straight from the inboxes (no merge by position: any order gives the
                //          same slice). The entries of ALL peers fill common chunks: the fixed cost of a chunk (entry ->
                //          directory -> words, three dependent reads, plus the staging barriers) is paid once per 64 markers,
                //          not once per peer.
                uint32_t n_peers = 0;
                for (uint32_t hh = 1; hh < NR; hh++) n_peers += pcnt[(me + hh) % NR];
                for (uint32_t x0 = 0; x0 < n_peers; x0 += kChgCap) {
                    const uint32_t nx = min((uint32_t)kChgCap, n_peers - x0);
                    if (tid < kChgCap) {
                        Blk bk;
                        bk.ptr = nullptr; bk.nw = 0; bk.b1 = 0; bk.b2 = 0; bk.n1 = 0; bk.n2 = 0; bk.nm = 0;
                        double dbs = 0.0, mv = 0.0;
                        if (tid < nx) {
                            uint32_t x = x0 + tid, h = me;
                            for (uint32_t hh = 1; hh < NR; hh++) {  // source GPU and index in its list
                                h = (me + hh) % NR;
                                const uint32_t n_h = pcnt[h];
                                if (x < n_h) break;
                                x -= n_h;
                            }
                            const unsigned char *reg = P.pc.inbox_local + ((size_t)par * NR + h) * P.pc.inbox_stride;
                            const uint4 *le = reinterpret_cast<const uint4 *>(reg + 16 + (size_t)x * kLLEntry);
                            uint64_t e1, e2;
                            if (x < kDirCap) {   // entry and directory units in one round trip
                                const uint4 *ld = reinterpret_cast<const uint4 *>(reg + kDirBase) + ((size_t)x * S + c) * 2;
                                uint64_t da, db;
                                ld_ll4(le + 1, le + 2, ld, ld + 1, (uint32_t)seq, P.pc.err, e1, e2, da, db);
                                bk = block_from_ll_dir(reinterpret_cast<const uint4 *>(reg + kInboxHeader), da, db, L);
                            } else {
                                uint64_t rr;
                                ld_ll3(le + 1, (uint32_t)seq, P.pc.err, e1, e2, rr);
                                bk = decode_block_ll(reinterpret_cast<const uint4 *>(reg + kInboxHeader) + (rr >> 4), (rr & 1ull) != 0,
                                                     c, S, L, (uint32_t)seq, P.pc.err);
                            }
                            dbs = __longlong_as_double((long long)e1);
                            mv = __longlong_as_double((long long)e2);
                            if (spec_win && (uint32_t)ld_ll(le, (uint32_t)seq, P.pc.err) / Tdiv != s_star) dbs = 0.0;   // repeated step
                        }
                        stage_chunk(chg, stg_scr, nx, bk, dbs, mv, (uint32_t)seq, P.q_scale, tid);
                    }
                    __syncthreads();
                    any |= chg->any != 0u;
                    off_q -= chg->qm_sum;
                    slice_sum_q += apply_chunk(chg, nx, true, Eq, L, (r == 0) ? &cnt_s[8] : nullptr, P.pc.err);
                }
            }
            if (off_q > (1ll << 61) || off_q < -(1ll << 61)) atomicExch(P.pc.err, 5u);
            HB_PHASE(7);
        } else if (P.mode == MODE_SCAADD) {  // unit mode: the caller's dense deltaBeta array, one entry per position
            for (uint32_t p0 = 0; p0 < W; p0 += blockDim.x) {
                const uint32_t p = p0 + tid;
                const double d = (p < W) ? __ldcg(P.dB + dslot + p) : 0.0;
                const bool ch = (d != 0.0);
                // every changed position decodes its own record: one dependent-load chain for all of them
                Blk b;
                b.ptr = nullptr; b.nw = 0; b.b1 = 0; b.b2 = 0; b.n1 = 0; b.n2 = 0; b.nm = 0;
                double mv = 0.0;
                if (ch) {
                    const int32_t m = P.order[base + p];
                    mv = P.mave[m];
                    b = decode_block(P.rec[m], c, S, L);
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, ch);
                if (lane == 0) chg_base[warp] = __popc(bal);
                __syncthreads();
                if (tid == 0) {
                    uint32_t a = 0;
                    for (uint32_t w = 0; w < (blockDim.x >> 5); w++) { uint32_t t = chg_base[w]; chg_base[w] = a; a += t; }
                    chg_n = a;
                }
                __syncthreads();
                const uint32_t nchg = chg_n;
                const uint32_t slot = chg_base[warp] + __popc(bal & ((1u << lane) - 1u));
                for (uint32_t x0 = 0; x0 < nchg; x0 += kChgCap) {
                    const uint32_t nx = min((uint32_t)kChgCap, nchg - x0);
                    if (tid == 0) { chg->delta_sum = 0ll; chg->qm_sum = 0ll; }
                    __syncthreads();
                    if (ch && slot >= x0 && slot < x0 + nx) stage_entry_any(chg, slot - x0, b, d, mv, P.q_scale);
                    __syncthreads();
                    if (warp == 0) finish_chunk(chg, nx, lane);
                    __syncthreads();
                    any |= chg->any != 0u;
                    off_q -= chg->qm_sum;
                    slice_sum_q += apply_chunk(chg, nx, false, Eq, L, (r == 0) ? &cnt_s[8] : nullptr, P.pc.err);
                }
            }
            if (off_q > (1ll << 61) || off_q < -(1ll << 61)) atomicExch(P.pc.err, 5u);
            HB_PHASE(4);
        }
        uint32_t n_done = n;
        if (spec_win) {
            if (s_star != 0xFFFFFFFFu) n_done = s_star + 1u;   // (then `any` is set)
            if (blockIdx.x == 0 && tid == 0) { spec_cnt[0]++; spec_cnt[1] += (unsigned long long)(n - n_done) * P.T; }
            if (tid == 0) {   // traffic counters: the steps that are kept (a barrier of the update lies behind the last add)
                unsigned long long a = 0, b = 0;
                for (uint32_t st = 0; st < n; st++) {
                    if (st < n_done) { a += cnt_step[st]; b += cnt_step[kSpecMax + st]; }
                    cnt_step[st] = 0; cnt_step[kSpecMax + st] = 0;
                }
                cnt_s[0] += a; cnt_s[1] += b;
            }
        }
        if (any) {
            since = 0;
            n_sync++;
        } else {
            since += n_done;
        }
        j0 += n_done;
        win++;
        HB_PHASE(5);
    }

    // ---- epilogue: group 0 stores the slices (back in double) and their sums ------------------
    if (P.mode != MODE_DOT) {
        if (r == 0) {
            double v = 0.0, v2 = 0.0, va = 0.0, vm = 0.0;
            for (uint32_t i = tid; i < L; i += blockDim.x) {
                const uint32_t gi = c * L + i;
                const long long qe = (long long)Eq[i];
                if (qe > (1ll << 62) || qe < -(1ll << 62)) atomicExch(P.pc.err, 5u);   // left the range the grid was chosen for
                const double e = (double)qe * P.q_inv;
                P.E_out[gi] = e;
                if (gi < P.N) { v += e; v2 += e * e; va += fabs(e); vm = fmax(vm, fabs(e)); }
            }
            const double s1 = block_sum(v, red);
            const double s2 = block_sum(v2, red);
            const double sa = block_sum(va, red);
            const double sx = block_max(vm, red);
            if (tid == 0) { P.slice_sum_out[c] = s1; P.slice_sq_out[c] = s2; P.slice_abs_out[c] = sa; P.slice_max_out[c] = sx; }
        }
        if (blockIdx.x == 0 && tid == 0) {
            *P.off_out = ldexp((double)off_q, -kOffShift);
            P.stats[0] = n_sync;
            P.stats[1] = win;
            for (int i = 0; i < 16; i++) P.stats[8 + i] = (unsigned long long)tph[i];
        }
    }
    if (P.cta_cycles && tid == 0)
        for (int i = 0; i < 16; i++) P.cta_cycles[(size_t)blockIdx.x * 16 + i] = gts[i];
    // traffic counters (lane 0 of every warp holds a share)
    __syncthreads();
    if (tid == 0) {
        const unsigned long long nd = cnt_s[0] + cnt_s[2] + cnt_s[4] + cnt_s[6], nb = cnt_s[1] + cnt_s[3] + cnt_s[5] + cnt_s[7];
        if (nd) atomicAdd(&P.stats[2], nd);
        if (nb && c == 0) atomicAdd(&P.stats[4], nb);
        if (r == 0 && cnt_s[8]) atomicAdd(&P.stats[3], cnt_s[8]);
        if (blockIdx.x == 0) { atomicAdd(&P.stats[5], cnt_s[9]); P.stats[6] = spec_cnt[0]; P.stats[7] = spec_cnt[1]; }   // changed markers; windows run ahead, draws repeated
    }
}

// ---------------------------------------------------------------------------------
// Window-ordered inputs of one iteration. One block per step j: position q = j*T + slot holds the marker of ONE local
// task at that step. perm holds the task-local marker order of every local task block (concatenated);
// ut/zt != NULL: positional tape values indexed like perm; else RNG spec v1 (Philox), keyed by the task, not by the slot.
// wts != NULL (chain mode, T <= kBalanceMaxT): the tasks of a step are laid out heaviest marker first (ties: lower task
// first). The marker loop deals the positions of a window round-robin to its CTA groups, so every group gets an equal
// share of the window's non-zeros; the order of the positions inside a step has no other meaning than the (fixed) order
// in which the epsilon updates of a window are summed.
constexpr uint32_t kBalanceMaxT = 1024;
__global__ void __launch_bounds__(256) k_window_order(const int32_t *__restrict__ perm, const double *__restrict__ ut,
                               const double *__restrict__ zt, const int32_t *__restrict__ task_len,
                               const int32_t *__restrict__ task_off, uint32_t T, uint32_t lmax, uint32_t seed,
                               uint32_t iteration, uint32_t task_first, int32_t *__restrict__ order,
                               double *__restrict__ u, double *__restrict__ z,
                               const uint64_t *__restrict__ rec, const double *__restrict__ mave, const double *__restrict__ mstd,
                               const double *__restrict__ beta, const int32_t *__restrict__ grp, const uint32_t *__restrict__ rec_bytes,
                               const uint32_t *__restrict__ wts,
                               uint32_t S, WinMeta *__restrict__ meta, uint4 *__restrict__ dirw) {
    __shared__ unsigned long long key[kBalanceMaxT];
    const uint32_t j = blockIdx.x;
    const size_t Q = (size_t)lmax * T;
    const bool balance = (wts != nullptr) && T >= 2 && T <= kBalanceMaxT;
    if (balance) {
        uint32_t n = 1;
        while (n < T) n <<= 1;
        for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) {
            unsigned long long k = 0ull;  // padding of the bitonic network sorts last
            if (t < T) {
                const unsigned long long w = ((int32_t)j < task_len[t]) ? (unsigned long long)wts[task_off[t] + perm[task_off[t] + j]] + 2ull : 1ull;
                k = (w << 16) | (unsigned long long)(0xFFFFu - t);
            }
            key[t] = k;
        }
        __syncthreads();
        for (uint32_t sz = 2; sz <= n; sz <<= 1) {
            for (uint32_t st = sz >> 1; st > 0; st >>= 1) {
                for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                    const uint32_t l = i ^ st;
                    if (l > i) {
                        const bool desc = ((i & sz) == 0);  // overall descending order
                        const unsigned long long a = key[i], b = key[l];
                        if ((a < b) == desc) { key[i] = b; key[l] = a; }
                    }
                }
                __syncthreads();
            }
        }
    }
    for (uint32_t slot = threadIdx.x; slot < T; slot += blockDim.x) {
        const uint32_t t = balance ? (0xFFFFu - (uint32_t)(key[slot] & 0xFFFFull)) : slot;
        const size_t q = (size_t)j * T + slot;
        WinMeta wm;
        wm.rec = 0; wm.mave = 0.0; wm.mstd = 0.0; wm.beta = 0.0; wm.u = 0.0; wm.z = 0.0; wm.m = -1; wm.grp = 0; wm.rec_bytes = 0; wm.pad = 0;
        if ((int32_t)j >= task_len[t]) {  // :2029-2034
            order[q] = -1; u[q] = 0.0; z[q] = 0.0;
        } else {
            const uint32_t o = (uint32_t)task_off[t] + j;
            const int32_t m = task_off[t] + perm[o];
            order[q] = m;
            if (ut) {
                wm.u = ut[o]; wm.z = zt[o];
            } else {
                uint32_t w[4];
                philox4x32(j, iteration, 0x48594452u, 0u, seed, task_first + t, w);
                wm.u = ((double)(w[0] >> 5) * 67108864.0 + (double)(w[1] >> 6)) * (1.0 / 9007199254740992.0);
                const double u1 = ((double)w[2] + 0.5) * (1.0 / 4294967296.0);
                const double u2 = ((double)w[3] + 0.5) * (1.0 / 4294967296.0);
                wm.z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
            }
            u[q] = wm.u; z[q] = wm.z;
            if (meta) { wm.m = m; wm.rec = rec[m]; wm.mave = mave[m]; wm.mstd = mstd[m]; wm.beta = beta[m]; wm.grp = grp[m]; wm.rec_bytes = rec_bytes[m]; }
        }
        if (!meta) continue;
        meta[q] = wm;
        if (wm.m >= 0 && !(wm.rec & 1ull)) {
            const uint32_t *dir = reinterpret_cast<const uint32_t *>(wm.rec);  // 12 bytes per slice, never written after staging
            uint32_t c = 0;
            for (; c + 4 <= S; c += 4) {  // 48 bytes = three 16-byte loads (records are 16-byte aligned)
                const uint4 a = __ldg(reinterpret_cast<const uint4 *>(dir + c * 3)), b = __ldg(reinterpret_cast<const uint4 *>(dir + c * 3) + 1),
                            d = __ldg(reinterpret_cast<const uint4 *>(dir + c * 3) + 2);
                dirw[(size_t)c * Q + q] = make_uint4(a.x, a.y, a.z, 0u);
                dirw[(size_t)(c + 1) * Q + q] = make_uint4(a.w, b.x, b.y, 0u);
                dirw[(size_t)(c + 2) * Q + q] = make_uint4(b.z, b.w, d.x, 0u);
                dirw[(size_t)(c + 3) * Q + q] = make_uint4(d.y, d.z, d.w, 0u);
            }
            for (; c < S; c++) dirw[(size_t)c * Q + q] = make_uint4(__ldg(dir + c * 3), __ldg(dir + c * 3 + 1), __ldg(dir + c * 3 + 2), 0u);
        }
    }
}

// sum of beta^2 per group over the local markers (src/BayesRRm.cpp:2496-2499): block (b, g) sums the markers of group g
// inside chunk b, a second small kernel adds the chunk sums in chunk order (fixed order: bit-reproducible)
constexpr uint32_t kSqChunks = 148;
__global__ void __launch_bounds__(256) k_beta_sqnorm(const double *__restrict__ beta, const int32_t *__restrict__ grp,
                                                     uint32_t M, double *__restrict__ part, const int32_t *__restrict__ comp,
                                                     const uint8_t *__restrict__ grp_active, uint32_t K, int32_t *__restrict__ cass) {
    __shared__ double red[32];
    __shared__ int32_t hist[32];
    const int g = blockIdx.y;
    const uint32_t per = (M + gridDim.x - 1) / gridDim.x, m0 = blockIdx.x * per, m1 = min(M, m0 + per);
    // component counts of the group (cass, :1904): taken from comp[] once the marker loop is over, every marker once
    const bool count = (cass != nullptr) && grp_active[g] && K <= 32u;
    if (threadIdx.x < 32) hist[threadIdx.x] = 0;
    __syncthreads();
    double v = 0.0;
    for (uint32_t m = m0 + threadIdx.x; m < m1; m += blockDim.x) {
        const double b = beta[m];
        if (grp[m] == g) {
            v += b * b;
            if (count) atomicAdd(&hist[comp[m]], 1);
        }
    }
    const double s = block_sum(v, red);
    if (threadIdx.x == 0) part[(size_t)g * gridDim.x + blockIdx.x] = s;
    if (count && threadIdx.x < K && hist[threadIdx.x]) atomicAdd(&cass[(size_t)g * K + threadIdx.x], hist[threadIdx.x]);
}
__global__ void k_beta_sqnorm_fin(const double *__restrict__ part, uint32_t nchunks, uint32_t G, double *__restrict__ out) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double s = 0.0;
    for (uint32_t b = 0; b < nchunks; b++) s += part[(size_t)g * nchunks + b];
    out[g] = s;
}

}  // namespace hb
