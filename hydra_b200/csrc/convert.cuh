// Genotype staging kernels: raw PLINK BED bytes / reference sparse lists -> sliced
// records in HBM, and back.  Integer work only; results are bit-exact with the
// reference's Data::sparse_data_fill_indices (src/data.cpp:1224-1290),
// sparse_data_correct_for_missing_phenotype (:1112-1158) and
// get_bed_marker_from_sparse (:826-865).
#pragma once
#include "common.cuh"

namespace hb {

// PLINK code -> class: 00 -> 2 copies, 10 -> 1 copy, 11 -> 0, 01 -> missing (class 3)
// (src/data.cpp:1248-1257, src/mk_lut.cpp:25-36)
__device__ __forceinline__ uint32_t code_to_class(uint32_t code) {
    // code: 0->2, 1->3, 2->1, 3->0   packed as 2-bit fields of 0b00011110
    return (0x1Eu >> (2u * code)) & 3u;
}
// class -> PLINK code (inverse map): 0->3, 1->2, 2->0, 3->1 : 0b01001011
__device__ __forceinline__ uint32_t class_to_code(uint32_t cls) { return (0x4Bu >> (2u * cls)) & 3u; }

// class of compacted individual i of a raw BED column (0 for i >= N)
__device__ __forceinline__ uint32_t raw_class(const uint8_t *__restrict__ row, const uint32_t *__restrict__ rmap,
                                              uint32_t i, uint32_t N) {
    if (i >= N) return 0u;
    uint32_t raw = rmap ? rmap[i] : i;
    uint32_t code = (row[raw >> 2] >> (2u * (raw & 3u))) & 3u;
    return code_to_class(code);
}

// ---------------------------------------------------------------------------
// Synthetic genotypes: cell (i,j) = word i%4 of
// philox(ctr=(i/4, j, attempt, 'GENO'), key=(seed,0)) compared with 3 integer
// thresholds per marker.  One thread = one BED byte.
// ---------------------------------------------------------------------------
__global__ void k_synth_bed(uint8_t *__restrict__ raw, size_t stride, uint32_t nb_raw, uint32_t n_ind_raw,
                            uint32_t seed, uint32_t j_global0, const uint32_t *__restrict__ thresholds,
                            const uint32_t *__restrict__ attempts) {
    const uint32_t m = blockIdx.y;
    const uint32_t t0 = thresholds[m * 3 + 0], t1 = thresholds[m * 3 + 1], t2 = thresholds[m * 3 + 2];
    const uint32_t att = attempts ? attempts[m] : 0u;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < nb_raw; q += gridDim.x * blockDim.x) {
        uint32_t w[4];
        philox4x32(q, j_global0 + m, att, 0x47454E4Fu, seed, 0u, w);
        uint32_t byte = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t i = q * 4 + k, code;
            if (i >= n_ind_raw) code = 0;
            else if (w[k] < t0) code = 1;
            else if (w[k] < t1) code = 0;
            else if (w[k] < t2) code = 2;
            else code = 3;
            byte |= code << (2 * k);
        }
        raw[(size_t)m * stride + q] = (uint8_t)byte;
    }
}

// ---------------------------------------------------------------------------
// Reference sparse lists -> raw BED bytes (rows pre-set to 0xFF):
// XOR 01 for ones, 11 for twos, 10 for missing (src/data.cpp:839-864).
// grid.y = marker, grid.x*block strides over the list entries; which: 0=I1 1=I2 2=IM.
// ---------------------------------------------------------------------------
__global__ void k_lists_to_bed(uint8_t *__restrict__ raw, size_t stride, const uint32_t *__restrict__ I,
                               const uint64_t *__restrict__ NS, const uint64_t *__restrict__ NL, uint32_t mask2) {
    const uint32_t m = blockIdx.y;
    const uint64_t s = NS[m], l = NL[m];
    uint32_t *row32 = reinterpret_cast<uint32_t *>(raw + (size_t)m * stride);  // stride is a multiple of 16
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < l; e += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t idx = I[s + e];
        atomicXor(&row32[idx >> 4], mask2 << (2u * (idx & 15u)));
    }
}

// ---------------------------------------------------------------------------
// Pass 1: per (marker, slice) class counts; per marker totals and record words.
// grid = markers, block = 256 (8 warps, one slice per warp at a time).
// cnt[(m*S+c)*3 + {0,1,2}] = n1,n2,nm of the slice;  start[m*S+c] = word offset
// of the slice block;  meta[m*4+{0..3}] = n1, n2, nm, total payload words.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_count(const uint8_t *__restrict__ raw, size_t stride,
                                               const uint32_t *__restrict__ rmap, uint32_t N, uint32_t S, uint32_t L,
                                               uint32_t *__restrict__ cnt, uint32_t *__restrict__ start,
                                               uint32_t *__restrict__ meta) {
    const uint32_t m = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint8_t *row = raw + (size_t)m * stride;
    for (uint32_t c = warp; c < S; c += nwarps) {
        uint32_t n1 = 0, n2 = 0, nm = 0;
        for (uint32_t it = 0; it < L; it += 32) {
            uint32_t cls = raw_class(row, rmap, c * L + it + lane, N);
            n1 += __popc(__ballot_sync(0xffffffffu, cls == 1));
            n2 += __popc(__ballot_sync(0xffffffffu, cls == 2));
            nm += __popc(__ballot_sync(0xffffffffu, cls == 3));
        }
        if (lane == 0) {
            uint32_t *o = cnt + ((size_t)m * S + c) * 3;
            o[0] = n1; o[1] = n2; o[2] = nm;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t1 = 0, t2 = 0, tm = 0, w = 0;
        for (uint32_t c = 0; c < S; c++) {
            const uint32_t *o = cnt + ((size_t)m * S + c) * 3;
            start[(size_t)m * S + c] = w;
            w += (o[0] + 3) / 4 + (o[1] + 3) / 4 + (o[2] + 3) / 4;
            t1 += o[0]; t2 += o[1]; tm += o[2];
        }
        meta[m * 4 + 0] = t1; meta[m * 4 + 1] = t2; meta[m * 4 + 2] = tm; meta[m * 4 + 3] = w;
    }
}

// ---------------------------------------------------------------------------
// Pass 2 (sparse records): directory + payload.  rec[m] is the record address
// with bit 0 = BED flag (BED markers are skipped here).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fill_sparse(const uint8_t *__restrict__ raw, size_t stride,
                                                     const uint32_t *__restrict__ rmap, uint32_t N, uint32_t S,
                                                     uint32_t L, const uint32_t *__restrict__ cnt,
                                                     const uint32_t *__restrict__ start,
                                                     const uint32_t *__restrict__ meta,
                                                     const uint64_t *__restrict__ rec) {
    const uint32_t m = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint64_t r = rec[m];
    if (r & 1ull) return;
    uint8_t *base = reinterpret_cast<uint8_t *>(r);
    uint32_t *dir = reinterpret_cast<uint32_t *>(base);
    uint64_t *payload = reinterpret_cast<uint64_t *>(base + dir_bytes(S));
    const uint8_t *row = raw + (size_t)m * stride;
    for (uint32_t c = threadIdx.x; c < S; c += blockDim.x) {
        const uint32_t *o = cnt + ((size_t)m * S + c) * 3;
        dir[c * 3 + 0] = start[(size_t)m * S + c];
        dir[c * 3 + 1] = o[0] | (o[1] << 16);
        dir[c * 3 + 2] = o[2];
    }
    const uint64_t pad = 0x0001000100010001ull * (uint64_t)L;
    const uint32_t words = meta[m * 4 + 3];
    for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) payload[w] = pad;
    __syncthreads();
    for (uint32_t c = warp; c < S; c += nwarps) {
        const uint32_t *o = cnt + ((size_t)m * S + c) * 3;
        uint16_t *blk = reinterpret_cast<uint16_t *>(payload + start[(size_t)m * S + c]);
        uint32_t off1 = 0, off2 = ((o[0] + 3) / 4) * 4, offm = off2 + ((o[1] + 3) / 4) * 4;
        const uint32_t lt = (1u << lane) - 1u;
        for (uint32_t it = 0; it < L; it += 32) {
            uint32_t cls = raw_class(row, rmap, c * L + it + lane, N);
            uint32_t b1 = __ballot_sync(0xffffffffu, cls == 1);
            uint32_t b2 = __ballot_sync(0xffffffffu, cls == 2);
            uint32_t bm = __ballot_sync(0xffffffffu, cls == 3);
            if (cls == 1) blk[off1 + __popc(b1 & lt)] = (uint16_t)(it + lane);
            else if (cls == 2) blk[off2 + __popc(b2 & lt)] = (uint16_t)(it + lane);
            else if (cls == 3) blk[offm + __popc(bm & lt)] = (uint16_t)(it + lane);
            off1 += __popc(b1); off2 += __popc(b2); offm += __popc(bm);
        }
    }
}

// ---------------------------------------------------------------------------
// Pass 3 (sparse records): bank-aware order inside every class run of every slice block.
// The dot product gathers epsilon (8-byte words, 16 distinct bank pairs) from shared memory with one 64-bit word
// of four indices per lane; the 16 lanes of a half-warp read, for index slot g, the elements at positions
// 64*chunk + 4*lane + g of the block (chunk = 16 consecutive words counted from the start of the slice block). The sum over a class does not depend on the order of its indices, so each
// run is permuted such that those 16 elements have distinct (index mod 16) wherever the residues allow it:
// elements are ranked by (depth inside their residue list, residue) and dealt to the half-warp groups in that order.
// Runs longer than kBankMax stay ascending. Exports go through the 2-bit form and are order independent.
// ---------------------------------------------------------------------------
constexpr uint32_t kBankMax = 4096;

__global__ void __launch_bounds__(256) k_bank_order(const uint64_t *__restrict__ rec, uint32_t S, uint32_t L) {
    extern __shared__ uint16_t bank_smem[];
    __shared__ uint32_t s_cnt[8][16];
    const uint32_t m = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint64_t r = rec[m];
    if (r & 1ull) return;
    uint8_t *base = reinterpret_cast<uint8_t *>(r);
    const uint32_t *dir = reinterpret_cast<const uint32_t *>(base);
    uint16_t *payload = reinterpret_cast<uint16_t *>(base + dir_bytes(S));
    uint16_t *in = bank_smem + (size_t)warp * 2 * kBankMax, *out = in + kBankMax;
    uint32_t *cnt = s_cnt[warp];
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t c = warp; c < S; c += nwarps) {
        const uint32_t st = dir[c * 3], n12 = dir[c * 3 + 1], nmiss = dir[c * 3 + 2];
        const uint32_t len[3] = {n12 & 0xFFFFu, n12 >> 16, nmiss};
        uint32_t off = 0;
        for (int cls = 0; cls < 3; cls++) {
            const uint32_t n = len[cls];
            uint16_t *run = payload + (size_t)st * 4 + off;
            const uint32_t ws = off / 4u;            // first word of the class inside the slice block
            const uint32_t nwc = (n + 3u) / 4u;      // its words
            off += nwc * 4u;
            if (n <= 16 || 4u * nwc > kBankMax) continue;
            if (lane < 16) cnt[lane] = 0;
            __syncwarp();
            // depth of every element inside its residue list (stable: runs are ascending)
            for (uint32_t i0 = 0; i0 < n; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool ok = i < n;
                const uint32_t idx = ok ? run[i] : 0u;
                const uint32_t res = ok ? (idx & 15u) : (16u + lane);  // inactive lanes match nobody
                const uint32_t peers = __match_any_sync(0xffffffffu, res);
                uint32_t depth = 0;
                if (ok) depth = cnt[res] + __popc(peers & lt);
                __syncwarp();
                if (ok && (peers >> lane) == 1u) cnt[res] += __popc(peers);  // highest lane of each group updates
                __syncwarp();
                if (ok) { in[i] = (uint16_t)idx; out[i] = (uint16_t)depth; }  // out temporarily holds the depth
            }
            __syncwarp();
            // Sequence index s = #elements with smaller (depth, residue): 16 consecutive elements of the sequence have
            // distinct residues wherever the residue lists are deep enough. The dot phase reads word w of the block with
            // lane w % 32, so one gather instruction of a half-warp takes the same lane q of 16 consecutive words
            // [16g, 16g + 16) of the block: the words of the class are cut at those boundaries and every segment of
            // `wseg` words is filled column by column (element s' of the segment -> word s' % wseg, lane s' / wseg),
            // the last, partly filled one included; PAD fills what is left of it.
            const uint32_t w0 = min(nwc, (16u - (ws & 15u)) & 15u);  // words before the first 16-word boundary
            uint32_t pos_of[kBankMax / 32];
            for (uint32_t t = 0, i = lane; i < n; i += 32, t++) {
                const uint32_t idx = in[i], res = idx & 15u, depth = out[i];
                uint32_t sq = 0;
#pragma unroll
                for (uint32_t rr = 0; rr < 16; rr++) {
                    const uint32_t cr = cnt[rr];
                    sq += min(cr, depth) + ((rr < res && cr > depth) ? 1u : 0u);
                }
                uint32_t segw, wseg, sp;  // first word of the segment (inside the class), its words, index inside it
                if (sq < 4u * w0) {
                    segw = 0; wseg = w0; sp = sq;
                } else {
                    const uint32_t s2 = sq - 4u * w0, g = s2 / 64u;
                    segw = w0 + 16u * g; wseg = min(16u, nwc - segw); sp = s2 - 64u * g;
                }
                pos_of[t] = 4u * (segw + sp % wseg) + sp / wseg;
            }
            __syncwarp();
            for (uint32_t i = lane; i < 4u * nwc; i += 32) out[i] = (uint16_t)L;  // PAD
            __syncwarp();
            for (uint32_t t = 0, i = lane; i < n; i += 32, t++) out[pos_of[t]] = in[i];
            __syncwarp();
            for (uint32_t i = lane; i < 4u * nwc; i += 32) run[i] = out[i];
            __syncwarp();
        }
    }
}

// Pass 2 (BED records): NA-compacted 2-bit codes, pad positions = 11. One thread = 16 individuals.
__global__ void k_fill_bed(const uint8_t *__restrict__ raw, size_t stride, const uint32_t *__restrict__ rmap,
                           uint32_t N, uint32_t S, uint32_t L, const uint64_t *__restrict__ rec) {
    const uint32_t m = blockIdx.y;
    const uint64_t r = rec[m];
    if (!(r & 1ull)) return;
    uint32_t *out = reinterpret_cast<uint32_t *>(r & ~15ull);
    const uint8_t *row = raw + (size_t)m * stride;
    const uint32_t nwords = S * L / 16;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += gridDim.x * blockDim.x) {
        uint32_t v = 0;
#pragma unroll
        for (uint32_t e = 0; e < 16; e++) {
            uint32_t cls = raw_class(row, rmap, w * 16 + e, N);
            v |= class_to_code(cls) << (2 * e);
        }
        out[w] = v;
    }
}

// ---------------------------------------------------------------------------
// Record -> NA-compacted BED bytes (S*L/4 bytes per marker, pad = 11).
// out rows must be pre-set to 0xFF for sparse records.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_record_to_bed(const uint64_t *__restrict__ rec, uint32_t m0, uint32_t S,
                                                       uint32_t L, uint8_t *__restrict__ out, size_t ostride) {
    const uint32_t mi = blockIdx.x, m = m0 + mi;
    const uint64_t r = rec[m];
    uint32_t *o32 = reinterpret_cast<uint32_t *>(out + (size_t)mi * ostride);
    if (r & 1ull) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(r & ~15ull);
        for (uint32_t w = threadIdx.x; w < S * L / 16; w += blockDim.x) o32[w] = src[w];
        return;
    }
    const uint8_t *base = reinterpret_cast<const uint8_t *>(r);
    const uint32_t *dir = reinterpret_cast<const uint32_t *>(base);
    const uint16_t *payload = reinterpret_cast<const uint16_t *>(base + dir_bytes(S));
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (uint32_t c = warp; c < S; c += nwarps) {
        const uint32_t st = dir[c * 3], n1 = dir[c * 3 + 1] & 0xFFFFu, n2 = dir[c * 3 + 1] >> 16, nm = dir[c * 3 + 2];
        const uint16_t *blk = payload + (size_t)st * 4;
        // every class is padded to whole words; PAD (= L) entries may sit anywhere inside a class (k_bank_order)
        const uint32_t o2 = ((n1 + 3) / 4) * 4, om = o2 + ((n2 + 3) / 4) * 4, oe = om + ((nm + 3) / 4) * 4;
        for (uint32_t e = lane; e < oe; e += 32) {
            const uint32_t idx = blk[e];
            if (idx == L) continue;
            const uint32_t i = c * L + idx, code = (e < o2) ? 1u : ((e < om) ? 3u : 2u);
            atomicXor(&o32[i >> 4], code << (2 * (i & 15)));
        }
    }
}

// ---------------------------------------------------------------------------
// NA-compacted BED bytes -> reference lists I1/I2/IM (ascending indices per
// marker): the device form of Data::sparse_data_fill_indices.  Output starts
// (relative to I1/I2/IM) are given per marker.  grid = markers, block = 256.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bed_to_lists(const uint8_t *__restrict__ bed, size_t stride, uint32_t N,
                                                      uint32_t *__restrict__ I1, const uint64_t *__restrict__ N1S,
                                                      uint32_t *__restrict__ I2, const uint64_t *__restrict__ N2S,
                                                      uint32_t *__restrict__ IM, const uint64_t *__restrict__ NMS) {
    __shared__ uint32_t wc[8][3];
    const uint32_t m = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint8_t *row = bed + (size_t)m * stride;
    const uint32_t chunk = ((N + 8 * 32 - 1) / (8 * 32)) * 32;  // individuals per warp, multiple of 32
    const uint32_t beg = warp * chunk, end = min(N, beg + chunk);
    uint32_t n1 = 0, n2 = 0, nm = 0;
    for (uint32_t it = beg; it < end; it += 32) {
        uint32_t cls = raw_class(row, nullptr, it + lane, N);
        n1 += __popc(__ballot_sync(0xffffffffu, cls == 1));
        n2 += __popc(__ballot_sync(0xffffffffu, cls == 2));
        nm += __popc(__ballot_sync(0xffffffffu, cls == 3));
    }
    if (lane == 0) { wc[warp][0] = n1; wc[warp][1] = n2; wc[warp][2] = nm; }
    __syncthreads();
    uint64_t o1 = N1S[m], o2 = N2S[m], om = NMS[m];
    for (uint32_t w = 0; w < warp; w++) { o1 += wc[w][0]; o2 += wc[w][1]; om += wc[w][2]; }
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t it = beg; it < end; it += 32) {
        uint32_t cls = raw_class(row, nullptr, it + lane, N);
        uint32_t b1 = __ballot_sync(0xffffffffu, cls == 1);
        uint32_t b2 = __ballot_sync(0xffffffffu, cls == 2);
        uint32_t bm = __ballot_sync(0xffffffffu, cls == 3);
        if (cls == 1) I1[o1 + __popc(b1 & lt)] = it + lane;
        else if (cls == 2) I2[o2 + __popc(b2 & lt)] = it + lane;
        else if (cls == 3) IM[om + __popc(bm & lt)] = it + lane;
        o1 += __popc(b1); o2 += __popc(b2); om += __popc(bm);
    }
}

}  // namespace hb
