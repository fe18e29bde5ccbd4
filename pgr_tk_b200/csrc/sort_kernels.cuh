// sort_kernels.cuh — stable LSD radix sort (8-bit digits) of (key, payload index) pairs held in HBM, plus the
// run-length step that turns sorted tuples into the ShmmrFragMap CSR.  Hand-written for sm_100a; all passes are
// HBM-streaming (each pass reads the keys twice and writes them once).
//
// Stability is what makes the per-key FragmentSignature order equal the reference's insertion order
// (seq_db.rs:326-340, :605-612: sequences in sid order, pairs in position order).
#pragma once
#include "common.cuh"

namespace pgr {

constexpr int RS_NT = 256;            // threads per CTA (8 warps)
constexpr int RS_SEG = 8192;          // elements per warp segment; a warp walks its segment in order, 32 at a time
constexpr int RS_WARPS = RS_NT / 32;

// sort record: 128-bit key (k0 major, k1 minor) + index of the tuple in insertion order
struct SortKey { uint64_t k0, k1; };

__device__ __forceinline__ uint32_t digit_of(const SortKey &k, int pass) {
    // pass 0..6 -> bytes of k1 (low to high, 56 significant bits), pass 7..13 -> bytes of k0
    return pass < 7 ? (uint32_t)(k.k1 >> (8 * pass)) & 0xFFu : (uint32_t)(k.k0 >> (8 * (pass - 7))) & 0xFFu;
}

// per-segment digit histogram: hist[digit * n_seg + seg]
__global__ void __launch_bounds__(RS_NT) rs_hist_kernel(const SortKey *keys, uint64_t n, int pass, uint32_t *hist, uint32_t n_seg) {
    __shared__ uint32_t h[RS_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < 256; i += 32) h[warp][i] = 0;
    __syncwarp();
    const uint32_t seg = blockIdx.x * RS_WARPS + warp;
    if (seg < n_seg) {
        const uint64_t b = (uint64_t)seg * RS_SEG, e = min(n, b + RS_SEG);
        for (uint64_t i = b + lane; i < e; i += 32) atomicAdd(&h[warp][digit_of(keys[i], pass)], 1u);
        __syncwarp();
        for (int i = lane; i < 256; i += 32) hist[(uint64_t)i * n_seg + seg] = h[warp][i];
    }
}

// exclusive scan of hist in (digit-major, segment-minor) order, in place; also reports whether one digit holds every
// element (then the pass is a no-op and the scatter is skipped).  Single CTA: every thread sums one contiguous chunk, the
// 1024 chunk sums are scanned once (two barriers), then every thread rewrites its chunk.  (The first version scanned
// 1024 items per iteration with three barriers each: 150 us per call on 10^5 items, the largest share of a sort pass.)
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t *hist, uint64_t n_items, uint64_t n, uint32_t n_seg, uint32_t *skip) {
    __shared__ uint64_t wtot[32];
    __shared__ uint32_t all_one;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) all_one = 0;
    const uint64_t per = (n_items + 1023) / 1024;
    const uint64_t b = min(n_items, (uint64_t)threadIdx.x * per), e = min(n_items, b + per);
    uint64_t sum = 0;
    for (uint64_t i = b; i < e; i++) sum += hist[i];
    uint64_t incl = sum;
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint64_t v = wtot[lane];
        uint64_t wi = v;
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= d) wi += t;
        }
        wtot[lane] = wi - v;   // exclusive prefix of the warp totals
    }
    __syncthreads();
    uint64_t run = wtot[warp] + incl - sum;
    for (uint64_t i = b; i < e; i++) { const uint32_t v = hist[i]; hist[i] = (uint32_t)run; run += v; }
    __syncthreads();
    // all-one detection: digit d holds everything iff its row starts at 0 and the next row starts at n
    for (uint32_t d = threadIdx.x; d < 256; d += 1024) {
        const uint64_t start = hist[(uint64_t)d * n_seg];
        const uint64_t next = (d + 1 < 256) ? hist[(uint64_t)(d + 1) * n_seg] : n;
        if (start == 0 && next == n) all_one = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) *skip = all_one;
}

// stable scatter: each warp walks its segment in order; lanes with the same digit are ranked by lane id
__global__ void __launch_bounds__(RS_NT) rs_scatter_kernel(const SortKey *keys, const uint32_t *idx, uint64_t n, int pass,
                                                            const uint32_t *hist, uint32_t n_seg, SortKey *keys_out,
                                                            uint32_t *idx_out, const uint32_t *skip) {
    if (*skip) return;
    __shared__ uint32_t off[RS_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t seg = blockIdx.x * RS_WARPS + warp;
    if (seg >= n_seg) return;
    for (int i = lane; i < 256; i += 32) off[warp][i] = hist[(uint64_t)i * n_seg + seg];
    __syncwarp();
    const uint64_t b = (uint64_t)seg * RS_SEG, e = min(n, b + RS_SEG);
    for (uint64_t i0 = b; i0 < e; i0 += 32) {
        const uint64_t i = i0 + lane;
        const bool live = i < e;
        SortKey k = {0, 0};
        uint32_t id = 0, d = 0xFFFFFFFFu;
        if (live) { k = keys[i]; id = idx ? idx[i] : (uint32_t)i; d = digit_of(k, pass); }
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        if (live) {
            const uint32_t rank = __popc(peers & ((1u << lane) - 1));
            const uint32_t pos = off[warp][d] + rank;
            keys_out[pos] = k;
            idx_out[pos] = id;
        }
        __syncwarp();
        if (live && (peers & ((1u << lane) - 1)) == 0) off[warp][d] += __popc(peers);  // leader advances the bucket
        __syncwarp();
    }
}

}  // namespace pgr
