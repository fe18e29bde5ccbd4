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

// ---- one-sweep form of the same sort ---------------------------------------------------------------------------------------
// The three-kernel pass above reads the keys twice and scatters 20-byte records lane by lane.  Here (a) ONE kernel reads the
// keys once and builds the digit histograms of ALL passes (so constant digits are known up front and their passes are not run
// at all), and (b) every pass is ONE kernel: a CTA takes the next tile (ticket), ranks its 2048 records by digit (warp-private
// counters + __match_any_sync, stable), learns where each digit's records go by a decoupled look-back over the tiles before it,
// stages the tile in digit order in shared memory and writes it out so that consecutive threads store consecutive records of
// a digit's run.  Per pass: 20 B read + 20 B written per record.
constexpr int OS_NT = 256, OS_WARPS = OS_NT / 32, OS_PER = 8, OS_TILE = OS_NT * OS_PER;   // 2048 records per tile
constexpr int OS_PASSES = 14;
constexpr uint32_t OS_FLAG_AGG = 1u << 30, OS_FLAG_PREFIX = 2u << 30, OS_VAL = (1u << 30) - 1;

// hist[pass * 256 + digit]: global counts of every digit of every pass in [first_pass, last_pass]
__global__ void __launch_bounds__(256) os_hist_kernel(const SortKey *keys, uint64_t n, int first_pass, int last_pass, uint32_t *hist) {
    __shared__ uint32_t h[OS_PASSES][256];
    for (int i = threadIdx.x; i < OS_PASSES * 256; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const SortKey k = keys[i];
        for (int p = first_pass; p <= last_pass; p++) atomicAdd(&h[p][digit_of(k, p)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < OS_PASSES * 256; i += blockDim.x) { const uint32_t v = (&h[0][0])[i]; if (v) atomicAdd(&hist[i], v); }
}
// per pass: exclusive scan of its 256 counts in place; skip[pass] = 1 when one digit holds every record
__global__ void __launch_bounds__(256) os_scan_kernel(uint32_t *hist, uint64_t n, uint32_t *skip) {
    __shared__ uint32_t ws[8];
    const int p = blockIdx.x, d = threadIdx.x, lane = d & 31, warp = d >> 5;
    const uint32_t v = hist[p * 256 + d];
    uint32_t incl = v;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    uint32_t wb = 0;
    for (int i = 0; i < warp; i++) wb += ws[i];
    hist[p * 256 + d] = wb + incl - v;
    const uint32_t one = __syncthreads_or(v == n && n > 0);
    if (d == 0) skip[p] = one ? 1u : 0u;
}

struct OsSmem {
    uint32_t wcnt[OS_WARPS][256];      // warp-private digit counts, then exclusive offsets of the warp inside the tile's digit run
    uint32_t tbase[256];               // first tile-local slot of each digit
    uint32_t gbase[256];               // first global slot of this tile's records of each digit
    uint32_t tile;
    alignas(16) SortKey skey[OS_TILE];
    uint32_t sidx[OS_TILE];
};

__global__ void __launch_bounds__(OS_NT) os_pass_kernel(const SortKey *keys, const uint32_t *idx, uint64_t n, int pass, const uint32_t *digit_start,
                                                         uint32_t *status /* [n_tiles][256] */, uint32_t *ticket, SortKey *keys_out, uint32_t *idx_out) {
    extern __shared__ __align__(16) unsigned char os_raw[];
    OsSmem &s = *reinterpret_cast<OsSmem *>(os_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s.tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < OS_WARPS * 256; i += OS_NT) (&s.wcnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s.tile;
    const uint64_t t0 = (uint64_t)tile * OS_TILE;
    // 1. load + rank inside the warp's 256 consecutive records (8 rounds of 32: record order = round-major, lane-minor)
    SortKey k[OS_PER]; uint32_t id[OS_PER], dg[OS_PER], lr[OS_PER];
#pragma unroll
    for (int r = 0; r < OS_PER; r++) {
        const uint64_t e = t0 + (uint64_t)warp * (OS_PER * 32) + r * 32 + lane;
        const bool live = e < n;
        k[r].k0 = 0; k[r].k1 = 0; id[r] = 0; dg[r] = 0xFFFFFFFFu;
        if (live) { k[r] = keys[e]; id[r] = idx ? idx[e] : (uint32_t)e; dg[r] = digit_of(k[r], pass); }
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, dg[r]);
        const uint32_t before = __popc(peers & ((1u << lane) - 1));
        lr[r] = 0;
        if (live) lr[r] = s.wcnt[warp][dg[r]] + before;
        __syncwarp();
        if (live && before == 0) s.wcnt[warp][dg[r]] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // 2. thread d owns digit d: tile count, per-warp offsets, look-back
    {
        const int d = tid;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < OS_WARPS; w++) { const uint32_t c = s.wcnt[w][d]; s.wcnt[w][d] = run; run += c; }
        const uint32_t cnt = run;
        // publish the tile's count, then accumulate the tiles before it (decoupled look-back)
        uint32_t *st = status + (size_t)tile * 256 + d;
        if (tile == 0) {
            atomicExch(st, cnt | OS_FLAG_PREFIX);
            s.gbase[d] = digit_start[d];
        } else {
            atomicExch(st, cnt | OS_FLAG_AGG);
            uint32_t excl = 0;
            for (int64_t t = (int64_t)tile - 1; t >= 0;) {
                const uint32_t v = *((volatile uint32_t *)(status + (size_t)t * 256 + d));
                if (v & OS_FLAG_PREFIX) { excl += v & OS_VAL; break; }
                if (v & OS_FLAG_AGG) { excl += v & OS_VAL; t--; }
            }
            atomicExch(st, (excl + cnt) | OS_FLAG_PREFIX);
            s.gbase[d] = digit_start[d] + excl;
        }
        // exclusive scan of the tile counts over the digits
        uint32_t incl = cnt;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
        __shared__ uint32_t ws[OS_WARPS];
        if (lane == 31) ws[warp] = incl;
        __syncthreads();
        uint32_t wb = 0;
        for (int i = 0; i < warp; i++) wb += ws[i];
        s.tbase[d] = wb + incl - cnt;
    }
    __syncthreads();
    // 3. stage in digit order
#pragma unroll
    for (int r = 0; r < OS_PER; r++) {
        if (dg[r] != 0xFFFFFFFFu) {
            const uint32_t slot = s.tbase[dg[r]] + s.wcnt[warp][dg[r]] + lr[r];
            s.skey[slot] = k[r]; s.sidx[slot] = id[r];
        }
    }
    __syncthreads();
    // 4. write out: slot i of the tile belongs to digit d(i); consecutive slots of a digit go to consecutive addresses
    const uint32_t n_tile = (uint32_t)min((uint64_t)OS_TILE, n - t0);
    for (uint32_t i = tid; i < n_tile; i += OS_NT) {
        const SortKey kk = s.skey[i];
        const uint32_t d = digit_of(kk, pass);
        const uint32_t pos = s.gbase[d] + (i - s.tbase[d]);
        keys_out[pos] = kk;
        idx_out[pos] = s.sidx[i];
    }
}

}  // namespace pgr
