// index_kernels.cuh — shimmer pairs -> (key, FragmentSignature) tuples -> ShmmrFragMap CSR; batched lookup.
#pragma once
#include "common.cuh"
#include "sort_kernels.cuh"

namespace pgr {

// one (shmmr-pair -> fragment) tuple, the unit exchanged between GPUs (40 bytes)
struct FragTuple {
    uint64_t h0, h1;                    // canonical key, h0 <= h1 (seq_db.rs:238-242, :391-395)
    uint32_t frg_id, sid, bgn, end;     // FragmentSignature (seq_db.rs:75)
    uint32_t ori;                       // 0/1
    uint32_t ord;                       // insertion ordinal of the sequence (multi-GPU merge, shard.cu); 0 elsewhere
};
static_assert(sizeof(FragTuple) == 40, "FragTuple layout");

struct PairParams {
    const pgr_mm128 *mm;         // final shimmers of the batch, flat
    const uint64_t *mm_off;      // [n_seq+1]
    uint32_t n_seq;
    const uint32_t *sid;         // [n_seq] caller's sequence ids
    const uint64_t *pair_off;    // [n_seq+1] exclusive scan of max(n_s - 1, 0)
    const uint32_t *frg_base;    // [n_seq] frg_id of the sequence's first pair
    FragTuple *out;              // [pair_off[n_seq]]
    uint64_t n_mm;
    uint32_t query_mode;         // 1: strict '<' canonicalisation (seq_db.rs:1213-1217), 0: '<=' (index build)
    uint32_t ord_base;           // FragTuple::ord = ord_base + s
};

// adjacent shimmers -> tuple (pair_shmmrs seq_db.rs:102-111; seq_to_compressed :233-245,:326-338; seq_to_index :386-400)
__global__ void pair_tuples_kernel(const PairParams p) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n_mm) return;
    // sequence of element e: largest s with mm_off[s] <= e
    uint32_t lo = 0, hi = p.n_seq;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (p.mm_off[mid] <= e) lo = mid; else hi = mid; }
    const uint32_t s = lo;
    if (e + 1 >= p.mm_off[s + 1]) return;  // last shimmer of its sequence: no pair starts here
    const uint64_t j = e - p.mm_off[s];
    const pgr_mm128 m0 = p.mm[e], m1 = p.mm[e + 1];
    const uint64_t s0 = m0.x >> 8, s1 = m1.x >> 8;
    FragTuple t;
    const bool fwd = p.query_mode ? (s0 < s1) : (s0 <= s1);
    t.h0 = fwd ? s0 : s1; t.h1 = fwd ? s1 : s0; t.ori = fwd ? 0u : 1u;
    t.bgn = ((uint32_t)(m0.y & 0xFFFFFFFFu) >> 1) + 1;
    t.end = ((uint32_t)(m1.y & 0xFFFFFFFFu) >> 1) + 1;
    t.sid = p.sid[s];
    t.frg_id = p.frg_base[s] + (uint32_t)j;
    t.ord = p.ord_base + s;
    p.out[p.pair_off[s] + j] = t;
}

__global__ void tuple_keys_kernel(const FragTuple *t, uint64_t n, SortKey *keys) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SortKey k; k.k0 = t[i].h0; k.k1 = t[i].h1;
    keys[i] = k;
}
// minor-key round of the multi-GPU merge: sort by the sequences' insertion ordinal first ...
__global__ void tuple_ord_keys_kernel(const FragTuple *t, uint64_t n, SortKey *keys) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SortKey k; k.k0 = 0; k.k1 = t[i].ord;
    keys[i] = k;
}
// ... then take the hash keys in that order
__global__ void tuple_keys_perm_kernel(const FragTuple *t, const uint32_t *idx, uint64_t n, SortKey *keys) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const FragTuple &x = t[idx[i]];
    SortKey k; k.k0 = x.h0; k.k1 = x.h1;
    keys[i] = k;
}

// after the sort: sigs[i] = tuple[idx[i]]; head[i] = key differs from the previous one
__global__ void csr_gather_kernel(const FragTuple *t, const SortKey *keys, const uint32_t *idx, uint64_t n, pgr_frag_sig *sigs,
                                  uint8_t *head) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const FragTuple &x = t[idx[i]];
    pgr_frag_sig sg;
    sg.frg_id = x.frg_id; sg.sid = x.sid; sg.bgn = x.bgn; sg.end = x.end; sg.ori = (uint8_t)x.ori;
    sg.pad_[0] = sg.pad_[1] = sg.pad_[2] = 0;
    sigs[i] = sg;
    head[i] = (i == 0 || keys[i].k0 != keys[i - 1].k0 || keys[i].k1 != keys[i - 1].k1) ? 1 : 0;
}

// ordered compaction of the heads (reuses the block-sum / block-prefix scheme of the level kernels)
constexpr int CS_NT = 256, CS_PER = 8, CS_BLK = CS_NT * CS_PER;
__global__ void __launch_bounds__(CS_NT) csr_count_kernel(const uint8_t *head, uint64_t n, uint32_t *block_sum) {
    __shared__ uint32_t wsum[CS_NT / 32];
    const uint64_t i0 = (uint64_t)blockIdx.x * CS_BLK;
    uint32_t cnt = 0;
    for (int j = 0; j < CS_PER; j++) {
        const uint64_t i = i0 + (uint64_t)j * CS_NT + threadIdx.x;
        if (i < n) cnt += head[i];
    }
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_down_sync(0xFFFFFFFFu, cnt, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int i = 0; i < CS_NT / 32; i++) t += wsum[i]; block_sum[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(CS_NT) csr_write_kernel(const uint8_t *head, const SortKey *keys, uint64_t n, const uint64_t *block_prefix,
                                                           SortKey *ukeys, uint64_t *offsets) {
    __shared__ uint32_t wsum[CS_NT / 32];
    const uint64_t i0 = (uint64_t)blockIdx.x * CS_BLK + (uint64_t)threadIdx.x * CS_PER;
    uint8_t f[CS_PER];
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < CS_PER; j++) { f[j] = (i0 + j < n) ? head[i0 + j] : 0; cnt += f[j]; }
    uint32_t incl = cnt;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((threadIdx.x & 31) >= d) incl += t; }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t wb = 0;
    for (uint32_t j = 0; j < (threadIdx.x >> 5); j++) wb += wsum[j];
    uint64_t rank = block_prefix[blockIdx.x] + wb + incl - cnt;
#pragma unroll
    for (int j = 0; j < CS_PER; j++)
        if (f[j]) { ukeys[rank] = keys[i0 + j]; offsets[rank] = i0 + j; rank++; }
}

// binary search of one key in the sorted unique keys; returns the key's index or -1
__device__ __forceinline__ int64_t find_key(const SortKey *ukeys, uint64_t n_keys, uint64_t h0, uint64_t h1) {
    uint64_t lo = 0, hi = n_keys;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        const SortKey k = ukeys[mid];
        if (k.k0 < h0 || (k.k0 == h0 && k.k1 < h1)) lo = mid + 1; else hi = mid;
    }
    if (lo < n_keys && ukeys[lo].k0 == h0 && ukeys[lo].k1 == h1) return (int64_t)lo;
    return -1;
}

// raw_query_fragment (seq_db.rs:1200-1228): per query pair the range of signatures of its key (count 0 if absent)
__global__ void lookup_kernel(const FragTuple *qt, uint64_t n_q, const SortKey *ukeys, const uint64_t *offsets, uint64_t n_keys,
                              uint64_t *hit_begin, uint32_t *hit_count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_q) return;
    const int64_t kx = find_key(ukeys, n_keys, qt[i].h0, qt[i].h1);
    if (kx < 0) { hit_begin[i] = 0; hit_count[i] = 0; }
    else { hit_begin[i] = offsets[kx]; hit_count[i] = (uint32_t)(offsets[kx + 1] - offsets[kx]); }
}

}  // namespace pgr
