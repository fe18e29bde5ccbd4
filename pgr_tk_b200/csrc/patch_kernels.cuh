// patch_kernels.cuh — exact local replay of the reference's level-0 machine (shmmrutils.rs:440-530) around the places
// where the tile kernel's local rule does not hold:
//   * a pushed position with fmmer == rmmer (reverse-complement palindrome, shmmrutils.rs:477): it is not pushed and
//     mdist is not advanced, while rescans reset mdist by position (:513) — duplicate emissions and a stuck mdist become
//     possible until the machine's next emission;
//   * a byte outside ACGTacgt: the k-mer registers keep their value and the stale k-mer is pushed again (:461-476), so the
//     key at p depends on the last k VALID bases, not on the k bytes ending at p.
// The tile kernel marks every 32-base block holding either in a bitmap (and the blocks made of 32 invalid bytes in a second
// one).  Marked blocks separated by fewer than `gap` blocks form a CLUSTER; one thread per cluster replays the machine:
//   start S = p* - 2w (p* = first position of the cluster's first block; a fresh machine is in sync with the true one after
//   w pushes without a skip; S is pulled back so that those w pushes lie before L-w+k, or set to k near the sequence start);
//   emissions at times < T0 = S + w are discarded, q0 = position of the last of them;
//   the replay ends at the first emission at a time t with p_last + 2w <= t < L-w+k - w once the next `gap` blocks are
//   unmarked (q1 = the position emitted last), else it runs to the end of the sequence.
// The patch replaces the tile kernel's entries with q0 < pos <= q1.  A long run of invalid bytes is not walked: once the
// ring buffer holds w copies of the stale key the machine's state is translation invariant (every position is emitted,
// shmmrutils.rs:516-524), so the thread jumps over the blocks of the all-invalid bitmap and leaves a FILL segment that a
// separate kernel expands; a run entered with fmmer == rmmer (e.g. leading N, all-zero registers) pushes nothing at all.
#pragma once
#include "shmmr_kernels.cuh"

namespace pgr {

struct Cluster { uint32_t sid, pos; };   // sequence ordinal, first position of the cluster's first marked block

struct ClusterFindParams {
    const uint32_t *bits;          // mark bitmap over the store's 32-base blocks
    uint64_t word_lo, word_hi;     // bitmap words to scan
    const uint64_t *s_off;         // [n_seq] sequence offsets sorted ascending ...
    const uint32_t *s_len;         // ... their lengths ...
    const uint32_t *s_sid;         // ... and their ordinals
    uint32_t n_seq;
    uint32_t gap;                  // blocks without a mark that separate two clusters
    Cluster *out; uint32_t cap; uint32_t *n_out;
};

__device__ __forceinline__ bool bit_at(const uint32_t *bits, uint64_t g) { return (bits[g >> 5] >> (g & 31)) & 1u; }

// a marked block starts a cluster iff none of the `gap` blocks before it (inside its own sequence) is marked
__global__ void cluster_find_kernel(const ClusterFindParams p) {
    const uint64_t wi = p.word_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= p.word_hi) return;
    uint32_t word = p.bits[wi];
    while (word) {
        const uint32_t b = __ffs(word) - 1;
        word &= word - 1;
        const uint64_t g = wi * 32 + b;
        // sequence of block g: last sorted sequence with off/32 <= g
        uint32_t lo = 0, hi = p.n_seq;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if ((p.s_off[mid] >> 5) <= g) lo = mid; else hi = mid; }
        const uint64_t first = p.s_off[lo] >> 5;
        const uint64_t from = (g - first > p.gap) ? g - p.gap : first;
        bool start = true;
        for (uint64_t q = from; q < g && start; q++) start = !bit_at(p.bits, q);
        if (!start) continue;
        const uint32_t slot = atomicAdd(p.n_out, 1u);
        if (slot < p.cap) { Cluster c; c.sid = p.s_sid[lo]; c.pos = (uint32_t)((g - first) << 5); p.out[slot] = c; }
    }
}

struct FillSeg { uint64_t dst; uint64_t x; uint32_t first_pos, count, sid, strand; };   // entries[dst + i] = {x, sid, first_pos + i, strand}

struct ReplayClusterParams {
    const uint8_t *seq; const uint64_t *off; const uint32_t *len;
    const uint32_t *mark_bits, *allinv_bits;
    const Cluster *clusters; uint32_t n_clusters;
    uint32_t w, k, gap;
    // per cluster outputs; q0 = UINT32_MAX encodes "from the start" (-1), q1 = UINT32_MAX "to the end"
    uint32_t *q0, *q1;
    uint32_t *t_stop;             // position at which the replay stopped (UINT32_MAX = ran to the end): marked blocks up to it
                                  // belong to this patch, also those that the finder took for the start of another cluster
    uint64_t *n_add;              // entries of the patch, fills included
    uint32_t *n_fill;             // fill segments of the patch
    // pass 1 inputs
    const uint64_t *entry_off;    // [n_clusters] where the patch's entries go
    const uint32_t *fill_off;     // [n_clusters] first fill segment of the patch
    pgr_mm128 *entries;
    FillSeg *fills;
};

template <int MODE>
__global__ void cluster_replay_kernel(const ReplayClusterParams p) {
    const uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= p.n_clusters) return;
    const uint32_t sid = p.clusters[ci].sid;
    const uint64_t soff = p.off[sid];
    const uint8_t *sq = p.seq + soff;
    const uint64_t blk0 = soff >> 5;                  // bitmap index of the sequence's block 0
    const int64_t L = p.len[sid];
    const int64_t w = p.w, k = p.k, gap = p.gap;
    const uint64_t mask = ~0ull >> (64 - k);
    const uint32_t shift = (uint32_t)k - 1;
    const int64_t E = L - w + k;                      // rule (2) active for pos < E
    const int64_t n_blk = (L + 31) >> 5;
    const int64_t pstar = p.clusters[ci].pos;
    int64_t S = pstar - 2 * w;
    if (S + w > E - 1) S = E - 1 - w;
    bool from_start = false;
    if (S < k + w + 1) { S = k; from_start = true; }
    const int64_t T0 = from_start ? k : S + w;
    // registers at S: the last k VALID bases before S (shmmrutils.rs:461-476 updates them on valid bases only); from the
    // true start they begin at zero.  The bytes before S are clean here: the previous cluster ended more than `gap` blocks
    // earlier.  `since_bad` counts the bytes since the last one outside ACGTacgt: the tile kernel's key at p is the true key
    // iff the k bytes ending at p are all ACGTacgt.
    uint64_t f0 = 0, f1 = 0, r0 = 0, r1 = 0;
    int64_t r_begin = 0;
    if (!from_start) {
        int64_t b = S, got = 0;
        while (b > 0 && got < k) { b--; if (base_code(sq[b]) < 4) got++; }
        r_begin = b;
    }
    int64_t since_bad = 1 << 30;
    for (int64_t q = r_begin; q < S; q++) {
        const uint32_t ch = sq[q];
        const uint32_t c = base_code(ch);
        if (c < 4) {
            f0 = ((f0 << 1) | (c & 1)) & mask; f1 = ((f1 << 1) | (c >> 1)) & mask;
            const uint64_t rc = 3 ^ c;
            r0 = ((r0 >> 1) | ((rc & 1) << shift)) & mask; r1 = ((r1 >> 1) | ((rc >> 1) << shift)) & mask;
        }
        since_bad = byte_is_acgt(ch) ? since_bad + 1 : 0;
    }
    uint64_t rx[128]; uint32_t ry[128];
    for (int64_t i = 0; i < w; i++) { rx[i] = ~0ull; ry[i] = ~0u; }
    uint32_t r_start = 0, r_end = 0, r_len = 0;
    uint64_t min_x = ~0ull, mdist = 0, last_x = ~0ull;
    int64_t eq_count = 0;                             // most recent pushes that carry the same key
    int64_t p_last = pstar + 31;                      // the disturbance can sit anywhere in the marked block
    int64_t D = pstar >> 5;                           // last marked block met so far
    uint32_t q0 = 0xFFFFFFFFu, q1 = 0xFFFFFFFFu, last_emit = 0xFFFFFFFFu, n_fill = 0, t_stop = 0xFFFFFFFFu;
    uint64_t n_add = 0;
    bool done = false;
    pgr_mm128 *dst = MODE ? p.entries + p.entry_off[ci] : nullptr;
    FillSeg *fills = MODE ? p.fills + p.fill_off[ci] : nullptr;
    for (int64_t pos = S; pos < L && !done; pos++) {
        if ((pos & 31) == 0 || pos == S) {
            // entering a marked block: the machine must walk all of it before it may stop (the stop test below looks at the
            // blocks AFTER the current one)
            if (bit_at(p.mark_bits, blk0 + (uint64_t)(pos >> 5))) { D = max(D, pos >> 5); p_last = max(p_last, (pos | 31)); }
        }
        const uint32_t ch = sq[pos];
        const uint32_t c = base_code(ch);
        if (c < 4) {
            f0 = ((f0 << 1) | (c & 1)) & mask; f1 = ((f1 << 1) | (c >> 1)) & mask;
            const uint64_t rc = 3 ^ c;
            r0 = ((r0 >> 1) | ((rc & 1) << shift)) & mask; r1 = ((r1 >> 1) | ((rc >> 1) << shift)) & mask;
        }
        since_bad = byte_is_acgt(ch) ? since_bad + 1 : 0;
        if (since_bad < k) p_last = max(p_last, pos); // the tile kernel's key here is not the reference's
        if (f0 == r0 && f1 == r1) {
            if (pos >= k) p_last = max(p_last, pos);
            if (c >= 4 && (pos & 31) == 31) {
                // inside an invalid run with palindromic stale registers: nothing is pushed until the run ends
                int64_t b = (pos >> 5) + 1;
                while (b < n_blk && bit_at(p.allinv_bits, blk0 + (uint64_t)b)) b++;
                if (b > (pos >> 5) + 1) { pos = (b << 5) - 1; p_last = max(p_last, pos); since_bad = 0; D = max(D, b - 1); }
            }
            continue;
        }
        if (pos < k) continue;
        const bool rev = r0 < f0;
        const uint64_t h = rev ? (u64hash(r0) ^ u64hash(r1 ^ HASH_XOR)) : (u64hash(f0) ^ u64hash(f1 ^ HASH_XOR));
        const uint64_t mx = (h << 8) | (uint64_t)k;
        const uint32_t my = ((uint32_t)pos << 1) | (rev ? 1u : 0u);
        rx[r_end] = mx; ry[r_end] = my;
        r_end = (r_end + 1) % (uint32_t)w;
        if (r_len < (uint32_t)w) r_len++; else r_start = (r_start + 1) % (uint32_t)w;
        eq_count = (mx == last_x) ? eq_count + 1 : 1;
        last_x = mx;
        bool emitted = false, emitted_here = false;
        if (mdist == (uint64_t)(w - 1)) {
            uint64_t mn = ~0ull;
            for (uint32_t i = 0; i < r_len; i++) if (rx[i] < mn) mn = rx[i];
            uint32_t last_y = 0;
            for (uint32_t i = 0; i < (uint32_t)w; i++) {
                const uint32_t sl = (r_start + i) % (uint32_t)w;
                if (rx[sl] == mn) {
                    if (pos >= T0) {
                        if (MODE) { pgr_mm128 mm; mm.x = rx[sl]; mm.y = ((uint64_t)sid << 32) | ry[sl]; dst[n_add] = mm; }
                        n_add++;
                    } else {
                        q0 = ry[sl] >> 1;
                    }
                    last_y = ry[sl];
                    last_emit = ry[sl] >> 1;
                    emitted = true;
                }
            }
            min_x = mn;
            mdist = (uint64_t)pos - (uint64_t)(last_y >> 1);
            emitted_here = (last_y >> 1) == (uint32_t)pos;
        } else if (mx <= min_x && pos >= w + k && pos < E && pos < L) {
            if (pos >= T0) {
                if (MODE) { pgr_mm128 mm; mm.x = mx; mm.y = ((uint64_t)sid << 32) | my; dst[n_add] = mm; }
                n_add++;
            } else {
                q0 = (uint32_t)pos;
            }
            last_emit = (uint32_t)pos;
            emitted = emitted_here = true;
            min_x = mx;
            mdist = 0;
        } else {
            mdist++;
        }
        if (c >= 4 && (pos & 31) == 31 && emitted_here && eq_count >= w && min_x == mx && mdist == 0 && pos >= T0 && pos >= w + k) {
            // saturated inside an invalid run: the ring holds w copies of the stale key, every further position of the run
            // below E is emitted (rule (2), shmmrutils.rs:516-524).  Jump over the all-invalid blocks; a fill segment
            // stands for their positions.
            int64_t b = (pos >> 5) + 1;
            while (b < n_blk && (b << 5) + 31 < E && bit_at(p.allinv_bits, blk0 + (uint64_t)b)) b++;
            const int64_t P = (b << 5) - 1;
            if (P > pos) {
                const uint64_t cnt = (uint64_t)(P - pos);
                if (MODE) {
                    FillSeg fs; fs.dst = p.entry_off[ci] + n_add; fs.x = mx; fs.first_pos = (uint32_t)(pos + 1); fs.count = (uint32_t)cnt;
                    fs.sid = sid; fs.strand = rev ? 1u : 0u;
                    fills[n_fill] = fs;
                }
                n_fill++;
                n_add += cnt;
                for (int64_t i = 0; i < w; i++) { rx[i] = mx; ry[i] = ((uint32_t)(P - w + 1 + i) << 1) | (rev ? 1u : 0u); }
                r_start = 0; r_end = 0; r_len = (uint32_t)w;
                last_emit = (uint32_t)P;
                pos = P; p_last = max(p_last, P); since_bad = 0; D = max(D, b - 1);
                continue;
            }
        }
        if (emitted && pos >= T0 && pos >= p_last + 2 * w && pos < E - w) {
            // the machine is back in its normal regime; stop unless another marked block follows within the cluster gap.
            // (A machine whose mdist got stuck behind a palindrome emits by rule (2) only and can run far beyond the
            // cluster's own blocks before it gets here: clusters it passed on the way are dropped by the host.)
            bool clear = true;
            for (int64_t b = (pos >> 5) + 1; b <= (pos >> 5) + gap && b < n_blk && clear; b++) clear = !bit_at(p.mark_bits, blk0 + (uint64_t)b);
            if (clear) { q1 = last_emit; t_stop = (uint32_t)pos; done = true; }
        }
    }
    if (from_start) q0 = 0xFFFFFFFFu;
    p.q0[ci] = q0; p.q1[ci] = q1;
    if (!MODE) { p.n_add[ci] = n_add; p.n_fill[ci] = n_fill; p.t_stop[ci] = t_stop; }
}

// expand the fill segments: one CTA per segment
__global__ void fill_segments_kernel(const FillSeg *fills, pgr_mm128 *entries) {
    const FillSeg fs = fills[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < fs.count; i += blockDim.x) {
        pgr_mm128 mm;
        mm.x = fs.x;
        mm.y = ((uint64_t)fs.sid << 32) | ((uint64_t)(fs.first_pos + i) << 1) | fs.strand;
        entries[fs.dst + i] = mm;
    }
}

// splice the patches into the flat level-0 list.  Patches are sorted by (sequence, position); per patch: lb = number of
// the sequence's entries with pos <= q0 (0 for "from the start"), ub = number with pos <= q1.
struct SpliceParams {
    const pgr_mm128 *flat0; const uint64_t *off0;      // before
    pgr_mm128 *flat1; const uint64_t *off1;            // after
    uint32_t n_seq;
    const int32_t *seq_first_patch;                    // [n_seq] first patch index of the sequence or -1
    const uint32_t *seq_n_patch;                       // [n_seq]
    const uint32_t *pq0, *pq1;                         // per patch
    uint32_t *plb, *pub;                               // per patch (bounds kernel output)
    const int64_t *pdelta;                             // per patch: cumulative (added - removed) of the sequence's patches up to and including this one
    const uint64_t *pdst;                              // per patch: where its entries start in flat1
    const uint32_t *pseq; const uint64_t *pn_add; const uint64_t *pentry_off;
    const pgr_mm128 *entries;
    uint32_t n_patches;
    uint64_t n0;
};

__device__ __forceinline__ uint32_t mm_pos32(const pgr_mm128 &m) { return (uint32_t)(m.y & 0xFFFFFFFFu) >> 1; }

__global__ void splice_bounds_kernel(const SpliceParams p) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.n_patches) return;
    const uint32_t s = p.pseq[j];
    const uint64_t b = p.off0[s], e = p.off0[s + 1];
    auto count_le = [&](uint32_t q) -> uint32_t {   // entries of s with pos <= q
        uint64_t lo = b, hi = e;
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (mm_pos32(p.flat0[mid]) <= q) lo = mid + 1; else hi = mid; }
        return (uint32_t)(lo - b);
    };
    p.plb[j] = (p.pq0[j] == 0xFFFFFFFFu) ? 0u : count_le(p.pq0[j]);
    p.pub[j] = (p.pq1[j] == 0xFFFFFFFFu) ? (uint32_t)(e - b) : count_le(p.pq1[j]);
}

__global__ void splice_copy_kernel(const SpliceParams p) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n0) return;
    const pgr_mm128 mm = p.flat0[i];
    const uint32_t s = (uint32_t)(mm.y >> 32);
    const uint64_t rel = i - p.off0[s];
    const int32_t fp = p.seq_first_patch[s];
    if (fp < 0) { p.flat1[p.off1[s] + rel] = mm; return; }
    // the sequence's patches are sorted and disjoint: last one whose lb <= rel
    uint32_t lo = 0, hi = p.seq_n_patch[s];
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (p.plb[fp + mid] <= rel) lo = mid + 1; else hi = mid; }
    int64_t delta = 0;
    if (lo > 0) {
        const uint32_t j = fp + lo - 1;
        if (rel < p.pub[j]) return;                    // inside (q0, q1]: replaced by the patch
        delta = p.pdelta[j];
    }
    p.flat1[p.off1[s] + (uint64_t)((int64_t)rel + delta)] = mm;
}

// one CTA per patch copies its entries to their place
__global__ void splice_patch_kernel(const SpliceParams p) {
    const uint32_t j = blockIdx.x;
    const uint64_t n = p.pn_add[j];
    const pgr_mm128 *src = p.entries + p.pentry_off[j];
    pgr_mm128 *dst = p.flat1 + p.pdst[j];
    for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

}  // namespace pgr
