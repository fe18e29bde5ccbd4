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
// shmmrutils.rs:516-524), so the thread jumps over the blocks of the all-invalid bitmap, emitting only the two ends of the
// stretch (the middle cannot survive the min_span filter); a run entered with fmmer == rmmer (e.g. leading N, all-zero
// registers) pushes nothing at all.
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

// a marked block starts a cluster iff none of the `gap` blocks before it (inside its own sequence) is marked.
// Two launches (WRITE = 0: starts per CTA -> block_scan_kernel -> WRITE = 1: clusters in bitmap order, which is sequence
// order for a store laid out by pgr_b200_ctx_upload): no atomics, no sort.
constexpr int CF_NT = 256;
// sequence (index into the offset-sorted table) that holds block g, found by walking forward from a cached one: the marks of a
// thread's word nearly always belong to one sequence, and sequences are far longer than a word's 1024 bases
struct SeqCursor { uint32_t lo = 0xFFFFFFFFu; uint64_t first = 0, next_first = 0; };
__device__ __forceinline__ void seq_of_block(const ClusterFindParams &p, uint64_t g, SeqCursor &c) {
    if (c.lo == 0xFFFFFFFFu) {
        uint32_t lo = 0, hi = p.n_seq;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if ((p.s_off[mid] >> 5) <= g) lo = mid; else hi = mid; }
        c.lo = lo;
    } else {
        while (c.lo + 1 < p.n_seq && (p.s_off[c.lo + 1] >> 5) <= g) c.lo++;
    }
    c.first = p.s_off[c.lo] >> 5;
    c.next_first = c.lo + 1 < p.n_seq ? (p.s_off[c.lo + 1] >> 5) : ~0ull;
}
template <int WRITE>
__global__ void __launch_bounds__(CF_NT) cluster_find_kernel(const ClusterFindParams p, uint32_t *cta_count, const uint64_t *cta_prefix) {
    __shared__ uint32_t wsum[CF_NT / 32];
    const uint64_t wi = p.word_lo + (uint64_t)blockIdx.x * CF_NT + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool live = wi < p.word_hi;
    const uint32_t word0 = live ? p.bits[wi] : 0u;
    const uint32_t prev = (live && word0 && wi > 0) ? p.bits[wi - 1] : 0u;   // the gap (< 32 blocks) reaches at most one word back
    uint32_t word = word0, starts = 0;                // starts bit b: block wi*32+b starts a cluster
    SeqCursor cur;
    while (word) {
        const uint32_t b = __ffs(word) - 1;
        word &= word - 1;
        const uint64_t g = wi * 32 + b;
        // the 32 blocks before g, block g-1 in bit 31: marks among the last `gap` of them (inside the block's own sequence) mean
        // that g continues a cluster
        const uint32_t before = __funnelshift_r(prev, word0, b);
        uint32_t win = before >> (32 - p.gap);
        if (win) {
            if (cur.lo == 0xFFFFFFFFu || g >= cur.next_first) seq_of_block(p, g, cur);
            const uint64_t room = g - cur.first;      // blocks of the sequence before g
            if (room < p.gap) win = room ? (before >> (32 - (uint32_t)room)) : 0u;
        }
        if (win) continue;
        starts |= 1u << b;
    }
    const uint32_t cnt = __popc(starts);
    uint32_t incl = cnt;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += v; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
    for (int i = 0; i < CF_NT / 32; i++) { if (i < warp) wbase += wsum[i]; total += wsum[i]; }
    if (!WRITE) { if (threadIdx.x == 0) cta_count[blockIdx.x] = total; return; }
    uint64_t dst = cta_prefix[blockIdx.x] + wbase + incl - cnt;
    while (starts) {
        const uint32_t b = __ffs(starts) - 1;
        starts &= starts - 1;
        const uint64_t g = wi * 32 + b;
        if (cur.lo == 0xFFFFFFFFu || g >= cur.next_first || g < cur.first) { if (g < cur.first) cur.lo = 0xFFFFFFFFu; seq_of_block(p, g, cur); }
        if (dst < p.cap) { Cluster c; c.sid = p.s_sid[cur.lo]; c.pos = (uint32_t)((g - cur.first) << 5); p.out[dst] = c; }
        dst++;
    }
}

constexpr int FILL_KEEP = 64;   // entries kept at both ends of a jumped-over stretch of equal keys (>= 4 * (r_max - 1) + slack)

struct ReplayClusterParams {
    const uint8_t *seq; const uint64_t *off; const uint32_t *len;
    const uint32_t *mark_bits, *allinv_bits;
    const Cluster *clusters;
    const uint32_t *list; uint32_t n_list;   // cluster indices this launch works on (NULL = 0 .. n_list-1)
    uint32_t w, k, gap;
    // per cluster outputs; q0 = UINT32_MAX encodes "from the start" (-1), q1 = UINT32_MAX "to the end"
    uint32_t *q0, *q1;
    uint32_t *t_stop;             // position at which the replay stopped (UINT32_MAX = ran to the end): marked blocks up to it
                                  // belong to this patch, also those that the finder took for the start of another cluster
    uint64_t *n_add;              // entries of the patch, fills included
    uint32_t *n_fill;             // invalid runs of the patch that were crossed by a jump
    unsigned long long *cycles;   // SM clocks spent on the cluster (count pass; diagnostics), or NULL
    // pass 1 inputs
    const uint64_t *entry_off;    // [n_clusters] where the patch's entries go (COOP = 1, MODE 1)
    pgr_mm128 *entries;
    pgr_mm128 *slab;              // [n_clusters * RC_SLAB] (COOP = 0)
    pgr_mm128 *hslab;             // [n_list * RC_HSLAB] (COOP = 1, MODE 0): the count pass of a heavy cluster keeps its first RC_HSLAB
                                  // entries, so that only the clusters with more need the write pass; NULL = count only
};

// k-mer registers after the bytes [.., upto): the last k machine-valid bases before `upto` (shmmrutils.rs:461-476 updates
// the registers on valid bases only); `since_bad` = bytes since the last one outside ACGTacgt
struct MachineRegs { uint64_t f0, f1, r0, r1; int64_t since_bad; };
__device__ __forceinline__ MachineRegs load_regs(const uint8_t *sq, int64_t upto, int64_t k, bool from_zero) {
    const uint64_t mask = ~0ull >> (64 - k);
    const uint32_t shift = (uint32_t)k - 1;
    MachineRegs m; m.f0 = m.f1 = m.r0 = m.r1 = 0; m.since_bad = 1 << 30;
    int64_t b = 0;
    if (!from_zero) { b = upto; int64_t got = 0; while (b > 0 && got < k) { b--; if (base_code(sq[b]) < 4) got++; } }
    for (int64_t q = b; q < upto; q++) {
        const uint32_t ch = sq[q];
        const uint32_t c = base_code(ch);
        if (c < 4) {
            m.f0 = ((m.f0 << 1) | (c & 1)) & mask; m.f1 = ((m.f1 << 1) | (c >> 1)) & mask;
            const uint64_t rc = 3 ^ c;
            m.r0 = ((m.r0 >> 1) | ((rc & 1) << shift)) & mask; m.r1 = ((m.r1 >> 1) | ((rc >> 1) << shift)) & mask;
        }
        m.since_bad = byte_is_acgt(ch) ? m.since_bad + 1 : 0;
    }
    return m;
}

// One WARP per cluster.  Lane 0 runs the machine.  When the machine's mdist is stuck (mdist >= w: it was set by position at
// a rescan behind skipped pushes, shmmrutils.rs:513, and rule (1) can never fire again) it emits nothing until rule (2)
// meets a key <= min_mer.x — a wait with a heavy tail (the minimum of a window is a small key) that the warp spends
// together: every lane hashes 32 positions per round and the first hit, or the first byte outside ACGTacgt, is where
// lane 0 resumes.
//
// Two forms.  COOP = 0: one THREAD per cluster, entries into the cluster's slab of RC_SLAB entries, at most RC_BUDGET machine
// steps — what nearly every cluster needs (a few hundred steps, a few dozen entries), with all 32 lanes of a warp at work.
// A cluster that runs out of steps or of slab is reported as heavy (t_stop = RC_HEAVY) and redone by COOP = 1: one WARP per
// cluster, count pass (MODE 0) and write pass (MODE 1) at exact offsets, lanes 1..31 joining for the stuck waits.
constexpr int RC_NT = 128;
constexpr int RC_SLAB = 96;
constexpr int RC_BUDGET = 2048;
constexpr int RC_HSLAB = 1024;
constexpr uint32_t RC_HEAVY = 0xFFFFFFFEu;
template <int MODE, int COOP>
__global__ void __launch_bounds__(RC_NT) cluster_replay_kernel(const ReplayClusterParams p) {
    const uint32_t li = COOP ? (blockIdx.x * (uint32_t)RC_NT + threadIdx.x) >> 5 : blockIdx.x * (uint32_t)RC_NT + threadIdx.x;
    const int lane = COOP ? (threadIdx.x & 31) : 0;
    if (li >= p.n_list) return;                       // COOP: whole warps leave together
    const uint32_t ci = p.list ? p.list[li] : li;
    const long long t_begin = clock64();
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
    // ---- machine state (meaningful in lane 0 only) ----
    uint64_t f0 = 0, f1 = 0, r0 = 0, r1 = 0;
    int64_t since_bad = 1 << 30;
    uint64_t rx[128]; uint32_t ry[128];
    uint32_t r_start = 0, r_end = 0, r_len = 0;
    uint64_t min_x = ~0ull, mdist = 0, last_x = ~0ull;
    int64_t eq_count = 0;                             // most recent pushes that carry the same key
    int64_t p_last = pstar + 31;                      // the disturbance can sit anywhere in the marked block
    uint32_t q0 = 0xFFFFFFFFu, q1 = 0xFFFFFFFFu, last_emit = 0xFFFFFFFFu, n_fill = 0, t_stop = 0xFFFFFFFFu;
    uint64_t n_add = 0;
    bool done = false;
    int64_t pos = S, no_scan_before = 0;
    pgr_mm128 *dst = COOP ? (MODE ? p.entries + p.entry_off[ci] : (p.hslab ? p.hslab + (size_t)li * RC_HSLAB : nullptr)) : p.slab + (size_t)ci * RC_SLAB;
    const uint64_t wcap = COOP ? (MODE ? ~0ull : (p.hslab ? (uint64_t)RC_HSLAB : 0ull)) : (uint64_t)RC_SLAB;   // entries below it are written
    int64_t steps = 0;
    bool heavy = false;
    if (lane == 0) {
        // registers at S; the bytes before S are clean here: the previous cluster ended more than `gap` blocks earlier
        const MachineRegs m = load_regs(sq, S, k, from_start);
        f0 = m.f0; f1 = m.f1; r0 = m.r0; r1 = m.r1; since_bad = m.since_bad;
        for (int64_t i = 0; i < w; i++) { rx[i] = ~0ull; ry[i] = ~0u; }
    }
    for (;;) {
        int64_t scan_from = -1;
        uint64_t thr = 0;
        if (lane == 0) {
            for (; pos < L && !done; pos++) {
                if (!COOP && (++steps > RC_BUDGET || n_add > RC_SLAB)) { heavy = true; break; }
                if ((pos & 31) == 0 || pos == S) {
                    // entering a marked block: the machine must walk all of it before it may stop (the stop test below looks
                    // at the blocks AFTER the current one)
                    if (bit_at(p.mark_bits, blk0 + (uint64_t)(pos >> 5))) p_last = max(p_last, (pos | 31));
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
                        if (b > (pos >> 5) + 1) { pos = (b << 5) - 1; p_last = max(p_last, pos); since_bad = 0; }
                    }
                    continue;
                }
                if (pos < k) continue;
                const bool rev = r0 < f0;
                const uint64_t h = u64hash_dev(rev ? r0 : f0) ^ u64hash_dev((rev ? r1 : f1) ^ HASH_XOR);   // one pair of hashes on the selected strand, no divergent branch
                const uint64_t mx = (h << 8) | (uint64_t)k;
                const uint32_t my = ((uint32_t)pos << 1) | (rev ? 1u : 0u);
                rx[r_end] = mx; ry[r_end] = my;
                r_end = (r_end + 1 == (uint32_t)w) ? 0u : r_end + 1;
                if (r_len < (uint32_t)w) r_len++; else r_start = (r_start + 1 == (uint32_t)w) ? 0u : r_start + 1;
                eq_count = (mx == last_x) ? eq_count + 1 : 1;
                last_x = mx;
                bool emitted = false, emitted_here = false;
                if (mdist == (uint64_t)(w - 1)) {
                    // one pass for the minimum, how often it occurs and where; the ring-order pass below only runs over a single
                    // slot when the minimum is unique (the common case outside repeats)
                    uint64_t mn = ~0ull;
                    uint32_t n_min = 0, slot_min = 0;
                    for (uint32_t i = 0; i < r_len; i++) {
                        const uint64_t x = rx[i];
                        if (x < mn) { mn = x; n_min = 1; slot_min = i; } else if (x == mn) n_min++;
                    }
                    uint32_t last_y = 0;
                    const bool one = n_min == 1 && r_len == (uint32_t)w;
                    for (uint32_t i = one ? (uint32_t)w - 1 : 0u, sl = one ? slot_min : r_start; i < (uint32_t)w; i++, sl = (sl + 1 == (uint32_t)w) ? 0u : sl + 1) {
                        if (rx[sl] == mn) {
                            if (pos >= T0) {
                                if (n_add < wcap) { pgr_mm128 mm; mm.x = rx[sl]; mm.y = ((uint64_t)sid << 32) | ry[sl]; dst[n_add] = mm; }
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
                        if (n_add < wcap) { pgr_mm128 mm; mm.x = mx; mm.y = ((uint64_t)sid << 32) | my; dst[n_add] = mm; }
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
                    // saturated inside an invalid run: the ring holds w copies of the stale key, every further position of the
                    // run below E is emitted (rule (2), shmmrutils.rs:516-524).  Jump over the all-invalid blocks.
                    int64_t b = (pos >> 5) + 1;
                    while (b < n_blk && (b << 5) + 31 < E && bit_at(p.allinv_bits, blk0 + (uint64_t)b)) b++;
                    const int64_t P = (b << 5) - 1;
                    if (P > pos) {
                        // positions pos+1 .. P are all emitted with the stale key.  Only the first and last FILL_KEEP of them
                        // are materialised: an entry whose neighbours in the list carry the same key survives reduce_shmmr
                        // (ties are inclusive, shmmrutils.rs:398-405) and is then dropped by the min_span filter (px != x &&
                        // x != nx, :547), and nothing outside the stretch sees more than r-1 <= 11 of its entries per level.
                        const int64_t cnt = P - pos;
                        for (int64_t i = 1; i <= cnt; i++) {
                            if (i > FILL_KEEP && i <= cnt - FILL_KEEP) { i = cnt - FILL_KEEP; continue; }
                            if (n_add < wcap) { pgr_mm128 mm; mm.x = mx; mm.y = ((uint64_t)sid << 32) | ((uint64_t)(uint32_t)(pos + i) << 1) | (rev ? 1u : 0u); dst[n_add] = mm; }
                            n_add++;
                        }
                        n_fill++;
                        for (int64_t i = 0; i < w; i++) { rx[i] = mx; ry[i] = ((uint32_t)(P - w + 1 + i) << 1) | (rev ? 1u : 0u); }
                        r_start = 0; r_end = 0; r_len = (uint32_t)w;
                        last_emit = (uint32_t)P;
                        pos = P; p_last = max(p_last, P); since_bad = 0;
                        continue;
                    }
                }
                if (c >= 4 && (pos & 31) == 31 && mdist >= (uint64_t)w && mx > min_x) {
                    // a STUCK machine inside an invalid run whose stale key is above min_mer.x: neither rule fires until the
                    // run ends; the pushes only refill the ring with copies of the stale key
                    int64_t b = (pos >> 5) + 1;
                    while (b < n_blk && bit_at(p.allinv_bits, blk0 + (uint64_t)b)) b++;
                    const int64_t P = (b << 5) - 1;
                    if (P - pos >= w) {
                        mdist += (uint64_t)(P - pos);
                        eq_count += P - pos;
                        for (int64_t i = 0; i < w; i++) { rx[i] = mx; ry[i] = ((uint32_t)(P - w + 1 + i) << 1) | (rev ? 1u : 0u); }
                        r_start = 0; r_end = 0; r_len = (uint32_t)w;
                        pos = P; p_last = max(p_last, P); since_bad = 0;
                        continue;
                    }
                }
                if (emitted && pos >= T0 && pos >= p_last + 2 * w && pos < E - w) {
                    // the machine is back in its normal regime; stop unless another marked block follows within the cluster
                    // gap.  (A stuck machine can run far beyond the cluster's own blocks before it gets here: clusters it
                    // passed on the way are dropped by the host.)
                    bool clear = true;
                    for (int64_t b = (pos >> 5) + 1; b <= (pos >> 5) + gap && b < n_blk && clear; b++) clear = !bit_at(p.mark_bits, blk0 + (uint64_t)b);
                    if (clear) { q1 = last_emit; t_stop = (uint32_t)pos; done = true; break; }
                }
                if (COOP && mdist >= (uint64_t)w && since_bad >= k && pos >= no_scan_before && pos + 1 + 8 * w < min(E, L)) {
                    scan_from = pos + 1; thr = min_x; pos++;     // stuck: look for the next rule (2) emission together
                    break;
                }
            }
        }
        if (!COOP) break;
        scan_from = __shfl_sync(0xFFFFFFFFu, scan_from, 0);
        if (scan_from < 0) break;                     // stopped, or walked to the end of the sequence
        thr = __shfl_sync(0xFFFFFFFFu, thr, 0);
        // ---- cooperative scan: first position >= scan_from (below E) whose key is <= thr, or first byte outside ACGTacgt ----
        const int64_t limit = min(E, L);
        uint32_t ev = 0xFFFFFFFFu;
        for (int64_t base = scan_from; base < limit && ev == 0xFFFFFFFFu; base += 1024) {
            const int64_t b = base + 32 * lane;
            uint32_t mine = 0xFFFFFFFFu;
            if (b < limit) {
                uint64_t g0 = 0, g1 = 0, h0 = 0, h1 = 0;       // forward / reverse registers of this lane
                bool lead_ok = true;
                for (int64_t q = max((int64_t)0, b - (k - 1)); q < b; q++) {
                    const uint32_t ch = sq[q];
                    const uint32_t c = base_code(ch);
                    lead_ok = lead_ok && byte_is_acgt(ch);
                    if (c < 4) {
                        g0 = ((g0 << 1) | (c & 1)) & mask; g1 = ((g1 << 1) | (c >> 1)) & mask;
                        const uint64_t rc = 3 ^ c;
                        h0 = ((h0 >> 1) | ((rc & 1) << shift)) & mask; h1 = ((h1 >> 1) | ((rc >> 1) << shift)) & mask;
                    }
                }
                // a lead-in with a foreign byte belongs to an earlier lane's range, which reports it: this lane stays silent
                for (int64_t q = b; q < min(b + 32, limit) && lead_ok && mine == 0xFFFFFFFFu; q++) {
                    const uint32_t ch = sq[q];
                    if (!byte_is_acgt(ch)) { mine = (uint32_t)q; break; }
                    const uint32_t c = base_code(ch);
                    g0 = ((g0 << 1) | (c & 1)) & mask; g1 = ((g1 << 1) | (c >> 1)) & mask;
                    const uint64_t rc = 3 ^ c;
                    h0 = ((h0 >> 1) | ((rc & 1) << shift)) & mask; h1 = ((h1 >> 1) | ((rc >> 1) << shift)) & mask;
                    if (g0 == h0 && g1 == h1) continue;        // palindrome: not pushed
                    const bool rv = h0 < g0;
                    const uint64_t hh = u64hash_dev(rv ? h0 : g0) ^ u64hash_dev((rv ? h1 : g1) ^ HASH_XOR);
                    if (((hh << 8) | (uint64_t)k) <= thr) mine = (uint32_t)q;
                }
            }
            ev = __reduce_min_sync(0xFFFFFFFFu, mine);
        }
        if (lane == 0) {
            if (ev == 0xFFFFFFFFu) {
                pos = L;                              // no key <= min_mer.x below E: the machine never emits again
            } else {
                // resume far enough before the event for the stop test to see every disturbance within 2w of its stop
                const int64_t r = (int64_t)ev - 4 * w;
                if (r > scan_from + k) {
                    const MachineRegs m = load_regs(sq, r, k, false);   // [scan_from, ev) is clean
                    f0 = m.f0; f1 = m.f1; r0 = m.r0; r1 = m.r1; since_bad = m.since_bad;
                    // pushes between scan_from and r: mdist stays stuck (it only grows), the ring is refilled from r on and is
                    // next read w-1 pushes after the rule (2) emission at ev, when it holds pushes from ev onwards only
                    mdist += (uint64_t)(r - scan_from);
                    for (int64_t i = 0; i < w; i++) { rx[i] = ~0ull; ry[i] = ~0u; }
                    r_start = 0; r_end = 0; r_len = 0; eq_count = 0; last_x = ~0ull;
                    pos = r;
                    if (bit_at(p.mark_bits, blk0 + (uint64_t)(pos >> 5))) p_last = max(p_last, (pos | 31));
                }
                no_scan_before = (int64_t)ev + 1;
            }
        }
    }
    if (lane == 0) {
        if (from_start) q0 = 0xFFFFFFFFu;
        if (!COOP && n_add > RC_SLAB) heavy = true;
        p.q0[ci] = q0; p.q1[ci] = q1;
        if (!MODE) { p.n_add[ci] = n_add; p.n_fill[ci] = n_fill; p.t_stop[ci] = heavy ? RC_HEAVY : t_stop; if (p.cycles) p.cycles[ci] = (unsigned long long)(clock64() - t_begin); }
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
    const uint32_t *pseq; const uint64_t *pn_add;
    const pgr_mm128 *const *psrc;                      // per patch: its entries (a slab or a slice of the heavy patches' buffer)
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
    const pgr_mm128 *src = p.psrc[j];
    pgr_mm128 *dst = p.flat1 + p.pdst[j];
    for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

}  // namespace pgr
