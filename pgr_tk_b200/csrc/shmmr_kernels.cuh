// shmmr_kernels.cuh — sm_100a kernels for sequence_to_shmmrs (shmmrutils.rs:417-669).
//
// The reference walks every sequence with a sequential state machine (ring buffer, min_mer, mdist).  These kernels
// use the equivalent LOCAL rule (DESIGN.md "Local rule", property-tested in tests/test_local_rule.py):
//   level 0 : position p in [k, L) is a minimizer  <=>  x[p] is a (tie-inclusive) minimum of some window of w
//             consecutive k-mer keys that lies inside [k, min(L, L-w+k)); the last w-k positions are replayed
//             (rescans only) by one thread of the sequence's last tile;
//   level 1,2: the same window-minimum rule over list indices with window r (reduce_shmmr, applied twice);
//   min_span filter on the level-2 list.
// The rule does not hold around a byte outside ACGTacgt (the reference keeps a stale k-mer, shmmrutils.rs:459-476) or a
// pushed reverse-complement palindrome (fmmer == rmmer, shmmrutils.rs:477).  The tile kernel marks the 32-base blocks
// that contain either in a bitmap; patch_kernels.cuh turns the marked blocks into clusters and replays the reference
// machine around every cluster (one thread per cluster, long invalid runs crossed in closed form).
#pragma once
#include "common.cuh"

namespace pgr {

// tile geometry; overridable at build time for tuning (make NVCCFLAGS+=-DPGR_L0_NT=256 ...)
#ifndef PGR_L0_NT
#define PGR_L0_NT 256
#endif
#ifndef PGR_L0_MIN_CTAS
#define PGR_L0_MIN_CTAS 3
#endif
// PGR_L0_BULK=1: the tile's bytes (L0_NT x 32 B, contiguous) are fetched by ONE bulk asynchronous copy (cp.async.bulk, the
// TMA engine's 1-D form) into shared memory behind an mbarrier, issued by thread 0 one tile ahead, instead of one 32-byte
// register prefetch per thread.  A/B variant (profiles/r2_l0_kernel_bulk_ab.txt): the key loop is ALU bound and the register
// prefetch already hides the loads, while the 8 KB staging slot costs the third resident CTA at 256 threads.
#ifndef PGR_L0_BULK
#define PGR_L0_BULK 1
#endif
#ifndef PGR_L0_PREFETCH_L2
#define PGR_L0_PREFETCH_L2 0   // cp.async.bulk.prefetch.L2 of the next tile at the head of this one: measured +-0.1 %, off
#endif
constexpr int L0_NT = PGR_L0_NT;           // threads per CTA; thread t owns the 32-base block t of the tile's load region
constexpr int L0_CTX = 2;                  // leading context-only blocks (a 56-mer reaches 55 bases back)
constexpr int L0_KB = L0_NT - L0_CTX;      // blocks that get keys
constexpr int L0_KPOS = L0_KB * 32;        // key positions per tile
constexpr int L0_PADB = 4;                 // spare blocks on both sides of the smem arrays (van Herk neighbours, w <= 128)
constexpr int L0_ARR = (L0_KB + 2 * L0_PADB) * 33;  // padded u32 array length
constexpr int L0_MIN_CTAS = PGR_L0_MIN_CTAS;  // resident CTAs per SM the kernel is compiled for
#if PGR_L0_BULK
// the last 8 KB of the P array (128-byte aligned in the CTA's shared memory) receive the next tile's bytes: P is only used in
// full by phase 3, the copy is issued after it and consumed by phase 1 of the next tile
constexpr int L0_SLOT_BYTES = L0_NT * 32;
constexpr int L0_SLOT_TOP = (L0_ARR - L0_PADB * 33) * 4;   // the spare blocks at the end of P keep their zeros (phase 3 reads them)
constexpr int L0_SLOT_OFF = ((L0_SLOT_TOP - L0_SLOT_BYTES) - (((L0_SLOT_TOP - L0_SLOT_BYTES) + L0_ARR * 4) % 128));   // byte offset inside P; H (L0_ARR words) precedes P
static_assert(L0_SLOT_OFF > 0 && (L0_ARR * 4 + L0_SLOT_OFF) % 128 == 0 && L0_SLOT_OFF + L0_SLOT_BYTES <= L0_SLOT_TOP, "tile slot placement");
static_assert(L0_SLOT_OFF / 2 >= L0_KPOS + 256, "the candidate list must fit below the slot");
constexpr int L0_LISTCAP = L0_SLOT_OFF / 2; // u16 entries of the P array below the slot
constexpr int L0_CH_TOP = L0_SLOT_OFF / 4;  // first P word that belongs to the slot
#else
constexpr int L0_LISTCAP = L0_ARR * 2;     // u16 entries that fit in the P array
constexpr int L0_CH_TOP = L0_ARR;
#endif
// key prefixes of the candidates, in list order, in the P words behind the largest possible candidate list: phase 5 compares
// neighbouring candidates through it with independent loads instead of list -> H chains
constexpr int L0_CH_OFF = (L0_KPOS + 256) / 2;      // first word
constexpr int L0_CH_CAP = L0_CH_TOP - L0_CH_OFF;    // entries; a tile with more candidates takes the chained form
static_assert(L0_CH_CAP >= 1024, "room for the candidates' key prefixes");

struct L0Params {
    const uint8_t *seq;          // device sequence store
    const uint64_t *off;         // [n_seq] byte offset of each sequence (32-byte aligned)
    const uint32_t *len;         // [n_seq]
    const uint32_t *tile_prefix; // [n_seq+1] cumulative tile count
    const uint32_t *cta_tile;    // [grid+1] static tile range per CTA
    uint32_t n_seq;
    uint32_t w, k;
    uint32_t tile_stride;        // TI: output positions per tile
    uint32_t halo;               // HB*32 >= w-1
    pgr_mm128 *arena;            // level-0 output, one private chunk per CTA
    uint64_t chunk_cap;          // entries per chunk
    uint64_t *chunk_count;       // [grid] entries demanded by each CTA (may exceed chunk_cap => host retries)
    uint32_t *seq_count;         // [n_seq] level-0 entries per sequence (atomicAdd per tile)
    uint32_t *seq_flag;          // [n_seq] != 0 => sequence must be replayed sequentially (candidate list overflow; never seen)
    // disturbance bitmaps over the 32-base blocks of the sequence store (block = (seq_off + pos) / 32):
    uint32_t *mark_bits;         //   block holds a byte outside ACGTacgt or a pushed position with fmmer == rmmer
    uint32_t *allinv_bits;       //   all 32 bytes of the block are outside ACGTacgt (interior of a long invalid run)
    uint32_t *n_marks;           // number of marking events (0 => the level-0 list needs no patch)
    uint64_t m1;                 // ~0 (the -1 of the hash's first step) as a parameter: an IMAD.WIDE addend straight from the constant bank
                                 // instead of two MOVs per position
    HiShift hs;                  // opaque power-of-two factors of the IMAD.HI shifts (PGR_L0_HIMASK)
};

__device__ __forceinline__ uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }

#if PGR_L0_BULK
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// thread 0: expect `bytes` on the barrier and start the bulk copy global -> shared
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy reads of dst are done (barrier before)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// ask the L2 for the bytes of a later bulk copy (no completion to wait for)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif

// 4 ASCII bases (one little-endian word) -> 4 plane bits each in the TOP nibble of a product word, first base most
// significant.  code bit0 = bit1 of (c ^ (c >> 1)), code bit1 = bit2 of c  (A=0,C=1,G=2,T=3; case-insensitive).
// Bits 1,9,17,25 (resp. 2,10,18,26) are gathered by one multiply (no partial products collide, so no carries).
__device__ __forceinline__ void planes4(uint32_t wd, uint32_t &p0, uint32_t &p1) {
    const uint32_t t = wd ^ (wd >> 1);
    p0 = (t & 0x02020202u) * 0x40201008u;    // bit 1+8i -> 31-i
    p1 = (wd & 0x04040404u) * 0x20100804u;   // bit 2+8i -> 31-i
}
// planes4 plus a validity word: bits outside 0x20 of some byte of `diff` are set iff that byte is not one of ACGTacgt.
// The 2-bit codes, doubled, become the four nibbles of a PRMT selector that looks the expected upper-case letter up in
// the byte pool {A,-,C,-,G,-,T,-}; the caller ORs (expected ^ byte) over the block and masks the case bit once.
__device__ __forceinline__ void planes4v(uint32_t wd, uint32_t &p0, uint32_t &p1, uint32_t &diff) {
    const uint32_t t = wd ^ (wd >> 1);
    const uint32_t c0 = t & 0x02020202u, c1 = wd & 0x04040404u;
    p0 = c0 * 0x40201008u;
    p1 = c1 * 0x20100804u;
    const uint32_t m = c0 | c1;                       // per byte: code << 1 (0, 2, 4, 6) in the low nibble
    const uint32_t n = m | (m >> 4);                  // byte 0 = nibbles (code0, code1), byte 2 = nibbles (code2, code3)
    const uint32_t sel = __byte_perm(n, 0u, 0x4420);  // low half = byte 0, byte 2
    const uint32_t expect = __byte_perm(0x00430041u, 0x00540047u, sel);   // 'A', 'C' | 'G', 'T' at even pool slots
    diff |= expect ^ wd;
}

__device__ __forceinline__ bool byte_is_acgt(uint32_t c) {
    const uint32_t code = (((c ^ (c >> 1)) >> 1) & 1u) | ((c >> 1) & 2u);
    return (c & 0xDFu) == ((0x54474341u >> (8u * code)) & 0xFFu);
}

__device__ __forceinline__ uint32_t min3u(uint32_t a, uint32_t b, uint32_t c) { return min(min(a, b), c); }
__device__ __forceinline__ uint32_t max3u(uint32_t a, uint32_t b, uint32_t c) { return max(max(a, b), c); }

struct DescCache {
    uint32_t sid;                    // UINT32_MAX = nothing cached
    uint32_t tp_lo, tp_hi;           // tile_prefix[sid], tile_prefix[sid + 1]
    uint32_t len;
    uint64_t off;
};
struct L0Smem {
    uint32_t H[L0_ARR];      // key prefix: the top bits of MM128.x (hash bits 32..55 in the K > 32 kernels, 24..55 in the
                             // generic one); padded index q + q/32 + PADB*33
    uint32_t P[L0_ARR];      // van Herk exchange: prefix minima (pass 1) then suffix maxima (pass 2); afterwards it
                             // holds `list`: ordered key indices (u16) of the selected positions + tail replay entries
    uint32_t F0[L0_NT + 8], F1[L0_NT + 8];   // bit planes per 32-base block, first base in the MOST significant bit;
                                             // the complement planes (first base in the LEAST significant bit) are ~brev()
    uint32_t bext[L0_KB + 2 * L0_PADB];   // block min (pass 1) / block max (pass 2)
    alignas(16) uint32_t wsum[(L0_NT / 32 + 3) / 4 * 4];
    uint32_t n_list;
    // tile descriptors, double-buffered: thread 0 prepares tile t+1 while the CTA works on tile t, and every thread
    // issues its 32-byte load for tile t+1 before the key loop of tile t (software prefetch across tiles)
    struct TileDesc {
        uint32_t seq_id, seq_len;
        int32_t keys_start;      // sequence position of key index 0 (multiple of 32, may be negative)
        int32_t out_lo, out_hi;  // output position range [out_lo, out_hi)
        uint32_t is_last;
        uint32_t bad;            // tile saw an invalid byte (or lost a palindrome record)
        uint32_t n_tail, any_reject;
        uint32_t n_pre, n_post;  // candidates below / above the output range (phase 4)
        uint32_t pad_;
        uint64_t seq_off;
    } td[2];
    DescCache dc;            // thread 0 only (shared memory instead of six registers of every thread)
};

__device__ __forceinline__ int pidx(int q) { return q + (q >> 5) + L0_PADB * 33; }  // q may be negative (>= -PADB*32)

// complement plane word of block j (blocks past the tile read as 0, like the reference's empty registers)
__device__ __forceinline__ uint32_t rplane(const uint32_t *F, int j) { return j < L0_NT ? ~__brev(F[j]) : 0u; }

// k-mer registers of key index q rebuilt from the plane words (same arithmetic as the key loop)
struct KmerRegs { uint64_t f0, f1, r0, r1; };
template <class SM>   // any shared-memory struct with the plane arrays F0 / F1 (L0Smem, SketchSmem)
__device__ __forceinline__ KmerRegs kmer_at(const SM &s, int q, uint32_t k) {
    const int t = (q >> 5) + L0_CTX, i = q & 31;
    const uint64_t kmask = ~0ull >> (64 - k);
    const uint32_t sh = 31 - i;
    const uint32_t cp = 65u - k, cs = cp >> 5, cb = cp & 31;
    const int g = t - 2 + (int)cs;
    KmerRegs r;
    r.f0 = (((uint64_t)fsr(s.F0[t - 1], s.F0[t - 2], sh) << 32) | fsr(s.F0[t], s.F0[t - 1], sh)) & kmask;
    r.f1 = (((uint64_t)fsr(s.F1[t - 1], s.F1[t - 2], sh) << 32) | fsr(s.F1[t], s.F1[t - 1], sh)) & kmask;
    const uint32_t ra0 = rplane(s.F0, g), ra1 = rplane(s.F0, g + 1), ra2 = rplane(s.F0, g + 2), ra3 = rplane(s.F0, g + 3);
    const uint32_t rb0 = rplane(s.F1, g), rb1 = rplane(s.F1, g + 1), rb2 = rplane(s.F1, g + 2), rb3 = rplane(s.F1, g + 3);
    const uint32_t q00 = fsr(ra0, ra1, cb), q01 = fsr(ra1, ra2, cb), q02 = fsr(ra2, ra3, cb);
    const uint32_t q10 = fsr(rb0, rb1, cb), q11 = fsr(rb1, rb2, cb), q12 = fsr(rb2, rb3, cb);
    r.r0 = (((uint64_t)fsr(q01, q02, i) << 32) | fsr(q00, q01, i)) & kmask;
    r.r1 = (((uint64_t)fsr(q11, q12, i) << 32) | fsr(q10, q11, i)) & kmask;
    return r;
}
// full 64-bit hash and strand of key index q (shmmrutils.rs:485-496)
template <class SM>
__device__ __forceinline__ uint64_t hash_at(const SM &s, int q, uint32_t k, uint32_t &strand) {
    const KmerRegs r = kmer_at(s, q, k);
    const bool rev = r.r0 < r.f0;
    strand = rev ? 1u : 0u;
    return rev ? (u64hash(r.r0) ^ u64hash(r.r1 ^ HASH_XOR)) : (u64hash(r.f0) ^ u64hash(r.f1 ^ HASH_XOR));
}
// the same for the K > 32 kernels with the arithmetic of the fast key loop (top word X of each strand, 32-bit halves,
// IMAD.WIDE hash): phase 7 runs it once per selected position, ~100 instructions instead of ~200
template <int K>
__device__ __forceinline__ uint64_t hash_at_fast(const L0Smem &s, int q, uint32_t &strand) {
    constexpr uint32_t PS = K - 32, HS = 64 - K;
    constexpr uint32_t cp = 65u - K, cs = cp >> 5, cb = cp & 31;
    const int t = (q >> 5) + L0_CTX, i = q & 31;
    const uint32_t sh = 31 - i;
    const int g = t - 2 + (int)cs;
    // plane 0, both strands: (X, lo) orders like the K-bit register
    const uint32_t a0 = s.F0[t], a1 = s.F0[t - 1], a2 = s.F0[t - 2];
    const uint32_t ra0 = rplane(s.F0, g), ra1 = rplane(s.F0, g + 1), ra2 = rplane(s.F0, g + 2), ra3 = rplane(s.F0, g + 3);
    const uint32_t q00 = fsr(ra0, ra1, cb), q01 = fsr(ra1, ra2, cb), q02 = fsr(ra2, ra3, cb);
    const uint32_t f0lo = fsr(a0, a1, sh), f0x = fsr(fsr(a0, a1, PS), fsr(a1, a2, PS), sh);
    const uint32_t r0lo = fsr(q00, q01, i), r0x = fsr(fsr(q00, q01, PS), fsr(q01, q02, PS), i);
    const bool rev = (r0x < f0x) || (r0x == f0x && r0lo < f0lo);   // rmmer.0 < fmmer.0 (shmmrutils.rs:486)
    strand = rev ? 1u : 0u;
    uint32_t ulo = rev ? r0lo : f0lo, uhi = (rev ? r0x : f0x) >> HS;
    // plane 1 of the chosen strand only
    uint32_t x0, x1, x2, amt;
    if (rev) {
        const uint32_t rb0 = rplane(s.F1, g), rb1 = rplane(s.F1, g + 1), rb2 = rplane(s.F1, g + 2), rb3 = rplane(s.F1, g + 3);
        x0 = fsr(rb0, rb1, cb); x1 = fsr(rb1, rb2, cb); x2 = fsr(rb2, rb3, cb); amt = (uint32_t)i;
    } else {
        x0 = s.F1[t]; x1 = s.F1[t - 1]; x2 = s.F1[t - 2]; amt = sh;
    }
    uint32_t vlo = fsr(x0, x1, amt) ^ (uint32_t)HASH_XOR, vhi = fsr(fsr(x0, x1, PS), fsr(x1, x2, PS), amt) >> HS;
    u64hash_dev32(ulo, uhi);
    u64hash_dev32(vlo, vhi);
    return ((uint64_t)(uhi ^ vhi) << 32) | (ulo ^ vlo);
}
// exact compare x[qa] < x[qb] (x = hash << 8 | k): 32-bit prefix first, the remaining 24 bits only on a prefix tie
__device__ __noinline__ bool key_lt_slow(const L0Smem &s, int qa, int qb, uint32_t k) {
    uint32_t st;
    const uint64_t a = hash_at(s, qa, k, st) << 8, b = hash_at(s, qb, k, st) << 8;
    return a < b;
}
__device__ __forceinline__ bool key_lt(const L0Smem &s, int qa, int qb, uint32_t k) {
    const uint32_t ha = s.H[pidx(qa)], hb = s.H[pidx(qb)];
    if (ha != hb) return ha < hb;
    return key_lt_slow(s, qa, qb, k);
}
// exact tie-inclusive selection test of key index q at sequence position pos (valid window starts [a_lo, a_hi])
__device__ __noinline__ bool selected_exact(const L0Smem &s, int q, int pos, int w, int a_lo, int a_hi, uint32_t k) {
    const int maxl = min(w - 1, pos - a_lo), maxr = min(w - 1, (a_hi + w - 1) - pos);
    if (maxl < 0 || maxr < 0) return false;
    int l = 0, r = 0;
    while (l < maxl && !key_lt(s, q - l - 1, q, k)) l++;
    while (r < maxr && !key_lt(s, q + r + 1, q, k)) r++;
    return l + r + 1 >= w;
}

// tile -> descriptor (executed by one thread).  The sequence of the newest descriptor is kept in shared memory
// (DescCache): tiles are handed out in order, so the next tile nearly always belongs to the same sequence and needs no
// global load at all — the three dependent loads of a fresh look-up (~1.5 us) sat in front of a barrier of every tile.
// A new sequence is found by walking forward from the cached one; the first tile of a CTA takes a binary search.
__device__ __forceinline__ void make_tile_desc(const L0Params &p, uint32_t tile, uint32_t w, L0Smem::TileDesc &d, DescCache &c) {
    if (c.sid == 0xFFFFFFFFu || tile >= c.tp_hi) {
        uint32_t sid;
        if (c.sid != 0xFFFFFFFFu) {
            sid = c.sid + 1;
            while (p.tile_prefix[sid + 1] <= tile) sid++;   // sequences without tiles (L <= k) are skipped
        } else {
            // (sequence, tile index): largest sid with tile_prefix[sid] <= tile
            uint32_t lo = 0, hi = p.n_seq;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (p.tile_prefix[mid] <= tile) lo = mid; else hi = mid;
            }
            sid = lo;
        }
        c.sid = sid; c.tp_lo = p.tile_prefix[sid]; c.tp_hi = p.tile_prefix[sid + 1]; c.len = p.len[sid]; c.off = p.off[sid];
    }
    const uint32_t sid = c.sid;
    const uint32_t j = tile - c.tp_lo;
    const uint32_t L = c.len;
    const uint32_t nt = c.tp_hi - c.tp_lo;
    int32_t ks = (int32_t)(j * p.tile_stride) - (int32_t)p.halo;
    const bool last = (j + 1 == nt);
    if (last) {  // the tail replay needs keys and selections back to E - 2w: pull the tile back if it is short
        const int32_t need = ((int32_t)L - 3 * (int32_t)w - 32) & ~31;
        if (need < ks) ks = need;
        if (ks < -(int32_t)p.halo) ks = -(int32_t)p.halo;
    }
    d.seq_id = sid; d.seq_len = L; d.keys_start = ks; d.is_last = last ? 1u : 0u;
    d.out_lo = (int32_t)(j * p.tile_stride);
    d.out_hi = (int32_t)min((uint64_t)L, (uint64_t)(j + 1) * p.tile_stride);
    d.seq_off = c.off;
    d.bad = 0; d.n_tail = 0; d.any_reject = 0; d.n_pre = 0; d.n_post = 0;
}

// cold path: mark the 32-base block at sequence position blk_pos as disturbed (idempotent; halo blocks are marked by
// two tiles); the neighbourhood is re-derived by cluster_replay_kernel
__device__ __noinline__ void mark_block(uint32_t *bits, uint32_t *n_marks, uint64_t seq_off, int32_t blk_pos) {
    const uint64_t g = (seq_off + (uint64_t)(int64_t)blk_pos) >> 5;
    atomicOr(&bits[g >> 5], 1u << (g & 31));
    atomicAdd(n_marks, 1u);
}

// key positions of block t whose k-byte window holds an invalid byte (i0 = invalid mask of the block, i1 / i2 = of the
// one / two blocks before it): their keys are not the reference's (stale k-mer) and are taken from the replay
__device__ __noinline__ uint32_t dirty_mask(uint32_t i0, uint32_t i1, uint32_t i2, uint32_t k) {
    const unsigned __int128 v = ((unsigned __int128)i0 << 64) | ((unsigned __int128)i1 << 32) | i2;
    const unsigned __int128 km = (((unsigned __int128)1) << k) - 1;
    uint32_t d = 0;
    for (uint32_t i = 0; i < 32; i++) if ((v >> (65 + i - k)) & km) d |= 1u << i;   // bytes [64+i-k+1, 64+i]
    return d;
}

// cold path of the fast key loop: some position of this thread's block has equal top words of plane 0 on both strands, so the
// strand picked from them may be wrong: redo the prefix of those positions exactly.  Returns true (and marks the block) if
// one of them is a pushed palindrome: the whole block is then re-derived by the replay and none of its keys matters.
__device__ __noinline__ bool strand_tie_scan(const L0Params &p, L0Smem &s, L0Smem::TileDesc &D, int kb, int32_t blk_pos, int32_t L, uint32_t k, uint32_t dirty) {
    for (int i = 0; i < 32; i++) {
        if ((dirty >> i) & 1u) continue;
        const int q = 32 * kb + i, pos = blk_pos + i;
        const KmerRegs r = kmer_at(s, q, k);
        if ((uint32_t)(r.f0 >> (k - 32)) != (uint32_t)(r.r0 >> (k - 32))) continue;
        if (r.f0 == r.r0 && r.f1 == r.r1 && pos >= (int)k && pos < L) {
            mark_block(p.mark_bits, p.n_marks, D.seq_off, blk_pos);
            return true;
        }
        const bool rev = r.r0 < r.f0;
        const uint64_t h = rev ? (u64hash(r.r0) ^ u64hash(r.r1 ^ HASH_XOR)) : (u64hash(r.f0) ^ u64hash(r.f1 ^ HASH_XOR));
        s.H[pidx(q)] = (uint32_t)(h >> 32) & 0x00FFFFFFu;   // the fast loop's 24-bit prefix
    }
    return false;
}

// Level-0 minimizers of one tile.  W, K > 0 are compile-time specialisations; 0 = read from params.
template <int W, int K, int U = 8>
__global__ void __launch_bounds__(L0_NT, L0_MIN_CTAS) l0_kernel(const L0Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    L0Smem &s = *reinterpret_cast<L0Smem *>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t w = W ? (uint32_t)W : p.w;
    const uint32_t k = K ? (uint32_t)K : p.k;
    const uint64_t kmask = ~0ull >> (64 - k);
    const uint32_t mlo = (uint32_t)kmask, mhi = (uint32_t)(kmask >> 32);

    uint16_t *const list = reinterpret_cast<uint16_t *>(s.P);
    const uint32_t t_begin = p.cta_tile[blockIdx.x], t_end = p.cta_tile[blockIdx.x + 1];
    uint64_t running = 0;  // entries this CTA has produced so far (identical in every thread)
    pgr_mm128 *chunk = p.arena + (uint64_t)blockIdx.x * p.chunk_cap;

    // spare blocks of the exchange arrays are never written with real data; give them harmless values once
    for (int i = tid; i < L0_KB + 2 * L0_PADB; i += L0_NT) s.bext[i] = 0;
    if (tid < (int)(sizeof(s.wsum) / sizeof(uint32_t))) s.wsum[tid] = 0;   // padding entries stay 0
    for (int i = tid; i < L0_ARR; i += L0_NT) { s.H[i] = 0; s.P[i] = 0; }

    // first tile: descriptor + this thread's 32 bases
    if (tid == 0) { s.dc.sid = 0xFFFFFFFFu; if (t_begin < t_end) make_tile_desc(p, t_begin, w, s.td[0], s.dc); }
    __syncthreads();
    uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
#if PGR_L0_BULK
    uint4 *const tile_bytes = reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(s.P) + L0_SLOT_OFF);
    __shared__ alignas(8) uint64_t tile_bar;
    uint32_t bar_phase = 0;
    if (tid == 0) {   // the zero fill of P above is complete (barrier): the slot may be written by the copy engine
        mbar_init(&tile_bar, 1);
        // the tile's load region starts L0_CTX blocks before key index 0; the store keeps 16 KiB of readable slack on both sides
        if (t_begin < t_end) bulk_load(tile_bytes, p.seq + s.td[0].seq_off + (int64_t)(s.td[0].keys_start - 32 * L0_CTX), L0_NT * 32, &tile_bar);
    }
    __syncthreads();
#else
    if (t_begin < t_end) {
        const int32_t bp = s.td[0].keys_start + 32 * (tid - L0_CTX);
        if (bp + 32 > 0 && bp < (int32_t)s.td[0].seq_len) {
            const uint4 *src = reinterpret_cast<const uint4 *>(p.seq + s.td[0].seq_off + bp);
            v0 = __ldg(src); v1 = __ldg(src + 1);
        }
    }
#endif
    int cur = 0;
    for (uint32_t tile = t_begin; tile < t_end; ++tile, cur ^= 1) {
        L0Smem::TileDesc &D = s.td[cur];
        const bool has_next = tile + 1 < t_end;
        if (tid == 0 && has_next) {
            make_tile_desc(p, tile + 1, w, s.td[cur ^ 1], s.dc);   // overlaps with phases 1-2 of this tile
#if PGR_L0_BULK && PGR_L0_PREFETCH_L2
            // the bulk copy of the next tile is issued after phase 3 of this one and consumed right after phase 7: have its bytes
            // in the L2 by then
            bulk_prefetch_l2(p.seq + s.td[cur ^ 1].seq_off + (int64_t)(s.td[cur ^ 1].keys_start - 32 * L0_CTX), L0_NT * 32);
#endif
        }
        const int32_t L = (int32_t)D.seq_len;
        const int32_t keys_start = D.keys_start;
        const int32_t blk_pos = keys_start + 32 * (tid - L0_CTX);  // sequence position of this thread's first base

        // ---- phase 1: 32 bases -> plane words -------------------------------------------------------------
        uint32_t f0 = 0, f1 = 0, inv = 0;
        const bool blk_live = (blk_pos + 32 > 0) && (blk_pos < L);  // block intersects the sequence
#if PGR_L0_BULK
        mbar_wait(&tile_bar, bar_phase);      // this tile's bytes have landed
        bar_phase ^= 1u;
        v0 = tile_bytes[2 * tid]; v1 = tile_bytes[2 * tid + 1];
#endif
        if (blk_live) {
            const uint32_t wd[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            uint32_t bad_bits = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t a, b;
                planes4v(wd[j], a, b, bad_bits);
                f0 = __funnelshift_l(a, f0, 4);      // (f0 << 4) | (a >> 28)
                f1 = __funnelshift_l(b, f1, 4);
            }
            bad_bits &= 0xDFDFDFDFu;                 // lower case is valid
            if (bad_bits || blk_pos < 0 || blk_pos + 32 > L) {
                // slow exact check restricted to the bytes that belong to the sequence: a block with a byte outside ACGTacgt
                // is marked (its neighbourhood is replayed exactly), like a palindrome's
                uint32_t minv = 0;   // bytes the reference's LUT maps to 4 (raw codes 0..3 are bases for it, shmmrutils.rs:426)
#pragma unroll 1
                for (int j = 0; j < 32; j++) {
                    const int pos = blk_pos + j;
                    const uint32_t ch = (wd[j >> 2] >> (8 * (j & 3))) & 0xFF;
                    if (pos >= 0 && pos < L && !byte_is_acgt(ch)) { inv |= 1u << j; if (ch > 3) minv |= 1u << j; }
                }
                if (inv) {
                    mark_block(p.mark_bits, p.n_marks, D.seq_off, blk_pos);
                    if (minv == 0xFFFFFFFFu) { const uint64_t g = (D.seq_off + (uint64_t)(int64_t)blk_pos) >> 5; atomicOr(&p.allinv_bits[g >> 5], 1u << (g & 31)); }
                }
            }
        }
        s.F0[tid] = f0; s.F1[tid] = f1;
        s.bext[tid + 2] = inv;   // invalid-byte mask of block tid, read by the two following blocks below; bext is rewritten in phase 3
        if (tid < 8) { s.F0[L0_NT + tid] = 0; s.F1[L0_NT + tid] = 0; }
        __syncthreads();   // planes visible; also publishes thread 0's descriptor of the next tile
#if !PGR_L0_BULK
        if (has_next) {    // the loads fly while the key loop runs
            const L0Smem::TileDesc &N = s.td[cur ^ 1];
            const int32_t bp = N.keys_start + 32 * (tid - L0_CTX);
            if (bp + 32 > 0 && bp < (int32_t)N.seq_len) {
                const uint4 *src = reinterpret_cast<const uint4 *>(p.seq + N.seq_off + bp);
                v0 = __ldg(src); v1 = __ldg(src + 1);
            }
        }
#endif

        // ---- phase 2: keys for blocks CTX.. ------------------------------------------------------------------
        const int kb = tid - L0_CTX;  // key block index (negative for the two context threads)
        uint32_t dirty = 0;           // key positions of this block whose k-byte window holds an invalid byte
        if (tid >= L0_CTX && blk_live) {
            {
                const uint32_t i1 = s.bext[tid + 1], i2 = s.bext[tid];
                if (inv | i1 | i2) dirty = (inv == 0xFFFFFFFFu) ? 0xFFFFFFFFu : dirty_mask(inv, i1, i2, k);
            }
            const uint32_t a2 = s.F0[tid - 2], a1 = s.F0[tid - 1], a0 = f0;
            const uint32_t b2 = s.F1[tid - 2], b1 = s.F1[tid - 1], b0 = f1;
            // r-planes: Q = G >> (32*(tid-2) + 65 - k), so that rmmer(i) = (Q >> i) & kmask
            const uint32_t cp = 65u - k, cs = cp >> 5, cb = cp & 31;
            const int g = tid - 2 + (int)cs;
            const uint32_t ra0 = rplane(s.F0, g), ra1 = rplane(s.F0, g + 1), ra2 = rplane(s.F0, g + 2), ra3 = rplane(s.F0, g + 3);
            const uint32_t rb0 = rplane(s.F1, g), rb1 = rplane(s.F1, g + 1), rb2 = rplane(s.F1, g + 2), rb3 = rplane(s.F1, g + 3);
            const uint32_t q00 = fsr(ra0, ra1, cb), q01 = fsr(ra1, ra2, cb), q02 = fsr(ra2, ra3, cb);
            const uint32_t q10 = fsr(rb0, rb1, cb), q11 = fsr(rb1, rb2, cb), q12 = fsr(rb2, rb3, cb);
            const int base = pidx(32 * kb);
            if (dirty == 0xFFFFFFFFu) {
                // interior of an invalid run: no key of this block is the reference's; none is a candidate
                for (int i = 0; i < 32; i++) s.H[base + i] = (K > 32) ? 0x00FFFFFFu : 0xFFFFFFFFu;
            } else if constexpr (K > 32) {
                // Fast key loop (K > 32; 61 instructions per position against 72 for the generic loop below).  Each
                // strand carries X = the TOP 32 bits of its K-bit plane-0 register (bits [K-32, K)), a plain funnel
                // extraction from the plane string pre-shifted by K-32 once per thread.  The strand (shmmrutils.rs:486) is
                // decided on X alone: one 32-bit compare + one min; the low word of plane 0 and both words of plane 1
                // are extracted for the chosen strand only (unconditional forward extraction, predicated overwrite), and
                // the high words are X >> (64-K), so no masks are needed.  Blocks in which some position has equal X on
                // both strands (every palindrome is one) are redone exactly after the loop (strand_tie_scan, cold).
                constexpr uint32_t PS = K - 32, HS = 64 - K;
                const uint32_t a0p = fsr(a0, a1, PS), a1p = fsr(a1, a2, PS);
                const uint32_t b0p = fsr(b0, b1, PS), b1p = fsr(b1, b2, PS);
                const uint32_t q0p0 = fsr(q00, q01, PS), q0p1 = fsr(q01, q02, PS);
                const uint32_t q1p0 = fsr(q10, q11, PS), q1p1 = fsr(q11, q12, PS);
                const uint64_t m1 = p.m1;
                const HiShift hsc = p.hs;
                bool no_tie = true;   // stays true unless some position of the block has X_f == X_r
#pragma unroll U
                for (int i = 0; i < 32; i++) {
                    // forward window: bits [31-i, 63-i) of (a1:a0) = left funnel by i+1 (clamped: 32 at i = 31), so that
                    // positions i and i+1 share the shift-amount register i+1 (reverse window: right funnel by i)
                    const uint32_t sl = (uint32_t)i + 1u;
                    const uint32_t f0x = __funnelshift_lc(a0p, a1p, sl), r0x = fsr(q0p0, q0p1, i);
                    no_tie &= (f0x != r0x);
                    const uint32_t ux = min(f0x, r0x);
                    uint32_t ulo, vlo, vx;
                    asm("{\n\t.reg .pred p;\n\t"
                        "setp.lt.u32 p, %3, %4;\n\t"
                        "shf.l.clamp.b32 %0, %11, %12, %18;\n\t@p shf.r.wrap.b32 %0, %5, %6, %17;\n\t"
                        "shf.l.clamp.b32 %1, %13, %14, %18;\n\t@p shf.r.wrap.b32 %1, %7, %8, %17;\n\t"
                        "shf.l.clamp.b32 %2, %15, %16, %18;\n\t@p shf.r.wrap.b32 %2, %9, %10, %17;\n\t}"
                        : "=&r"(ulo), "=&r"(vlo), "=&r"(vx)
                        : "r"(r0x), "r"(f0x), "r"(q00), "r"(q01), "r"(q10), "r"(q11), "r"(q1p0), "r"(q1p1),
                          "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"(b0p), "r"(b1p), "r"((uint32_t)i), "r"(sl));
                    uint32_t uhi, vhi;
                    if ((PGR_L0_HIMASK & 8) != 0) { uhi = __umulhi(ux, hsc.chs); vhi = __umulhi(vx, hsc.chs); }
                    else { uhi = ux >> HS; vhi = vx >> HS; }
                    vlo ^= (uint32_t)HASH_XOR;
                    u64hash_dev32x(ulo, uhi, m1, hsc);
                    u64hash_dev32x(vlo, vhi, m1, hsc);
                    // key prefix of this kernel: the top 24 bits of MM128.x = hash bits 32..55 (one LOP3); any prefix of x
                    // orders consistently with x, and prefix ties between candidates are resolved exactly in phase 5
                    s.H[base + i] = (uhi ^ vhi) & 0x00FFFFFFu;
                }
                // a block with a pushed palindrome lies inside a replay patch as a whole (patch_kernels.cuh): none of its
                // positions needs to be a candidate, which spares the exact tie tests of phase 5 in low-complexity sequence
                if (!no_tie && strand_tie_scan(p, s, D, kb, blk_pos, L, k, dirty)) dirty = 0xFFFFFFFFu;
                if (dirty) {   // never candidates; the maximal prefix keeps them from shadowing clean keys
                    for (int i = 0; i < 32; i++) if ((dirty >> i) & 1u) s.H[base + i] = 0x00FFFFFFu;
                }
            } else {
                bool pal_blk = false;   // the block holds a pushed palindrome: replayed as a whole, no candidates (see above)
    #pragma unroll U
                for (int i = 0; i < 32; i++) {
                    const uint32_t sh = 31 - i;
                    const uint32_t f0lo = fsr(a0, a1, sh) & mlo, f0hi = fsr(a1, a2, sh) & mhi;
                    const uint32_t f1lo = fsr(b0, b1, sh) & mlo, f1hi = fsr(b1, b2, sh) & mhi;
                    const uint32_t r0lo = fsr(q00, q01, i) & mlo, r0hi = fsr(q01, q02, i) & mhi;
                    const uint32_t r1lo = fsr(q10, q11, i) & mlo, r1hi = fsr(q11, q12, i) & mhi;
                    if (f0lo == r0lo) {                             // rare: possible palindrome (shmmrutils.rs:477)
                        const int pos = blk_pos + i;
                        if (f0hi == r0hi && f1lo == r1lo && f1hi == r1hi && pos >= (int)k && pos < L && !((dirty >> i) & 1u) && !pal_blk) {
                            mark_block(p.mark_bits, p.n_marks, D.seq_off, blk_pos);
                            pal_blk = true;
                        }
                    }
                    // strand: reverse iff rmmer.0 < fmmer.0 (shmmrutils.rs:486, plane 0 only); one 64-bit compare, four selects
                    uint32_t ulo, uhi, vlo, vhi;
                    asm("{\n\t.reg .pred p;\n\t.reg .b64 a, b;\n\t"
                        "mov.b64 a, {%4, %5};\n\tmov.b64 b, {%6, %7};\n\tsetp.lt.u64 p, a, b;\n\t"
                        "selp.b32 %0, %4, %6, p;\n\tselp.b32 %1, %5, %7, p;\n\tselp.b32 %2, %8, %10, p;\n\tselp.b32 %3, %9, %11, p;\n\t}"
                        : "=r"(ulo), "=r"(uhi), "=r"(vlo), "=r"(vhi)
                        : "r"(r0lo), "r"(r0hi), "r"(f0lo), "r"(f0hi), "r"(r1lo), "r"(r1hi), "r"(f1lo), "r"(f1hi));
                    vlo ^= (uint32_t)HASH_XOR;
                    u64hash_dev32(ulo, uhi);
                    u64hash_dev32(vlo, vhi);
                    // MM128.x high word = hash bits 24..55
                    s.H[base + i] = ((dirty >> i) & 1u) ? 0xFFFFFFFFu : __funnelshift_r(ulo ^ vlo, uhi ^ vhi, 24);
                }
                if (pal_blk) {
                    dirty = 0xFFFFFFFFu;
                    for (int i = 0; i < 32; i++) s.H[base + i] = 0xFFFFFFFFu;
                }
            }
        }
        __syncthreads();

        // ---- phase 3: window-minimum candidates on the 32-bit key prefix ----------------------------------
        // valid window starts a (sequence positions): [k, Eb - w], Eb = min(L, L - w + k)
        const int32_t Eb = min(L, L - (int32_t)w + (int32_t)k);
        const int32_t a_lo = (int32_t)k, a_hi = Eb - (int32_t)w;            // inclusive bounds
        uint32_t cand = 0, omask = 0, bmask = 0;
        {   // every thread runs this phase (threads 0,1 work on spare blocks) so that barriers stay uniform
            const int base = pidx(32 * kb);
            const int c = (int)w - 1, d = c >> 5, cr = c & 31;
            // masks of this block: valid window starts (amask), positions that lie in some valid window (vmask), and
            // the part of those inside / below this tile's output range (omask / bmask).  Candidates are collected over
            // vmask, halo included: a halo candidate can tie on the key prefix with a candidate of the output range
            uint32_t amask = 0, vmask = 0;
            {
                const int lo = max(a_lo - blk_pos, 0), hi = min(a_hi - blk_pos, 31);
                if (lo <= hi) amask = (0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo);
                const int vlo = max(a_lo - blk_pos, 0), vhi = min(a_hi + (int)w - 1 - blk_pos, 31);
                if (vlo <= vhi) vmask = (0xFFFFFFFFu >> (31 - vhi)) & (0xFFFFFFFFu << vlo);
                const int plo = max(D.out_lo - blk_pos, 0), phi = min(D.out_hi - 1 - blk_pos, 31);
                if (plo <= phi) omask = (0xFFFFFFFFu >> (31 - phi)) & (0xFFFFFFFFu << plo);
                const int bhi = min(D.out_lo - 1 - blk_pos, 31);
                if (bhi >= 0) bmask = 0xFFFFFFFFu >> (31 - bhi);
            }
            if (tid < L0_CTX) vmask = 0;
            if (w > 32) {
                // van Herk / Gil-Werman with one 32-position block per thread
                uint32_t m[32];
                {   // prefix minima -> smem, suffix minima -> registers
                    uint32_t run = 0xFFFFFFFFu;
#pragma unroll
                    for (int o = 0; o < 32; o++) { m[o] = s.H[base + o]; run = min(run, m[o]); s.P[base + o] = run; }
                    run = 0xFFFFFFFFu;
#pragma unroll
                    for (int o = 31; o >= 0; o--) { run = min(run, m[o]); m[o] = run; }
                    s.bext[kb + L0_PADB] = run;
                }
                __syncthreads();
                {
                    uint32_t mid1 = 0xFFFFFFFFu;
                    for (int b = 1; b < d; b++) mid1 = min(mid1, s.bext[kb + L0_PADB + b]);
                    const uint32_t mid2 = min(mid1, s.bext[kb + L0_PADB + d]);
                    const int pb = pidx(32 * (kb + d)) + cr;   // padded index of P_{kb+d}[cr]
#pragma unroll
                    for (int o = 0; o < 32; o++) {
                        const bool carry = (o + cr) >= 32;
                        const uint32_t far = s.P[pb + o + (carry ? 1 : 0)];
                        m[o] = min3u(m[o], carry ? mid2 : mid1, far);
                    }
                    if (amask != 0xFFFFFFFFu) {   // blocks at the sequence ends: invalid windows contribute nothing to the max
#pragma unroll
                        for (int o = 0; o < 32; o++) m[o] = ((amask >> o) & 1u) ? m[o] : 0u;
                    }
                }
                __syncthreads();
                {   // suffix maxima of m -> smem; block max
                    uint32_t run = 0;
#pragma unroll
                    for (int o = 31; o >= 0; o--) { run = max(run, m[o]); s.P[base + o] = run; }
                    s.bext[kb + L0_PADB] = run;
                }
                __syncthreads();
                if (tid >= L0_CTX) {   // the two context threads have no blocks to their left inside the arrays
                    uint32_t mid1 = 0;
                    for (int b = 1; b < d; b++) mid1 = max(mid1, s.bext[kb + L0_PADB - b]);
                    const uint32_t mid2 = max(mid1, s.bext[kb + L0_PADB - d]);
                    const int sb = pidx(32 * (kb - d)) - cr;
                    uint32_t run = 0;
#pragma unroll
                    for (int o = 0; o < 32; o++) {
                        const bool borrow = o < cr;
                        run = max(run, m[o]);
                        const uint32_t far = s.P[sb + o - (borrow ? 1 : 0)];
                        const uint32_t M = max3u(run, borrow ? mid2 : mid1, far);
                        if (M == s.H[base + o]) cand |= 1u << o;
                    }
                }
                cand &= vmask & ~dirty;
            } else {
                // small windows: direct evaluation (w <= 32); w is uniform, so no barrier mismatch with the branch above
                for (int o = 0; o < 32; o++) {
                    if (!(((vmask & ~dirty) >> o) & 1u)) continue;
                    const int q = 32 * kb + o;
                    const uint32_t hq = s.H[pidx(q)];
                    // l / r = run of neighbours with H >= hq inside the valid position range (and inside the tile's keys:
                    // a halo position may be under-counted, which only matters for windows that leave the tile)
                    const int pos = blk_pos + o;
                    const int maxl = min(min((int)w - 1, pos - a_lo), q), maxr = min(min((int)w - 1, (a_hi + (int)w - 1) - pos), L0_KPOS - 1 - q);
                    int l = 0, r = 0;
                    while (l < maxl && s.H[pidx(q - l - 1)] >= hq) l++;
                    while (r < maxr && s.H[pidx(q + r + 1)] >= hq) r++;
                    if (l + r + 1 >= (int)w) cand |= 1u << o;
                }
            }
        }

        // ---- phase 4: ordered list of the candidates (halo included) -----------------------------------------------
        // block-wide scan of the candidate counts; the few candidates below / above the output range (halo blocks only)
        // are counted with shared-memory atomics: the list is position-ordered, so the output range is list[n_pre, n_pre + n_in)
        const uint32_t cnt = __popc(cand);
        {
            const uint32_t c_pre = __popc(cand & bmask), c_post = __popc(cand & ~(omask | bmask));
            if (c_pre) atomicAdd(&D.n_pre, c_pre);
            if (c_post) atomicAdd(&D.n_post, c_post);
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, dlt);
            if (lane >= dlt) incl += v;
        }
        if (lane == 31) s.wsum[warp] = incl;
        __syncthreads();
#if PGR_L0_BULK
        if (has_next && tid == 0) {   // phase 3 is over for every thread (barrier above): P's tail is free until phase 3 of the next tile
            const L0Smem::TileDesc &N = s.td[cur ^ 1];
            bulk_load(tile_bytes, p.seq + N.seq_off + (int64_t)(N.keys_start - 32 * L0_CTX), L0_NT * 32, &tile_bar);
        }
#endif
        uint32_t wbase = 0, total = 0;
        {
            const uint4 *wv = reinterpret_cast<const uint4 *>(s.wsum);
#pragma unroll
            for (int i = 0; i < (L0_NT / 32 + 3) / 4; i++) {
                const uint4 v = wv[i];
                const uint32_t e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; j++) { if (4 * i + j < warp) wbase += e[j]; total += e[j]; }
            }
        }
        const uint32_t lo_idx = D.n_pre;                       // list index of the first candidate in the output range
        const uint32_t n_in = total - D.n_pre - D.n_post;      // candidates in the output range
        uint32_t *const chv = s.P + L0_CH_OFF;
        const bool ch_ok = total <= (uint32_t)L0_CH_CAP;
        {
            uint32_t dst = wbase + incl - cnt, rem = cand;
            const int hb = pidx(32 * kb);
            while (rem) {
                const int o = __ffs(rem) - 1;
                rem &= rem - 1;
                list[dst] = (uint16_t)(32 * kb + o);
                if (ch_ok) chv[dst] = s.H[hb + o];
                dst++;
            }
        }
        __syncthreads();

        // ---- phase 5: candidates that tie on the key prefix with a neighbouring candidate: exact 64-bit test -----------
        // A candidate without such a tie is the strict minimum of one of its windows, hence selected.  One with a tie
        // (the other may sit in the halo) is tested exactly; a rejected one is marked (bit 15) and dropped below.
        for (uint32_t j = lo_idx + tid; j < lo_idx + n_in; j += L0_NT) {
            const int q = list[j] & 0x7FFF;
            bool tie = false;
            if (ch_ok) {
                // four neighbours per side and round, all loads independent (positions from the list, prefixes from chv); a
                // round whose farthest neighbour is still within w of q is followed by another one (rare)
                const uint32_t hq = chv[j];
                for (int jj = (int)j - 1; jj >= 0 && !tie;) {
                    const int i0 = jj, i1 = max(jj - 1, 0), i2 = max(jj - 2, 0), i3 = max(jj - 3, 0);
                    const int q0 = list[i0] & 0x7FFF, q1 = list[i1] & 0x7FFF, q2 = list[i2] & 0x7FFF, q3 = list[i3] & 0x7FFF;
                    const uint32_t h0 = chv[i0], h1 = chv[i1], h2 = chv[i2], h3 = chv[i3];
                    const bool n0 = q - q0 < (int)w, n1 = n0 && jj >= 1 && q - q1 < (int)w, n2 = n1 && jj >= 2 && q - q2 < (int)w,
                               n3 = n2 && jj >= 3 && q - q3 < (int)w;
                    tie = (n0 && h0 == hq) || (n1 && h1 == hq) || (n2 && h2 == hq) || (n3 && h3 == hq);
                    if (!n3) break;
                    jj -= 4;
                }
                for (uint32_t jj = j + 1; jj < total && !tie;) {
                    const uint32_t last = total - 1;
                    const uint32_t i0 = jj, i1 = min(jj + 1, last), i2 = min(jj + 2, last), i3 = min(jj + 3, last);
                    const int q0 = list[i0] & 0x7FFF, q1 = list[i1] & 0x7FFF, q2 = list[i2] & 0x7FFF, q3 = list[i3] & 0x7FFF;
                    const uint32_t h0 = chv[i0], h1 = chv[i1], h2 = chv[i2], h3 = chv[i3];
                    const bool n0 = q0 - q < (int)w, n1 = n0 && jj + 1 <= last && q1 - q < (int)w, n2 = n1 && jj + 2 <= last && q2 - q < (int)w,
                               n3 = n2 && jj + 3 <= last && q3 - q < (int)w;
                    tie = (n0 && h0 == hq) || (n1 && h1 == hq) || (n2 && h2 == hq) || (n3 && h3 == hq);
                    if (!n3) break;
                    jj += 4;
                }
            } else {
                const uint32_t hq = s.H[pidx(q)];
                for (int jj = (int)j - 1; jj >= 0 && q - (int)(list[jj] & 0x7FFF) < (int)w && !tie; jj--) tie = (s.H[pidx(list[jj] & 0x7FFF)] == hq);
                for (uint32_t jj = j + 1; jj < total && (int)(list[jj] & 0x7FFF) - q < (int)w && !tie; jj++) tie = (s.H[pidx(list[jj] & 0x7FFF)] == hq);
            }
            if (tie && !selected_exact(s, q, q + keys_start, (int)w, a_lo, a_hi, k)) { list[j] = (uint16_t)(q | 0x8000); D.any_reject = 1u; }
        }
        __syncthreads();
        uint16_t *const olist = list + lo_idx;   // the output range is a contiguous part of the position-ordered list
        uint32_t n_list = n_in;
        if (D.any_reject) {   // rare: drop the marked entries
            if (tid == 0) {
                uint32_t o = 0;
                for (uint32_t j = 0; j < n_in; j++) if (!(olist[j] & 0x8000)) olist[o++] = olist[j];
                s.n_list = o;
            }
            __syncthreads();
            n_list = s.n_list;
        }

        // ---- phase 6: tail replay (last tile of the sequence; shmmrutils.rs:503-515 with rule (2) disabled) ---
        if (D.is_last && (int32_t)w > (int32_t)k && L > (int32_t)k) {
            if (tid == 0) {
                // q = last selected position below Eb (selections of the whole sequence, not only this tile's range)
                const int32_t t_lo = max(Eb, (int32_t)k);
                int32_t q = -1;
                for (int32_t pos = Eb - 1; pos >= a_lo && pos >= Eb - (int32_t)w; pos--) {
                    const int qq = pos - keys_start;
                    if (qq < 0) break;
                    if (selected_exact(s, qq, pos, (int)w, a_lo, a_hi, k)) { q = pos; break; }
                }
                uint32_t n = 0;
                for (int32_t pos = t_lo; pos < L; pos++) {
                    const bool fire = (q < 0) ? (pos == (int32_t)(k + w - 1)) : (pos == q + (int32_t)w);
                    if (!fire) continue;
                    const int qe = pos - keys_start, qa = qe - (int)w + 1;
                    int best = qa;
                    for (int j = qa + 1; j <= qe; j++) if (key_lt(s, j, best, k)) best = j;
                    for (int j = qa; j <= qe; j++) {
                        if (!key_lt(s, best, j, k)) {  // x[j] == min
                            if (lo_idx + n_list + n < (uint32_t)L0_LISTCAP) olist[n_list + n] = (uint16_t)j;
                            n++;
                            q = j + keys_start;
                        }
                    }
                }
                D.n_tail = n;
            }
            __syncthreads();
        }
        const uint32_t n_tail = D.n_tail;
        const uint32_t n_all = n_list + n_tail;

        // ---- phase 7: rebuild the full MM128 of every selected position and write it (coalesced) -------------
        for (uint32_t j = tid; j < min(n_all, (uint32_t)L0_LISTCAP - lo_idx); j += L0_NT) {
            const uint64_t dst = running + j;
            if (dst >= p.chunk_cap) break;
            const int q = olist[j];
            uint32_t strand;
            uint64_t h;
            if constexpr (K > 32) h = hash_at_fast<K>(s, q, strand); else h = hash_at(s, q, k, strand);
            pgr_mm128 mm;
            mm.x = (h << 8) | k;
            mm.y = ((uint64_t)D.seq_id << 32) | ((uint64_t)(uint32_t)(q + keys_start) << 1) | strand;
            chunk[dst] = mm;
        }
        if (tid == 0) {
            atomicAdd(&p.seq_count[D.seq_id], n_all);
            if (D.bad || lo_idx + n_all > (uint32_t)L0_LISTCAP) atomicOr(&p.seq_flag[D.seq_id], 1u);
        }
        running += n_all;
        __syncthreads();  // tile fully consumed before its smem is reused
    }
    if (tid == 0) p.chunk_count[blockIdx.x] = running;
}

// ---------------------------------------------------------------------------------------------------------------
// Exact sequential restatement of the level-0 loop (shmmrutils.rs:440-530) for flagged sequences: one thread per
// sequence, ring buffer in local memory.  mode 0 = count only, mode 1 = write at dst[seq_l0_off[s] + i].
struct ReplayParams {
    const uint8_t *seq; const uint64_t *off; const uint32_t *len;
    const uint32_t *list;       // [n_list] sequence ordinals to replay
    uint32_t n_list;
    uint32_t w, k;
    uint32_t *count;            // [n_seq] out (mode 0)
    const uint64_t *dst_off;    // [n_seq] (mode 1)
    pgr_mm128 *dst;
};

// LUT of shmmrutils.rs:426-436 without a branch (the sequential kernels call it once per base from lanes that hold different
// bases: a switch diverges five ways): the code bits of a letter are bit 1 of c ^ (c >> 1) and bit 2 of c (A 0, C 1, G 2, T 3,
// either case); the byte is that letter iff it equals the expected upper-case letter once the case bit is dropped; raw 0..3 are
// codes themselves; anything else maps to 4.
__device__ __forceinline__ uint32_t base_code(uint32_t c) {
    const uint32_t code = (((c ^ (c >> 1)) >> 1) & 1u) | ((c >> 1) & 2u);
    const uint32_t expect = (0x54474341u >> (8u * code)) & 0xFFu;   // 'A', 'C', 'G', 'T'
    return c < 4u ? c : ((c & 0xDFu) == expect ? code : 4u);
}

template <int MODE>
__global__ void replay_l0_kernel(const ReplayParams p) {
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= p.n_list) return;
    const uint32_t sid = p.list[li];
    const uint8_t *sq = p.seq + p.off[sid];
    const uint64_t L = p.len[sid];
    const uint32_t w = p.w, k = p.k;
    const uint64_t mask = ~0ull >> (64 - k);
    const uint32_t shift = k - 1;
    uint64_t rx[128]; uint32_t ry[128];   // ring buffer (w <= 128): x and (pos<<1|strand)
    for (uint32_t i = 0; i < w; i++) { rx[i] = ~0ull; ry[i] = ~0u; }
    uint32_t r_start = 0, r_end = 0, r_len = 0;
    uint64_t f0 = 0, f1 = 0, r0 = 0, r1 = 0;
    uint64_t min_x = ~0ull, mdist = 0, n_out = 0;
    pgr_mm128 *dst = MODE ? p.dst + p.dst_off[sid] : nullptr;
    const uint64_t rule2_end = L - (uint64_t)w + (uint64_t)k;  // wrapping, shmmrutils.rs:518
    for (uint64_t pos = 0; pos < L; pos++) {
        const uint32_t c = base_code(sq[pos]);
        if (c < 4) {
            f0 = ((f0 << 1) | (c & 1)) & mask;
            f1 = ((f1 << 1) | (c >> 1)) & mask;
            const uint64_t rc = 3 ^ c;
            r0 = ((r0 >> 1) | ((rc & 1) << shift)) & mask;
            r1 = ((r1 >> 1) | ((rc >> 1) << shift)) & mask;
        }
        if (f0 == r0 && f1 == r1) continue;
        if (pos < k) continue;
        const bool rev = r0 < f0;
        const uint64_t h = rev ? (u64hash(r0) ^ u64hash(r1 ^ HASH_XOR)) : (u64hash(f0) ^ u64hash(f1 ^ HASH_XOR));
        const uint64_t mx = (h << 8) | k;
        const uint32_t my = ((uint32_t)pos << 1) | (rev ? 1u : 0u);
        rx[r_end] = mx; ry[r_end] = my;
        r_end = (r_end + 1) % w;
        if (r_len < w) r_len++; else r_start = (r_start + 1) % w;
        if (mdist == (uint64_t)(w - 1)) {
            uint64_t mn = ~0ull;
            for (uint32_t i = 0; i < r_len; i++) if (rx[i] < mn) mn = rx[i];
            uint32_t last_y = 0;
            for (uint32_t i = 0; i < w; i++) {
                const uint32_t sl = (r_start + i) % w;
                if (rx[sl] == mn) {
                    if (MODE) { pgr_mm128 mm; mm.x = rx[sl]; mm.y = ((uint64_t)sid << 32) | ry[sl]; dst[n_out] = mm; }
                    n_out++;
                    last_y = ry[sl];
                }
            }
            min_x = mn;
            mdist = pos - (uint64_t)(last_y >> 1);
            continue;
        } else if (mx <= min_x && pos >= (uint64_t)(w + k) && pos < rule2_end && pos < L) {
            if (MODE) { pgr_mm128 mm; mm.x = mx; mm.y = ((uint64_t)sid << 32) | my; dst[n_out] = mm; }
            n_out++;
            min_x = mx;
            mdist = 0;
            continue;
        }
        mdist++;
    }
    if (!MODE) p.count[sid] = (uint32_t)n_out;
}

// ---------------------------------------------------------------------------------------------------------------
// chunked arena -> flat level-0 list.  Entry i of chunk c has logical (fast-path) index chunk_prefix[c] + i; it goes
// to flat[seq_dst[s] + (logical - seq_fast[s])] unless its sequence was flagged for replay.
struct GatherParams {
    const pgr_mm128 *arena; uint64_t chunk_cap;
    const uint64_t *chunk_prefix;   // [grid+1]
    uint32_t n_chunks;
    const uint64_t *seq_fast;       // [n_seq] exclusive scan of fast-path counts
    const uint64_t *seq_dst;        // [n_seq+1] exclusive scan of final counts
    const uint32_t *seq_flag;
    pgr_mm128 *flat;
};

__global__ void gather_l0_kernel(const GatherParams p) {
    const uint32_t c = blockIdx.y;
    const uint64_t n = p.chunk_prefix[c + 1] - p.chunk_prefix[c];
    const uint64_t base = p.chunk_prefix[c];
    const pgr_mm128 *src = p.arena + (uint64_t)c * p.chunk_cap;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const pgr_mm128 mm = src[i];
        const uint32_t sid = (uint32_t)(mm.y >> 32);
        if (p.seq_flag[sid]) continue;
        p.flat[p.seq_dst[sid] + (base + i - p.seq_fast[sid])] = mm;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Levels 1 and 2 (reduce_shmmr, shmmrutils.rs:359-415) and the min_span filter (:536-555) on the flat list.
// y>>32 holds the sequence ordinal; seq_off[] are the list offsets of each sequence.
// flag kinds: 0 = window-minimum rule with window r (padding = virtual MAX sentinels), 1 = min_span filter.
struct LevelParams {
    const pgr_mm128 *in; uint64_t n_in;
    const uint64_t *seq_off_in;  // [n_seq+1]
    uint32_t n_seq;
    uint32_t r, padding, min_span;
    uint8_t *flags;              // [n_in]
    uint32_t *block_sum;         // [n_blocks]
    uint64_t *block_prefix;      // [n_blocks+1]
    pgr_mm128 *out;
    uint64_t *seq_off_out;       // [n_seq+1]
    const uint32_t *rid;         // final pass only: ordinal -> caller's rid
    uint32_t patch_rid;
};

constexpr int LV_NT = 256, LV_PER = 8, LV_BLK = LV_NT * LV_PER;

template <int KIND>
__global__ void __launch_bounds__(LV_NT) level_flags_kernel(const LevelParams p) {
    __shared__ uint32_t wsum[LV_NT / 32];
    const uint64_t i0 = (uint64_t)blockIdx.x * LV_BLK;
    uint32_t cnt = 0;
    for (int j = 0; j < LV_PER; j++) {
        const uint64_t i = i0 + (uint64_t)j * LV_NT + threadIdx.x;
        if (i >= p.n_in) break;
        const pgr_mm128 me = p.in[i];
        const uint32_t sid = (uint32_t)(me.y >> 32);
        const uint64_t b = p.seq_off_in[sid], e = p.seq_off_in[sid + 1];
        bool keep;
        if (KIND == 0) {
            const uint32_t r = p.r;
            uint32_t l = 0, rr = 0;
            for (uint32_t d = 1; d < r; d++) {
                if (i < b + d) { if (p.padding) l = r - 1; break; }
                if (p.in[i - d].x < me.x) break;
                l = d;
            }
            for (uint32_t d = 1; d < r; d++) {
                if (i + d >= e) { if (p.padding) rr = r - 1; break; }
                if (p.in[i + d].x < me.x) break;
                rr = d;
            }
            keep = (l + rr + 1 >= r);
        } else {
            if (i == b || i + 1 == e) {
                keep = true;
            } else {
                const pgr_mm128 pv = p.in[i - 1], nx = p.in[i + 1];
                const uint32_t pp = (uint32_t)(pv.y & 0xFFFFFFFFu) >> 1, cp = (uint32_t)(me.y & 0xFFFFFFFFu) >> 1,
                               np = (uint32_t)(nx.y & 0xFFFFFFFFu) >> 1;
                keep = (uint32_t)(cp - pp) > p.min_span && (uint32_t)(np - cp) > p.min_span && pv.x != me.x && me.x != nx.x;
            }
        }
        p.flags[i] = keep ? 1 : 0;
        cnt += keep ? 1u : 0u;
    }
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_down_sync(0xFFFFFFFFu, cnt, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < LV_NT / 32; i++) t += wsum[i];
        p.block_sum[blockIdx.x] = t;
    }
}

// single-CTA exclusive scan of block sums (n_blocks = n_in / 2048): every thread sums one contiguous chunk, the 1024
// chunk sums are scanned once, every thread writes its chunk's prefixes (two barriers in all)
__global__ void __launch_bounds__(1024) block_scan_kernel(const uint32_t *block_sum, uint64_t *block_prefix, uint32_t n_blocks) {
    __shared__ uint64_t wtot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t per = (n_blocks + 1023) / 1024;
    const uint32_t b = min(n_blocks, threadIdx.x * per), e = min(n_blocks, b + per);
    uint64_t sum = 0;
    for (uint32_t i = b; i < e; i++) sum += block_sum[i];
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
        wtot[lane] = wi - v;
    }
    __syncthreads();
    uint64_t run = wtot[warp] + incl - sum;
    for (uint32_t i = b; i < e; i++) { block_prefix[i] = run; run += block_sum[i]; }
    if (threadIdx.x == 1023) block_prefix[n_blocks] = run;   // the last thread's running sum is the total (empty chunks add 0)
}

__global__ void __launch_bounds__(LV_NT) level_scatter_kernel(const LevelParams p) {
    __shared__ uint32_t excl[LV_BLK + 1];
    __shared__ uint32_t wsum[LV_NT / 32];
    const uint64_t i0 = (uint64_t)blockIdx.x * LV_BLK;
    const uint64_t bp = p.block_prefix[blockIdx.x];
    // thread t owns elements i0 + t*LV_PER .. +LV_PER (contiguous, so ranks are in order)
    uint8_t f[LV_PER];
    uint32_t cnt = 0;
    const uint64_t my0 = i0 + (uint64_t)threadIdx.x * LV_PER;
#pragma unroll
    for (int j = 0; j < LV_PER; j++) {
        f[j] = (my0 + j < p.n_in) ? p.flags[my0 + j] : 0;
        cnt += f[j];
    }
    uint32_t incl = cnt;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((threadIdx.x & 31) >= d) incl += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t wb = 0;
    for (uint32_t j = 0; j < (threadIdx.x >> 5); j++) wb += wsum[j];
    uint32_t rank = wb + incl - cnt;
#pragma unroll
    for (int j = 0; j < LV_PER; j++) {
        excl[threadIdx.x * LV_PER + j] = rank;
        if (f[j]) {
            pgr_mm128 mm = p.in[my0 + j];
            if (p.patch_rid) mm.y = ((uint64_t)p.rid[(uint32_t)(mm.y >> 32)] << 32) | (mm.y & 0xFFFFFFFFull);
            p.out[bp + rank] = mm;
            rank++;
        }
    }
    if (threadIdx.x == LV_NT - 1) excl[LV_BLK] = rank;
    __syncthreads();
    // sequences whose first input element lies in this block get their output offset from the local scan
    const uint64_t i1 = min(i0 + (uint64_t)LV_BLK, p.n_in);
    const bool last_block = (i1 == p.n_in);
    // first sid with seq_off_in[sid] >= i0
    uint32_t lo = 0, hi = p.n_seq + 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (p.seq_off_in[mid] < i0) lo = mid + 1; else hi = mid;
    }
    for (uint32_t sid = lo + threadIdx.x; sid <= p.n_seq; sid += LV_NT) {
        const uint64_t b = p.seq_off_in[sid];
        if (b < i1 || (last_block && b == i1)) p.seq_off_out[sid] = bp + excl[b - i0]; else break;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Tiled versions of the two kernels above (the default path).  The list is read with coalesced 16-byte loads into a
// shared-memory tile with a halo of r-1 entries; sequence boundaries come from the ordinal in y>>32 of the neighbours
// (the list is grouped by sequence), so there are no per-entry offset look-ups; the scatter assigns consecutive
// entries to consecutive lanes (ballot ranks) so that its loads and stores coalesce.  The input may also be the
// level-0 ARENA itself (n_chunks > 0): logical index i lives in chunk c with chunk_prefix[c] <= i < chunk_prefix[c+1] at
// arena[c * chunk_cap + (i - chunk_prefix[c])], which saves the gather pass when no sequence needs patching.
struct ChunkView {
    const uint64_t *prefix;        // [n_chunks+1] logical index of each chunk's first entry; nullptr = flat list
    uint64_t cap;                  // entries per chunk
    uint32_t n_chunks;
    const uint32_t *block_chunk;   // [n_blocks] chunk of the first index a block touches (block_chunk_kernel): without it
                                   // every thread would start with a binary search of dependent loads
};
struct ChunkCursor {          // per-thread cache of the chunk that held the last index
    uint32_t c; uint64_t lo, hi;
};
__device__ __forceinline__ void cursor_seek(const ChunkView &v, ChunkCursor &k, uint64_t i) {
    uint32_t lo = 0, hi = v.n_chunks;   // largest c with prefix[c] <= i, skipping empty chunks (prefix[c+1] > i)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (v.prefix[mid] <= i) lo = mid; else hi = mid;
    }
    k.c = lo; k.lo = v.prefix[lo]; k.hi = v.prefix[lo + 1];
}
__device__ __forceinline__ uint64_t physical_index(const ChunkView &v, ChunkCursor &k, uint64_t i) {
    if (v.prefix == nullptr) return i;
    if (i < k.lo || i >= k.hi) cursor_seek(v, k, i);
    return (uint64_t)k.c * v.cap + (i - k.lo);
}
__device__ __forceinline__ pgr_mm128 load_mm(const pgr_mm128 *p) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p);   // entries are 16-byte aligned
    pgr_mm128 m; m.x = v.x; m.y = v.y;
    return m;
}
__device__ __forceinline__ void store_mm(pgr_mm128 *p, const pgr_mm128 &m) {
    *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(m.x, m.y);
}

__device__ __forceinline__ ChunkCursor cursor_for_block(const ChunkView &v, uint32_t block) {
    ChunkCursor k = {0, 1, 0};   // empty range: the first access seeks (flat lists never look at it)
    if (v.prefix != nullptr) { k.c = v.block_chunk[block]; k.lo = v.prefix[k.c]; k.hi = v.prefix[k.c + 1]; }
    return k;
}

constexpr int LT_HALO = 12;   // r - 1 <= 11
constexpr uint32_t LT_NOSEQ = 0xFFFFFFFFu;

// chunk of the first logical index block b of the tiled kernels touches (its left halo)
__global__ void block_chunk_kernel(const ChunkView v, uint32_t n_blocks, uint64_t n_in, uint32_t *block_chunk) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint64_t i0 = (uint64_t)b * LV_BLK;
    uint64_t i = i0 >= (uint64_t)LT_HALO ? i0 - LT_HALO : 0;
    if (i >= n_in) i = n_in - 1;
    ChunkCursor k;
    cursor_seek(v, k, i);
    block_chunk[b] = k.c;
}

template <int KIND>
__global__ void __launch_bounds__(LV_NT) level_flags_tiled_kernel(const LevelParams p, const ChunkView cv) {
    __shared__ uint64_t sx[LV_BLK + 2 * LT_HALO];
    __shared__ uint32_t ssid[LV_BLK + 2 * LT_HALO];
    __shared__ uint32_t spos[KIND == 1 ? LV_BLK + 2 * LT_HALO : 1];
    __shared__ uint32_t wsum[LV_NT / 32];
    const uint64_t i0 = (uint64_t)blockIdx.x * LV_BLK;
    ChunkCursor cur = cursor_for_block(cv, blockIdx.x);
    for (int t = threadIdx.x; t < LV_BLK + 2 * LT_HALO; t += LV_NT) {
        const int64_t i = (int64_t)i0 - LT_HALO + t;
        uint64_t x = 0; uint32_t sid = LT_NOSEQ, pos = 0;
        if (i >= 0 && (uint64_t)i < p.n_in) {
            const pgr_mm128 m = load_mm(p.in + physical_index(cv, cur, (uint64_t)i));
            x = m.x; sid = (uint32_t)(m.y >> 32); pos = (uint32_t)(m.y & 0xFFFFFFFFu) >> 1;
        }
        sx[t] = x; ssid[t] = sid;
        if (KIND == 1) spos[t] = pos;
    }
    __syncthreads();
    uint32_t cnt = 0;
    const int r = (int)p.r;
#pragma unroll
    for (int j = 0; j < LV_PER; j++) {
        const int e = j * LV_NT + threadIdx.x;       // element of the block
        const uint64_t i = i0 + e;
        if (i >= p.n_in) break;
        const int t = e + LT_HALO;
        const uint64_t x = sx[t];
        const uint32_t sid = ssid[t];
        bool keep;
        if (KIND == 0) {
            int l = 0, rr = 0;
            for (int d = 1; d < r; d++) {
                if (ssid[t - d] != sid) { if (p.padding) l = r - 1; break; }   // left end of the sequence
                if (sx[t - d] < x) break;
                l = d;
            }
            for (int d = 1; d < r; d++) {
                if (ssid[t + d] != sid) { if (p.padding) rr = r - 1; break; }
                if (sx[t + d] < x) break;
                rr = d;
            }
            keep = (l + rr + 1 >= r);
        } else {
            if (ssid[t - 1] != sid || ssid[t + 1] != sid) keep = true;           // first / last of its sequence
            else {
                const uint32_t pp = spos[t - 1], cp = spos[t], np = spos[t + 1];
                keep = (uint32_t)(cp - pp) > p.min_span && (uint32_t)(np - cp) > p.min_span && sx[t - 1] != x && x != sx[t + 1];
            }
        }
        p.flags[i] = keep ? 1 : 0;
        cnt += keep ? 1u : 0u;
    }
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_down_sync(0xFFFFFFFFu, cnt, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < LV_NT / 32; i++) t += wsum[i];
        p.block_sum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(LV_NT) level_scatter_tiled_kernel(const LevelParams p, const ChunkView cv) {
    __shared__ uint32_t excl[LV_BLK + 1];                  // exclusive rank of every element of the block
    __shared__ uint32_t wcnt[LV_PER][LV_NT / 32];          // kept entries per (row, warp)
    const uint64_t i0 = (uint64_t)blockIdx.x * LV_BLK;
    const uint64_t bp = p.block_prefix[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t masks[LV_PER];
    uint8_t f[LV_PER];
#pragma unroll
    for (int j = 0; j < LV_PER; j++) {
        const uint64_t i = i0 + (uint64_t)j * LV_NT + threadIdx.x;
        f[j] = (i < p.n_in) ? p.flags[i] : 0;
        masks[j] = __ballot_sync(0xFFFFFFFFu, f[j] != 0);
        if (lane == 0) wcnt[j][warp] = __popc(masks[j]);
    }
    __syncthreads();
    if (threadIdx.x < 32) {   // exclusive scan of the LV_PER * (LV_NT/32) counts in element order (row-major), in place
        constexpr int NC = LV_PER * (LV_NT / 32);
        uint32_t *flat = &wcnt[0][0];
        uint32_t carry = 0;
        for (int base = 0; base < NC; base += 32) {
            const uint32_t v = (base + lane < NC) ? flat[base + lane] : 0;
            uint32_t incl = v;
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
            if (base + lane < NC) flat[base + lane] = carry + incl - v;
            carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        if (lane == 0) excl[LV_BLK] = carry;
    }
    __syncthreads();
    ChunkCursor cur = cursor_for_block(cv, blockIdx.x);
#pragma unroll
    for (int j = 0; j < LV_PER; j++) {
        const int e = j * LV_NT + threadIdx.x;
        const uint32_t rank = wcnt[j][warp] + __popc(masks[j] & ((1u << lane) - 1u));
        excl[e] = rank;
        if (f[j]) {
            pgr_mm128 mm = load_mm(p.in + physical_index(cv, cur, i0 + e));
            if (p.patch_rid) mm.y = ((uint64_t)p.rid[(uint32_t)(mm.y >> 32)] << 32) | (mm.y & 0xFFFFFFFFull);
            store_mm(p.out + bp + rank, mm);
        }
    }
    __syncthreads();
    // sequences whose first input element lies in this block get their output offset from the local scan
    const uint64_t i1 = min(i0 + (uint64_t)LV_BLK, p.n_in);
    const bool last_block = (i1 == p.n_in);
    uint32_t lo = 0, hi = p.n_seq + 1;   // first sid with seq_off_in[sid] >= i0
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (p.seq_off_in[mid] < i0) lo = mid + 1; else hi = mid;
    }
    for (uint32_t sid = lo + threadIdx.x; sid <= p.n_seq; sid += LV_NT) {
        const uint64_t b = p.seq_off_in[sid];
        if (b < i1 || (last_block && b == i1)) p.seq_off_out[sid] = bp + excl[b - i0]; else break;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Single-pass version of one level (flags + ordered compaction in ONE kernel): the list is read from HBM once, the tile
// and a halo of r-1 entries live in shared memory, and the global output offset of a tile comes from a decoupled
// look-back over per-tile status words (aggregate / inclusive prefix), tiles being handed out in order by a ticket.
constexpr int LF_NT = 256, LF_PER = 8, LF_TILE = LF_NT * LF_PER, LF_HALO = 12;
constexpr int LF_N = LF_TILE + 2 * LF_HALO;

struct LevelFusedParams {
    const pgr_mm128 *in; uint64_t n_in;
    const uint64_t *seq_off_in;  // [n_seq+1]
    uint32_t n_seq;
    uint32_t r, padding, min_span;
    pgr_mm128 *out;
    uint64_t *seq_off_out;       // [n_seq+1]
    const uint32_t *rid; uint32_t patch_rid;
    unsigned long long *tile_status;   // [n_tiles] zeroed: bits 62..63 = 0 invalid / 1 aggregate / 2 inclusive prefix
    uint32_t *ticket;                  // zeroed
    uint64_t *total_out;
};

struct LevelFusedSmem {
    uint64_t x[LF_N + LF_N / 8 + 8];   // index q + q/8: one pad per 8 entries keeps the 8-per-thread accesses conflict-light
    uint64_t y[LF_N + LF_N / 8 + 8];
    uint32_t excl[LF_TILE + 1];
    uint32_t wsum[LF_NT / 32];
    uint32_t tile;
    uint64_t base;
};
__device__ __forceinline__ int lf_idx(int q) { return q + (q >> 3); }

template <int KIND>
__global__ void __launch_bounds__(LF_NT) level_fused_kernel(const LevelFusedParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LevelFusedSmem &s = *reinterpret_cast<LevelFusedSmem *>(smem_raw);
    if (threadIdx.x == 0) s.tile = atomicAdd(p.ticket, 1u);
    __syncthreads();
    const uint32_t tile = s.tile;
    const int64_t i0 = (int64_t)tile * LF_TILE;
    const int64_t n = (int64_t)p.n_in;
    // tile + halo -> shared memory; q = 0 is global index i0 - LF_HALO
    for (int q = threadIdx.x; q < LF_N; q += LF_NT) {
        const int64_t g = i0 - LF_HALO + q;
        pgr_mm128 m; m.x = 0; m.y = ~0ull;       // sequence ordinal 0xFFFFFFFF never matches: acts as a boundary
        if (g >= 0 && g < n) m = p.in[g];
        s.x[lf_idx(q)] = m.x; s.y[lf_idx(q)] = m.y;
    }
    __syncthreads();
    const uint32_t r = p.r;
    uint32_t fl = 0, cnt = 0;
    const int q0 = LF_HALO + threadIdx.x * LF_PER;
#pragma unroll
    for (int j = 0; j < LF_PER; j++) {
        const int q = q0 + j;
        const int64_t g = i0 - LF_HALO + q;
        if (g >= n) break;
        const uint64_t xe = s.x[lf_idx(q)], ye = s.y[lf_idx(q)];
        const uint32_t sq = (uint32_t)(ye >> 32);
        bool keep;
        if (KIND == 0) {
            uint32_t l = 0, rr = 0;
            for (uint32_t d = 1; d < r; d++) {
                if ((uint32_t)(s.y[lf_idx(q - d)] >> 32) != sq) { if (p.padding) l = r - 1; break; }
                if (s.x[lf_idx(q - d)] < xe) break;
                l = d;
            }
            for (uint32_t d = 1; d < r; d++) {
                if ((uint32_t)(s.y[lf_idx(q + d)] >> 32) != sq) { if (p.padding) rr = r - 1; break; }
                if (s.x[lf_idx(q + d)] < xe) break;
                rr = d;
            }
            keep = (l + rr + 1 >= r);
        } else {
            const uint64_t yp = s.y[lf_idx(q - 1)], yn = s.y[lf_idx(q + 1)];
            if ((uint32_t)(yp >> 32) != sq || (uint32_t)(yn >> 32) != sq) {
                keep = true;   // first or last shimmer of its sequence
            } else {
                const uint32_t pp = (uint32_t)(yp & 0xFFFFFFFFu) >> 1, cp = (uint32_t)(ye & 0xFFFFFFFFu) >> 1, np = (uint32_t)(yn & 0xFFFFFFFFu) >> 1;
                keep = (uint32_t)(cp - pp) > p.min_span && (uint32_t)(np - cp) > p.min_span && s.x[lf_idx(q - 1)] != xe && xe != s.x[lf_idx(q + 1)];
            }
        }
        if (keep) { fl |= 1u << j; cnt++; }
    }
    // CTA scan of the per-thread counts
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = cnt;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s.wsum[warp] = incl;
    __syncthreads();
    uint32_t wb = 0, total = 0;
    for (int i = 0; i < LF_NT / 32; i++) { const uint32_t v = s.wsum[i]; if (i < warp) wb += v; total += v; }
    // decoupled look-back (warp 0): publish the aggregate, then sum predecessors until an inclusive prefix is met
    if (warp == 0) {
        const unsigned long long FLAG_A = 1ull << 62, FLAG_P = 2ull << 62, VMASK = (1ull << 62) - 1;
        volatile unsigned long long *st = p.tile_status;
        if (lane == 0) { st[tile] = (tile == 0 ? FLAG_P : FLAG_A) | (unsigned long long)total; __threadfence(); }
        uint64_t excl = 0;
        if (tile > 0) {
            int64_t look = (int64_t)tile - 1;
            for (;;) {
                const int64_t j = look - lane;
                unsigned long long v = FLAG_P;   // lanes past the beginning read as "prefix 0"
                if (j >= 0) { do { v = st[j]; } while ((v >> 62) == 0); }
                const uint32_t is_p = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2);
                // sum the values of lanes up to and including the first prefix lane
                const int first_p = is_p ? __ffs(is_p) - 1 : 32;
                uint64_t val = (lane <= first_p) ? (uint64_t)(v & VMASK) : 0;
                for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xFFFFFFFFu, val, d);
                val = __shfl_sync(0xFFFFFFFFu, val, 0);
                excl += val;
                if (is_p) break;
                look -= 32;
            }
        }
        if (lane == 0) {
            st[tile] = FLAG_P | (unsigned long long)(excl + total);
            __threadfence();
            s.base = excl;
            if (i0 + LF_TILE >= n) *p.total_out = excl + total;
        }
    }
    __syncthreads();
    const uint64_t base = s.base;
    uint32_t rank = wb + incl - cnt;
#pragma unroll
    for (int j = 0; j < LF_PER; j++) {
        s.excl[threadIdx.x * LF_PER + j] = rank;
        if ((fl >> j) & 1u) {
            const int q = q0 + j;
            pgr_mm128 mm; mm.x = s.x[lf_idx(q)]; mm.y = s.y[lf_idx(q)];
            if (p.patch_rid) mm.y = ((uint64_t)p.rid[(uint32_t)(mm.y >> 32)] << 32) | (mm.y & 0xFFFFFFFFull);
            p.out[base + rank] = mm;
            rank++;
        }
    }
    if (threadIdx.x == LF_NT - 1) s.excl[LF_TILE] = rank;
    __syncthreads();
    // sequences whose first input element lies in this tile get their output offset from the local scan
    const int64_t i1 = min(i0 + (int64_t)LF_TILE, n);
    const bool last_tile = (i1 == n);
    uint32_t lo = 0, hi = p.n_seq + 1;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((int64_t)p.seq_off_in[mid] < i0) lo = mid + 1; else hi = mid; }
    for (uint32_t sid = lo + threadIdx.x; sid <= p.n_seq; sid += LF_NT) {
        const int64_t b = (int64_t)p.seq_off_in[sid];
        if (b < i1 || (last_tile && b == i1)) p.seq_off_out[sid] = base + s.excl[b - i0]; else break;
    }
}

}  // namespace pgr
