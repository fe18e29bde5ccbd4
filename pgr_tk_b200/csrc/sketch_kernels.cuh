// sketch_kernels.cuh — sketch mode of sequence_to_shmmrs (sequence_to_shmmrs2, shmmrutils.rs:558-630): every pushed position
// whose 64-bit hash is below u64::MAX >> 4 >> r is kept; the min_span filter follows (run_level kind 1).
//
// There is no window in this mode, so a position depends on its k-mer only and the work is tile-parallel:
//   sketch_mask_kernel   one CTA per tile of 254 32-base blocks (+2 context blocks), one thread per block: 32-byte load ->
//                        bit planes in shared memory (as l0_kernel's phase 1) -> for a CLEAN block (no byte outside ACGTacgt in
//                        the block or in the two before it) 32 k-mers from funnel shifts of the plane words, exact strand
//                        and palindrome tests, two hashes, threshold -> a keep mask and a strand mask per block (8 B per 32
//                        bases).  A block that is not clean is marked in a bitmap (and in the all-invalid bitmap when no byte
//                        of it is a base for the reference's LUT).
//   sketch_dirty_kernel  the marked blocks, one thread per bitmap word: the reference's state machine on the block's 32
//                        positions, started from the last k valid bases before the block; runs of invalid bytes are skipped
//                        through the all-invalid bitmap, a word (1024 bases) at a time, so that a Mb-scale gap costs its
//                        length / 1024 per block and not its length.
//   sketch_count_kernel  kept positions per tile (a warp per tile) -> host scan -> tile offsets
//   sketch_write_kernel  one CTA per tile again: planes, block-wide scan of the mask populations, the full MM128 of the kept
//                        positions (hash_at on the plane words for clean blocks, the state machine for marked ones).
#pragma once
#include "patch_kernels.cuh"

namespace pgr {

constexpr int SKT_KPOS = L0_KB * 32;   // positions per tile

struct SketchTParams {
    const uint8_t *seq; const uint64_t *off; const uint32_t *len;
    const uint32_t *tile_prefix;     // [n_seq+1] cumulative tile count
    uint32_t n_seq, n_tiles;
    uint32_t k, r;
    uint64_t blk_base;               // store block of keep[0] / strand[0]
    uint32_t *keep, *strand;         // per store block of the chunk
    uint32_t *dirty_bits, *allinv_bits, *n_dirty;   // bitmaps over the store's blocks (absolute block index)
    uint32_t *tile_count;            // [n_tiles]
    const uint64_t *tile_off;        // [n_tiles+1]
    pgr_mm128 *out;
};

// plane words of one tile (the part of L0Smem this mode needs: 3 KB, so that the SM holds its full complement of threads)
struct SketchSmem {
    uint32_t F0[L0_NT + 8], F1[L0_NT + 8];
    uint32_t bext[L0_NT + 2];          // invalid-byte mask of block t at [t + 2]
    uint32_t wsum[L0_NT / 32];
    uint32_t sid, j;
};

__device__ __forceinline__ uint64_t sketch_threshold(uint32_t r) { return (~0ull >> 4) >> r; }

// largest block index g' in [g_lo, g] whose all-invalid bit is clear, or g_lo - 1 (as int64) when there is none
__device__ __forceinline__ int64_t skip_allinv_back(const uint32_t *bits, int64_t g_lo, int64_t g) {
    while (g >= g_lo) {
        const uint32_t word = bits[g >> 5];
        const uint32_t upto = (uint32_t)(g & 31);
        uint32_t clear = ~word & (0xFFFFFFFFu >> (31 - upto));   // blocks <= g of this word that are not all-invalid
        if (clear) {
            const int64_t cand = (g & ~(int64_t)31) + (31 - __clz(clear));
            return cand >= g_lo ? cand : g_lo - 1;
        }
        g = (g & ~(int64_t)31) - 1;
    }
    return g_lo - 1;
}

// the reference's registers right before sequence position `upto`: built from the last k bases the LUT accepts, found by
// walking back (most recent base = bit 0 of the forward planes, bit k-1 of the complement planes)
__device__ __forceinline__ void sketch_regs_before(const uint8_t *sq, int64_t upto, uint32_t k, const uint32_t *allinv, uint64_t blk0,
                                                   uint64_t &f0, uint64_t &f1, uint64_t &r0, uint64_t &r1) {
    f0 = f1 = r0 = r1 = 0;
    uint32_t got = 0;
    int64_t b = upto;
    while (b > 0 && got < k) {
        if ((b & 31) == 0) {   // at a block boundary: jump over the all-invalid blocks below it
            const int64_t blk = skip_allinv_back(allinv, (int64_t)blk0, (int64_t)blk0 + (b >> 5) - 1) - (int64_t)blk0;
            b = (blk + 1) << 5;
            if (b <= 0) break;
        }
        b--;
        const uint32_t c = base_code(sq[b]);
        if (c < 4) {
            f0 |= (uint64_t)(c & 1) << got; f1 |= (uint64_t)(c >> 1) << got;
            const uint64_t rc = 3 ^ c;
            r0 |= (rc & 1) << (k - 1 - got); r1 |= (rc >> 1) << (k - 1 - got);
            got++;
        }
    }
}

// the state machine of sequence_to_shmmrs2 over the 32 positions of one block; emit(i, hash, reverse) for the kept ones
template <class Emit>
__device__ __forceinline__ void sketch_walk_block(const uint8_t *sq, int64_t L, int64_t blk_pos, uint32_t k, uint64_t thr, const uint32_t *allinv,
                                                  uint64_t blk0, Emit emit) {
    const uint64_t mask = ~0ull >> (64 - k);
    const uint32_t shift = k - 1;
    uint64_t f0, f1, r0, r1;
    sketch_regs_before(sq, max(blk_pos, (int64_t)0), k, allinv, blk0, f0, f1, r0, r1);
    for (int i = 0; i < 32; i++) {
        const int64_t pos = blk_pos + i;
        if (pos < 0 || pos >= L) continue;
        const uint32_t c = base_code(sq[pos]);
        if (c < 4) {
            f0 = ((f0 << 1) | (c & 1)) & mask; f1 = ((f1 << 1) | (c >> 1)) & mask;
            const uint64_t rc = 3 ^ c;
            r0 = ((r0 >> 1) | ((rc & 1) << shift)) & mask; r1 = ((r1 >> 1) | ((rc >> 1) << shift)) & mask;
        }
        if (f0 == r0 && f1 == r1) continue;
        if (pos < (int64_t)k) continue;
        const bool rev = r0 < f0;
        const uint64_t h = u64hash_dev(rev ? r0 : f0) ^ u64hash_dev((rev ? r1 : f1) ^ HASH_XOR);   // one pair of hashes on the selected strand, no divergent branch
        if (h < thr) emit(i, h, rev);
    }
}

// tile -> (sequence, tile index in it); one thread, result through shared memory
__device__ __forceinline__ void sketch_tile_lookup(const SketchTParams &p, uint32_t tile, uint32_t &sid, uint32_t &j) {
    uint32_t lo = 0, hi = p.n_seq;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (p.tile_prefix[mid] <= tile) lo = mid; else hi = mid; }
    sid = lo; j = tile - p.tile_prefix[lo];
}

// phase 1 shared by the mask and the write kernel: this thread's 32 bases -> plane words in shared memory; returns the mask of
// bytes outside ACGTacgt (inv) and of bytes the reference's LUT maps to no base (minv), restricted to the sequence
__device__ __forceinline__ void sketch_planes(SketchSmem &s, const uint8_t *sq, int64_t L, int64_t blk_pos, int tid, uint32_t &inv, uint32_t &minv) {
    uint32_t f0 = 0, f1 = 0;
    inv = 0; minv = 0;
    if (blk_pos + 32 > 0 && blk_pos < L) {
        const uint4 *src = reinterpret_cast<const uint4 *>(sq + blk_pos);   // the store keeps 16 KiB of readable slack on both sides
        const uint4 v0 = __ldg(src), v1 = __ldg(src + 1);
        const uint32_t wd[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        uint32_t bad_bits = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint32_t a, b;
            planes4v(wd[j], a, b, bad_bits);
            f0 = __funnelshift_l(a, f0, 4);
            f1 = __funnelshift_l(b, f1, 4);
        }
        bad_bits &= 0xDFDFDFDFu;
        if (bad_bits || blk_pos < 0 || blk_pos + 32 > L) {
#pragma unroll 1
            for (int j = 0; j < 32; j++) {
                const int64_t pos = blk_pos + j;
                const uint32_t ch = (wd[j >> 2] >> (8 * (j & 3))) & 0xFF;
                if (pos >= 0 && pos < L && !byte_is_acgt(ch)) { inv |= 1u << j; if (ch > 3) minv |= 1u << j; }
            }
        }
    }
    s.F0[tid] = f0; s.F1[tid] = f1;
    s.bext[tid + 2] = inv;
    if (tid < 8) { s.F0[L0_NT + tid] = 0; s.F1[L0_NT + tid] = 0; }
}

// K > 32: the key loop of l0_kernel's fast form (strand from the top 32 bits X of the plane-0 registers, plane 1 and the low
// words for the chosen strand only); a block in which some position has X_f == X_r — every palindrome is one — is redone by
// the generic loop, which tests strand and palindrome exactly.  K = 0: generic loop only.
template <int K>
__global__ void __launch_bounds__(L0_NT) sketch_mask_kernel(const SketchTParams p) {
    __shared__ SketchSmem s;
    const int tid = threadIdx.x;
    if (tid == 0) sketch_tile_lookup(p, blockIdx.x, s.sid, s.j);
    __syncthreads();
    const uint32_t sh_sid = s.sid, sh_j = s.j;
    const uint32_t sid = sh_sid;
    const int64_t L = p.len[sid];
    const uint64_t soff = p.off[sid];
    const uint8_t *sq = p.seq + soff;
    const uint32_t k = p.k;
    const int64_t blk_pos = (int64_t)sh_j * SKT_KPOS + 32 * (tid - L0_CTX);
    uint32_t inv, minv;
    sketch_planes(s, sq, L, blk_pos, tid, inv, minv);
    __syncthreads();
    if (tid < L0_CTX || blk_pos >= L) return;   // context blocks belong to the previous tile
    const uint64_t g = (soff + (uint64_t)blk_pos) >> 5;
    const uint32_t i1 = s.bext[tid + 1], i2 = s.bext[tid];
    uint32_t keep = 0, strand = 0;
    if (inv | i1 | i2) {
        atomicOr(&p.dirty_bits[g >> 5], 1u << (g & 31));
        atomicAdd(p.n_dirty, 1u);
        if (minv == 0xFFFFFFFFu) atomicOr(&p.allinv_bits[g >> 5], 1u << (g & 31));
    } else {
        const uint64_t kmask = ~0ull >> (64 - k);
        const uint32_t mlo = (uint32_t)kmask, mhi = (uint32_t)(kmask >> 32);
        const uint32_t a2 = s.F0[tid - 2], a1 = s.F0[tid - 1], a0 = s.F0[tid];
        const uint32_t b2 = s.F1[tid - 2], b1 = s.F1[tid - 1], b0 = s.F1[tid];
        const uint32_t cp = 65u - k, cs = cp >> 5, cb = cp & 31;
        const int gq = tid - 2 + (int)cs;
        const uint32_t ra0 = rplane(s.F0, gq), ra1 = rplane(s.F0, gq + 1), ra2 = rplane(s.F0, gq + 2), ra3 = rplane(s.F0, gq + 3);
        const uint32_t rb0 = rplane(s.F1, gq), rb1 = rplane(s.F1, gq + 1), rb2 = rplane(s.F1, gq + 2), rb3 = rplane(s.F1, gq + 3);
        const uint32_t q00 = fsr(ra0, ra1, cb), q01 = fsr(ra1, ra2, cb), q02 = fsr(ra2, ra3, cb);
        const uint32_t q10 = fsr(rb0, rb1, cb), q11 = fsr(rb1, rb2, cb), q12 = fsr(rb2, rb3, cb);
        const uint64_t thr = sketch_threshold(p.r);
        bool fast_done = false;
        if constexpr (K > 32) {
            constexpr uint32_t PS = K - 32, HS = 64 - K;
            const uint32_t a0p = fsr(a0, a1, PS), a1p = fsr(a1, a2, PS);
            const uint32_t b0p = fsr(b0, b1, PS), b1p = fsr(b1, b2, PS);
            const uint32_t q0p0 = fsr(q00, q01, PS), q0p1 = fsr(q01, q02, PS);
            const uint32_t q1p0 = fsr(q10, q11, PS), q1p1 = fsr(q11, q12, PS);
            const uint32_t thr_hi = (uint32_t)(thr >> 32), thr_lo = (uint32_t)thr;
            bool no_tie = true;
            uint32_t kp = 0, sd = 0;
#pragma unroll 4
            for (int i = 0; i < 32; i++) {
                const uint32_t sl = (uint32_t)i + 1u;
                const uint32_t f0x = __funnelshift_lc(a0p, a1p, sl), r0x = fsr(q0p0, q0p1, i);
                no_tie &= (f0x != r0x);
                const bool rev = r0x < f0x;
                const uint32_t ux = min(f0x, r0x);
                uint32_t ulo = rev ? fsr(q00, q01, i) : __funnelshift_lc(a0, a1, sl);
                uint32_t vlo = (rev ? fsr(q10, q11, i) : __funnelshift_lc(b0, b1, sl)) ^ (uint32_t)HASH_XOR;
                const uint32_t vx = rev ? fsr(q1p0, q1p1, i) : __funnelshift_lc(b0p, b1p, sl);
                uint32_t uhi = ux >> HS, vhi = vx >> HS;
                u64hash_dev32(ulo, uhi);
                u64hash_dev32(vlo, vhi);
                const uint32_t hh = uhi ^ vhi, hl = ulo ^ vlo;
                if (hh < thr_hi || (hh == thr_hi && hl < thr_lo)) { kp |= 1u << i; if (rev) sd |= 1u << i; }
            }
            if (no_tie) {
                // positions outside [k, L) are masked here (the loop above is branch-free)
                uint32_t ok = 0xFFFFFFFFu;
                if (blk_pos < (int64_t)k) ok &= (blk_pos + 32 <= (int64_t)k) ? 0u : (0xFFFFFFFFu << (uint32_t)((int64_t)k - blk_pos));
                if (blk_pos + 32 > L) ok &= 0xFFFFFFFFu >> (uint32_t)(blk_pos + 32 - L);
                keep = kp & ok; strand = sd & ok;
                fast_done = true;
            }
        }
        if (!fast_done)
#pragma unroll 4
        for (int i = 0; i < 32; i++) {
            const uint32_t sh = 31 - i;
            const uint32_t f0lo = fsr(a0, a1, sh) & mlo, f0hi = fsr(a1, a2, sh) & mhi;
            const uint32_t f1lo = fsr(b0, b1, sh) & mlo, f1hi = fsr(b1, b2, sh) & mhi;
            const uint32_t r0lo = fsr(q00, q01, i) & mlo, r0hi = fsr(q01, q02, i) & mhi;
            const uint32_t r1lo = fsr(q10, q11, i) & mlo, r1hi = fsr(q11, q12, i) & mhi;
            const bool pal = f0lo == r0lo && f0hi == r0hi && f1lo == r1lo && f1hi == r1hi;
            const bool rev = (r0hi < f0hi) || (r0hi == f0hi && r0lo < f0lo);   // rmmer.0 < fmmer.0 (shmmrutils.rs:611)
            uint32_t ulo = rev ? r0lo : f0lo, uhi = rev ? r0hi : f0hi;
            uint32_t vlo = (rev ? r1lo : f1lo) ^ (uint32_t)HASH_XOR, vhi = rev ? r1hi : f1hi;
            u64hash_dev32(ulo, uhi);
            u64hash_dev32(vlo, vhi);
            const int64_t pos = blk_pos + i;
            const uint64_t h = ((uint64_t)(uhi ^ vhi) << 32) | (ulo ^ vlo);
            if (!pal && pos >= (int64_t)k && pos < L && h < thr) { keep |= 1u << i; if (rev) strand |= 1u << i; }
        }
    }
    p.keep[g - p.blk_base] = keep;
    p.strand[g - p.blk_base] = strand;
}

struct SketchSeqTable { const uint64_t *s_off; const uint32_t *s_len; const uint32_t *s_sid; uint32_t n; };   // sorted by offset

// one thread per bitmap word of the chunk: the marked blocks of the word through the exact machine
__global__ void sketch_dirty_kernel(const SketchTParams p, const SketchSeqTable t, uint64_t word_lo, uint64_t word_hi) {
    const uint64_t wi = word_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= word_hi) return;
    uint32_t word = p.dirty_bits[wi];
    const uint64_t thr = sketch_threshold(p.r);
    while (word) {
        const uint32_t b = __ffs(word) - 1;
        word &= word - 1;
        const uint64_t g = wi * 32 + b;
        uint32_t lo = 0, hi = t.n;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if ((t.s_off[mid] >> 5) <= g) lo = mid; else hi = mid; }
        const uint64_t soff = t.s_off[lo];
        const int64_t L = t.s_len[lo];
        const int64_t blk_pos = (int64_t)((g - (soff >> 5)) << 5);
        uint32_t keep = 0, strand = 0;
        sketch_walk_block(p.seq + soff, L, blk_pos, p.k, thr, p.allinv_bits, soff >> 5,
                          [&](int i, uint64_t, bool rev) { keep |= 1u << i; if (rev) strand |= 1u << i; });
        p.keep[g - p.blk_base] = keep;
        p.strand[g - p.blk_base] = strand;
    }
}

// kept positions per tile: one warp per tile
__global__ void sketch_count_kernel(const SketchTParams p) {
    const uint32_t tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (tile >= p.n_tiles) return;
    uint32_t sid = 0, j = 0;
    if (lane == 0) sketch_tile_lookup(p, tile, sid, j);
    sid = __shfl_sync(0xFFFFFFFFu, sid, 0); j = __shfl_sync(0xFFFFFFFFu, j, 0);
    const uint64_t L = p.len[sid];
    const uint64_t b0 = (uint64_t)j * L0_KB, b1 = min((L + 31) >> 5, b0 + L0_KB);   // blocks of the sequence in this tile
    const uint64_t g0 = (p.off[sid] >> 5) - p.blk_base;
    uint32_t c = 0;
    for (uint64_t b = b0 + lane; b < b1; b += 32) c += __popc(p.keep[g0 + b]);
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if (lane == 0) p.tile_count[tile] = c;
}

// One CTA per tile again.  (A variant without plane words that rebuilt the k-mer of every kept position from its k bytes was
// measured 3.4 x slower: 12 % of the lanes walk 56 bytes each while the rest of their warp waits.)
__global__ void __launch_bounds__(L0_NT) sketch_write_kernel(const SketchTParams p) {
    __shared__ SketchSmem s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t t_off = p.tile_off[blockIdx.x];
    if (p.tile_off[blockIdx.x + 1] == t_off) return;   // nothing kept in this tile (uniform for the CTA)
    if (tid == 0) sketch_tile_lookup(p, blockIdx.x, s.sid, s.j);
    __syncthreads();
    const uint32_t sid = s.sid;
    const int64_t L = p.len[sid];
    const uint64_t soff = p.off[sid];
    const uint8_t *sq = p.seq + soff;
    const int64_t blk_pos = (int64_t)s.j * SKT_KPOS + 32 * (tid - L0_CTX);
    uint32_t inv, minv;
    sketch_planes(s, sq, L, blk_pos, tid, inv, minv);
    const bool mine = tid >= L0_CTX && blk_pos < L;
    const uint64_t g = (soff + (uint64_t)max(blk_pos, (int64_t)0)) >> 5;
    const uint32_t keep = mine ? p.keep[g - p.blk_base] : 0u, strand = mine ? p.strand[g - p.blk_base] : 0u;
    // block-wide exclusive scan of the populations
    const uint32_t cnt = __popc(keep);
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += v; }
    if (lane == 31) s.wsum[warp] = incl;
    __syncthreads();   // planes and warp sums visible
    uint32_t wbase = 0;
    for (int i = 0; i < warp; i++) wbase += s.wsum[i];
    if (!cnt) return;
    pgr_mm128 *dst = p.out + t_off + wbase + incl - cnt;
    const uint32_t kk = p.k;
    const bool dirty = (p.dirty_bits[g >> 5] >> (g & 31)) & 1u;
    uint32_t n = 0;
    if (!dirty) {
        uint32_t rem = keep;
        while (rem) {
            const int i = __ffs(rem) - 1;
            rem &= rem - 1;
            uint32_t st;
            const uint64_t h = hash_at(s, 32 * (tid - L0_CTX) + i, kk, st);
            pgr_mm128 mm;
            mm.x = (h << 8) | kk;
            mm.y = ((uint64_t)sid << 32) | ((uint64_t)(uint32_t)(blk_pos + i) << 1) | ((strand >> i) & 1u);
            dst[n++] = mm;
        }
    } else {
        sketch_walk_block(sq, L, blk_pos, kk, sketch_threshold(p.r), p.allinv_bits, soff >> 5, [&](int i, uint64_t h, bool rev) {
            pgr_mm128 mm;
            mm.x = (h << 8) | kk;
            mm.y = ((uint64_t)sid << 32) | ((uint64_t)(uint32_t)(blk_pos + i) << 1) | (rev ? 1u : 0u);
            dst[n++] = mm;
        });
    }
}

}  // namespace pgr
