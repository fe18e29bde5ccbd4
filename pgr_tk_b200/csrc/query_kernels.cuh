// query_kernels.cuh — batched query_fragment_to_hps (aln.rs:147-242) and sparse_aln (aln.rs:12-142), and
// frag_map_to_adj_list (seq_db.rs:876-944).
#pragma once
#include "index_kernels.cuh"

namespace pgr {

// ---- generic exclusive scan of u32 counts into u64 offsets (n+1 entries) --------------------------------------------
constexpr int SC_NT = 256, SC_PER = 8, SC_BLK = SC_NT * SC_PER;
__global__ void __launch_bounds__(SC_NT) scan_reduce_kernel(const uint32_t *v, uint64_t n, uint32_t *block_sum) {
    __shared__ uint32_t wsum[SC_NT / 32];
    const uint64_t i0 = (uint64_t)blockIdx.x * SC_BLK;
    uint32_t c = 0;
    for (int j = 0; j < SC_PER; j++) { const uint64_t i = i0 + (uint64_t)j * SC_NT + threadIdx.x; if (i < n) c += v[i]; }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int i = 0; i < SC_NT / 32; i++) t += wsum[i]; block_sum[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(SC_NT) scan_apply_kernel(const uint32_t *v, uint64_t n, const uint64_t *block_prefix, uint64_t *out) {
    __shared__ uint32_t wsum[SC_NT / 32];
    const uint64_t i0 = (uint64_t)blockIdx.x * SC_BLK + (uint64_t)threadIdx.x * SC_PER;
    uint32_t x[SC_PER], c = 0;
#pragma unroll
    for (int j = 0; j < SC_PER; j++) { x[j] = (i0 + j < n) ? v[i0 + j] : 0; c += x[j]; }
    uint32_t incl = c;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((threadIdx.x & 31) >= d) incl += t; }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t wb = 0;
    for (uint32_t j = 0; j < (threadIdx.x >> 5); j++) wb += wsum[j];
    uint64_t r = block_prefix[blockIdx.x] + wb + incl - c;
#pragma unroll
    for (int j = 0; j < SC_PER; j++) { if (i0 + j < n) out[i0 + j] = r; r += x[j]; }
    if (i0 <= n && n < i0 + SC_PER) out[n] = block_prefix[blockIdx.x] + wb + incl;  // total, written by the thread that owns slot n
}

// ---- per-signature count of signatures with the same sid inside its key's vector (aln.rs:183-191) ----------------------
// run-length when the vector is sid-monotone (the normal case: sequences are added in sid order), full scan otherwise
__global__ void sid_count_kernel(const pgr_frag_sig *sigs, const uint64_t *offsets, uint64_t n_keys, uint32_t *sid_count) {
    const uint64_t kx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (kx >= n_keys) return;
    const uint64_t b = offsets[kx], e = offsets[kx + 1];
    bool mono = true;
    for (uint64_t i = b + 1; i < e; i++) mono = mono && (sigs[i - 1].sid <= sigs[i].sid);
    if (mono) {
        uint64_t i = b;
        while (i < e) {
            uint64_t j = i + 1;
            while (j < e && sigs[j].sid == sigs[i].sid) j++;
            for (uint64_t q = i; q < j; q++) sid_count[q] = (uint32_t)(j - i);
            i = j;
        }
    } else {
        for (uint64_t i = b; i < e; i++) {
            uint32_t c = 0;
            for (uint64_t j = b; j < e; j++) c += (sigs[j].sid == sigs[i].sid) ? 1u : 0u;
            sid_count[i] = c;
        }
    }
}

// ---- query pair statistics -------------------------------------------------------------------------------------------
// qcount[i] = number of pairs of the same query with the same key (shmmr_pair_hash_count, aln.rs:172-182)
__global__ void qpair_count_kernel(const FragTuple *qt, uint64_t n_qp, const uint64_t *qp_off, uint32_t *qcount) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_qp) return;
    const uint32_t q = qt[i].sid;
    const uint64_t b = qp_off[q], e = qp_off[q + 1];
    const uint64_t h0 = qt[i].h0, h1 = qt[i].h1;
    uint32_t c = 0;
    for (uint64_t j = b; j < e; j++) c += (qt[j].h0 == h0 && qt[j].h1 == h1) ? 1u : 0u;
    qcount[i] = c;
}

struct QueryFilter { uint32_t max_count, max_count_query, max_count_target; };

// number of hit pairs a query pair contributes after the count filters (aln.rs:197-228)
__global__ void hit_count_kernel(uint64_t n_qp, const uint32_t *qcount, const uint64_t *hit_begin, const uint32_t *hit_cnt,
                                 const uint32_t *sid_count, QueryFilter f, uint32_t *n_hits) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_qp) return;
    const uint32_t c = qcount[i];
    uint32_t n = 0;
    if (c <= f.max_count && c <= f.max_count_query) {
        const uint64_t b = hit_begin[i];
        for (uint32_t j = 0; j < hit_cnt[i]; j++) {
            const uint64_t tc = (uint64_t)c * sid_count[b + j];
            n += (tc <= f.max_count_target) ? 1u : 0u;
        }
    }
    n_hits[i] = n;
}

struct HitRec {          // 32 bytes
    uint32_t qid, sid;   // query ordinal, target sequence id
    uint32_t qb, qe, tb, te;
    uint8_t qo, to, pad_[2];
    uint32_t pad2_;
};
static_assert(sizeof(HitRec) == 32, "HitRec layout");

__global__ void hit_expand_kernel(const FragTuple *qt, uint64_t n_qp, const uint32_t *qcount, const uint64_t *hit_begin,
                                  const uint32_t *hit_cnt, const uint32_t *sid_count, const pgr_frag_sig *sigs, QueryFilter f,
                                  const uint64_t *hit_off, HitRec *hits, SortKey *keys) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_qp) return;
    const uint32_t c = qcount[i];
    if (!(c <= f.max_count && c <= f.max_count_query)) return;
    const FragTuple t = qt[i];
    uint64_t o = hit_off[i];
    const uint64_t b = hit_begin[i];
    for (uint32_t j = 0; j < hit_cnt[i]; j++) {
        const uint64_t tc = (uint64_t)c * sid_count[b + j];
        if (tc > f.max_count_target) continue;
        const pgr_frag_sig sg = sigs[b + j];
        HitRec h;
        h.qid = t.sid; h.sid = sg.sid; h.qb = t.bgn; h.qe = t.end; h.qo = (uint8_t)t.ori;
        h.tb = sg.bgn; h.te = sg.end; h.to = sg.ori; h.pad_[0] = h.pad_[1] = 0; h.pad2_ = 0;
        hits[o] = h;
        SortKey k; k.k0 = t.sid; k.k1 = sg.sid;
        keys[o] = k;
        o++;
    }
}

__global__ void hit_gather_kernel(const HitRec *in, const SortKey *keys, const uint32_t *idx, uint64_t n, HitRec *out, uint8_t *head) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = in[idx[i]];
    head[i] = (i == 0 || keys[i].k0 != keys[i - 1].k0 || keys[i].k1 != keys[i - 1].k1) ? 1 : 0;
}

// ---- generate_smp_adj_list_for_seq (seq_db.rs:946-1000), all sequences of a batch at once ---------------------------------
// pair i of a sequence and its successor w give two entries iff both keys are in the map with at least min_count signatures
// and v.end == w.bgn
__global__ void smp_adj_flag_kernel(const FragTuple *qt, uint64_t n_qp, const uint64_t *qp_off, const uint32_t *hit_cnt, const uint64_t *min_count,
                                    uint32_t *flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_qp) return;
    const uint32_t q = qt[i].sid;   // ordinal of the sequence in the batch
    uint32_t f = 0;
    if (i + 1 < qp_off[q + 1]) {
        const uint64_t mc = min_count[q];
        const uint32_t cv = hit_cnt[i], cw = hit_cnt[i + 1];
        f = (cv > 0 && cw > 0 && cv >= mc && cw >= mc && qt[i].end == qt[i + 1].bgn) ? 1u : 0u;
    }
    flag[i] = f;
}
__global__ void smp_adj_emit_kernel(const FragTuple *qt, uint64_t n_qp, const uint32_t *flag, const uint64_t *rank, const uint32_t *sids, pgr_adj_pair *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_qp || !flag[i]) return;
    const FragTuple v = qt[i], w = qt[i + 1];
    pgr_adj_pair a, b;
    a.sid = b.sid = sids[v.sid];
    a.pad_[0] = a.pad_[1] = b.pad_[0] = b.pad_[1] = 0;
    a.a0 = v.h0; a.a1 = v.h1; a.ori0 = (uint8_t)v.ori; a.b0 = w.h0; a.b1 = w.h1; a.ori1 = (uint8_t)w.ori;
    b.a0 = w.h0; b.a1 = w.h1; b.ori0 = (uint8_t)(1u - w.ori); b.b0 = v.h0; b.b1 = v.h1; b.ori1 = (uint8_t)(1u - v.ori);
    out[2 * rank[i]] = a;
    out[2 * rank[i] + 1] = b;
}

// ---- hits of one query -> stable order by target sid (aln.rs:213-228 builds one list per target in query-pair order) --------
// The hits of a batch are generated query by query; only the order by sid INSIDE a query is missing.  One CTA per query:
// (sid, position in the query's hit list) keys are sorted in shared memory (bitonic on 64-bit keys: position as the minor key
// makes it stable) and the 32-byte records are moved once.  A global LSD radix sort of the same data takes 8 passes.
constexpr int QS_NT = 256, QS_CAP = 8192;
__global__ void hit_query_offsets_kernel(const uint64_t *qp_off, const uint64_t *hit_off, uint64_t n_q, uint64_t *hq_off) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q <= n_q) hq_off[q] = hit_off[qp_off[q]];
}
__global__ void __launch_bounds__(QS_NT) query_sort_kernel(const HitRec *in, const uint64_t *hq_off, HitRec *out, uint8_t *head) {
    extern __shared__ __align__(16) unsigned char qs_raw[];
    uint64_t *key = reinterpret_cast<uint64_t *>(qs_raw);
    const uint64_t b = hq_off[blockIdx.x], e = hq_off[blockIdx.x + 1];
    const uint32_t n = (uint32_t)(e - b);
    if (n == 0) return;
    uint32_t m = 1;
    while (m < n) m <<= 1;
    for (uint32_t i = threadIdx.x; i < m; i += QS_NT) key[i] = i < n ? (((uint64_t)in[b + i].sid << 32) | i) : ~0ull;
    __syncthreads();
    for (uint32_t k = 2; k <= m; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < m; i += QS_NT) {
                const uint32_t l = i ^ j;
                if (l > i) {
                    const uint64_t x = key[i], y = key[l];
                    if ((x > y) == ((i & k) == 0)) { key[i] = y; key[l] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t r = threadIdx.x; r < n; r += QS_NT) {
        const uint64_t kx = key[r];
        out[b + r] = in[b + (uint32_t)kx];
        head[b + r] = (r == 0 || (uint32_t)(key[r - 1] >> 32) != (uint32_t)(kx >> 32)) ? 1 : 0;
    }
}
// segment keys (qid, sid) of the sorted hits for the run-length step
__global__ void hit_keys_kernel(const HitRec *h, uint64_t n, SortKey *keys) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SortKey k; k.k0 = h[i].qid; k.k1 = h[i].sid;
    keys[i] = k;
}

// ---- sparse_aln (aln.rs:12-142), one thread per (query, target) segment ---------------------------------------------
struct ChainParams {
    const HitRec *hits;          // sorted by (qid, sid), each segment already non-decreasing in qb (stable)
    const uint64_t *seg_off;     // [n_seg+1]
    uint64_t n_seg;
    uint32_t max_span;
    float penalty;
    uint32_t has_gap; float max_gap;
    uint32_t oriented;
    // per-hit scratch / outputs (same indexing as hits)
    float *v_s;                  // score of each vertex
    int32_t *best_pre;           // index (within segment) of the best predecessor, -1 = none
    uint32_t *cls_first, *cls_last;  // duplicate classes (HashMap keyed by HitPair value)
    uint32_t *order;             // heap-sorted vertex order for head selection
    uint8_t *visited;
    // outputs
    uint32_t *out_idx;           // chain members as segment-relative hit indices, chains back to back
    uint8_t *out_start;          // 1 where a chain starts
    float *out_score;            // chain score at its start slot
    uint32_t *seg_n_out, *seg_n_chains;
    uint32_t *seg_err;           // aln.rs would spin forever (all scores <= 0)
};

__device__ __forceinline__ bool same_hp(const HitRec &a, const HitRec &b) {
    return a.qb == b.qb && a.qe == b.qe && a.qo == b.qo && a.tb == b.tb && a.te == b.te && a.to == b.to;
}
__device__ __forceinline__ bool same_q(const HitRec &a, const HitRec &b) { return a.qb == b.qb && a.qe == b.qe && a.qo == b.qo; }

// order[] comparison: higher score first, then lower index (canonical replacement for FxHashSet iteration order)
__device__ __forceinline__ bool head_before(const float *vs, uint32_t a, uint32_t b) {
    return vs[a] > vs[b] || (vs[a] == vs[b] && a < b);
}

// sparse_aln on one segment: h[0..n) sorted by qb (stable); every array has n entries and may live in shared or global memory.
// PAR > 1: the duplicate classes are computed by PAR cooperating lanes (lane = 0..PAR-1, all lanes of the warp must call);
// everything sequential (DP, head order, traceback) runs on lane 0 only.
template <int PAR>
__device__ __forceinline__ void chain_segment(const ChainParams &p, const HitRec *h, uint32_t n, float *vs, int32_t *bp, uint32_t *cf, uint32_t *cl,
                                              uint32_t *ord, uint8_t *vis, uint32_t *oi, uint8_t *ost, float *osc, int lane, uint32_t &n_out_r,
                                              uint32_t &n_ch_r, uint32_t &err_r) {
    // duplicate classes: identical HitPairs share one map entry; equal values have equal qb, so they sit in one qb run
    for (uint32_t i = (uint32_t)lane; i < n; i += PAR) {
        uint32_t f = i, l = i;
        for (uint32_t j = i; j > 0 && h[j - 1].qb == h[i].qb; j--) if (same_hp(h[j - 1], h[i])) f = j - 1;
        for (uint32_t j = i + 1; j < n && h[j].qb == h[i].qb; j++) if (same_hp(h[j], h[i])) l = j;
        cf[i] = f; cl[i] = l;
    }
    if (PAR > 1) __syncwarp();
    n_out_r = 0; n_ch_r = 0; err_r = 0;
    if (lane != 0) return;
    // DP (aln.rs:25-103).  v_s.get(pre) sees the LATEST inserted duplicate with index < i
    vs[0] = __fsub_rn((float)h[0].qe, (float)h[0].qb);
    bp[0] = -1;
    for (uint32_t i = 1; i < n; i++) {
        const HitRec hp = h[i];
        const float len = __fsub_rn((float)hp.qe, (float)hp.qb);
        float best_s = 0.0f;
        int32_t best_v = -1;
        // span_set: distinct (qb,qe,qo) seen so far; hits are sorted by qb, so a small list suffices only when max_span
        // is small; in general count distinct values by scanning the already visited range
        uint32_t span = 0;
        for (uint32_t j = i; j-- > 0;) {
            const HitRec pre = h[j];
            if (p.oriented && ((pre.qo ^ pre.to) != (hp.qo ^ hp.to))) continue;
            if (p.has_gap) {
                const float dq = fabsf(__fsub_rn((float)hp.qb, (float)pre.qe));
                const float dt = (hp.qo == hp.to) ? fabsf(__fsub_rn((float)hp.tb, (float)pre.te)) : fabsf(__fsub_rn((float)hp.te, (float)pre.tb));
                if (dq > p.max_gap || dt > p.max_gap) continue;
            }
            if (same_q(pre, hp)) continue;
            // span_set.insert(pre.0): new iff no admitted predecessor with the same (qb,qe,qo) was seen at a larger index
            bool seen = false;
            for (uint32_t jj = j + 1; jj < i && !seen; jj++) {
                const HitRec o = h[jj];
                if (!same_q(o, pre)) { if (o.qb != pre.qb) break; continue; }
                // o has the same left fragment: it was inserted unless it was skipped by the gates above
                bool skipped = false;
                if (p.oriented && ((o.qo ^ o.to) != (hp.qo ^ hp.to))) skipped = true;
                if (!skipped && p.has_gap) {
                    const float dq = fabsf(__fsub_rn((float)hp.qb, (float)o.qe));
                    const float dt = (hp.qo == hp.to) ? fabsf(__fsub_rn((float)hp.tb, (float)o.te)) : fabsf(__fsub_rn((float)hp.te, (float)o.tb));
                    if (dq > p.max_gap || dt > p.max_gap) skipped = true;
                }
                if (!skipped) seen = true;
            }
            if (!seen) span++;
            // latest duplicate of pre with index < i
            uint32_t jl = j;
            for (uint32_t jj = j + 1; jj < i && h[jj].qb == pre.qb; jj++) if (same_hp(h[jj], pre)) jl = jj;
            const float p_s = vs[jl];
            float s = __fadd_rn(p_s, len);
            const float dq = fabsf(__fsub_rn((float)hp.qb, (float)pre.qe));
            const float dt = (hp.qo == hp.to) ? fabsf(__fsub_rn((float)hp.tb, (float)pre.te)) : fabsf(__fsub_rn((float)hp.te, (float)pre.tb));
            s = __fsub_rn(s, __fmul_rn(p.penalty, __fadd_rn(dq, dt)));
            if (s > best_s) { best_s = s; best_v = (int32_t)j; }
            if (span >= p.max_span) break;
        }
        if (best_s > 0.0f) { vs[i] = best_s; bp[i] = best_v; }
        else { vs[i] = len; bp[i] = -1; }
    }
    // final map values: a class reads the entry of its last duplicate
    // head order: heap sort of class representatives (first occurrences) by (score desc, index asc)
    uint32_t m = 0;
    for (uint32_t i = 0; i < n; i++) { vis[i] = 0; if (cf[i] == i) ord[m++] = i; }
    auto key_before = [&](uint32_t a, uint32_t bidx) {  // a, bidx are class representatives
        const float sa = vs[cl[a]], sb = vs[cl[bidx]];
        return sa > sb || (sa == sb && a < bidx);
    };
    // heap sort ascending in "before" order: build a max-heap on the inverse relation
    auto sift = [&](uint32_t start, uint32_t end) {
        uint32_t root = start;
        for (;;) {
            uint32_t child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && key_before(ord[child], ord[child + 1])) child++;   // pick the one that comes LATER
            if (key_before(ord[root], ord[child])) { const uint32_t t = ord[root]; ord[root] = ord[child]; ord[child] = t; root = child; }
            else break;
        }
    };
    if (m > 1) {
        for (uint32_t s = m / 2; s-- > 0;) sift(s, m);
        for (uint32_t end = m; end-- > 1;) { const uint32_t t = ord[0]; ord[0] = ord[end]; ord[end] = t; sift(0, end); }
    }
    // traceback (aln.rs:105-141)
    uint32_t n_out = 0, n_ch = 0;
        for (uint32_t t = 0; t < m; t++) {
        const uint32_t head = ord[t];
        if (vis[head]) continue;
        const float best_s = vs[cl[head]];
        if (!(best_s > 0.0f)) { err_r = 1; break; }   // the reference loops forever here
        const uint32_t start = n_out;
        int32_t v = (int32_t)head;
        while (v >= 0) {
            const uint32_t rep = cf[v];
            if (vis[rep]) break;
            vis[rep] = 1;
            oi[n_out++] = rep;
            v = bp[cl[rep]];
        }
        // reverse the track in place
        for (uint32_t a = start, z = n_out; a + 1 < z; a++, z--) { const uint32_t tt = oi[a]; oi[a] = oi[z - 1]; oi[z - 1] = tt; }
        for (uint32_t a = start; a < n_out; a++) ost[a] = 0;
        ost[start] = 1;
        osc[start] = __fsub_rn(best_s, vs[cl[oi[start]]]);
        n_ch++;
    }
    n_out_r = n_out;
    n_ch_r = n_ch;
}

// one thread per (query, target) segment, all arrays in global memory.  (Measured on config 4, 1.0 M segments / 28.6 M hits:
// 8.3 ms.  A warp per segment with the records staged in shared memory and the look-back of every hit spread over the lanes
// — ballots for span_set, warp arg-max for the best predecessor, rank sort for the heads, bit-exact — took 16.4 ms: a
// look-back visits ~8 predecessors, so three quarters of the lanes idle and the kernel becomes issue bound; with lane 0 alone
// doing the DP it took 60 ms.  Both were removed.  Round 2 also tried one thread per segment with the records of a CTA's 32
// segments staged in shared memory by one coalesced pass (48 KB per CTA): the DRAM traffic falls, but only 128 threads fit per
// SM and the DP is a chain of dependent instructions that lives on latency hiding: ~17 ms.  Removed as well.)
__global__ void chain_kernel(const ChainParams p) {
    const uint64_t sgi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sgi >= p.n_seg) return;
    const uint64_t b = p.seg_off[sgi], e = p.seg_off[sgi + 1];
    const uint32_t n = (uint32_t)(e - b);
    p.seg_n_out[sgi] = 0; p.seg_n_chains[sgi] = 0; p.seg_err[sgi] = 0;
    if (n < 2) return;  // aln.rs:237 filter(|(_sid, hps)| hps.len() > 1)
    uint32_t n_out, n_ch, err;
    chain_segment<1>(p, p.hits + b, n, p.v_s + b, p.best_pre + b, p.cls_first + b, p.cls_last + b, p.order + b, p.visited + b, p.out_idx + b,
                     p.out_start + b, p.out_score + b, 0, n_out, n_ch, err);
    p.seg_n_out[sgi] = n_out; p.seg_n_chains[sgi] = n_ch; p.seg_err[sgi] = err;
}

// nested result arrays built on the device: one thread per segment copies its chains to their final places
struct AssembleParams {
    const HitRec *hits; const uint64_t *seg_off; const SortKey *seg_keys; uint64_t n_seg;
    const uint32_t *out_idx; const uint8_t *out_start; const float *out_score;
    const uint32_t *seg_n_out;
    const uint64_t *hit_prefix, *chain_prefix, *target_prefix;     // exclusive scans over segments
    uint32_t *target_sid, *target_qid; uint64_t *target_chain_off;
    float *chain_score; uint64_t *chain_hit_off; pgr_hit_pair *hits_out;
    uint64_t chain_base, hit_base;   // chains / hits of the query groups before this one: the offsets written are global
};
__global__ void seg_has_kernel(const uint32_t *seg_n_out, uint64_t n_seg, uint32_t *has) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_seg) has[s] = seg_n_out[s] ? 1u : 0u;
}
__global__ void assemble_kernel(const AssembleParams p) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p.n_seg) return;
    const uint32_t n_out = p.seg_n_out[s];
    if (n_out == 0) return;
    const uint64_t t = p.target_prefix[s], b = p.seg_off[s], hb = p.hit_prefix[s];
    p.target_sid[t] = (uint32_t)p.seg_keys[s].k1;
    p.target_qid[t] = (uint32_t)p.seg_keys[s].k0;
    p.target_chain_off[t] = p.chain_base + p.chain_prefix[s];
    uint64_t c = p.chain_prefix[s];
    for (uint32_t a = 0; a < n_out; a++) {
        if (p.out_start[b + a]) { p.chain_score[c] = p.out_score[b + a]; p.chain_hit_off[c] = p.hit_base + hb + a; c++; }
        const HitRec h = p.hits[b + p.out_idx[b + a]];
        pgr_hit_pair hp;
        hp.qb = h.qb; hp.qe = h.qe; hp.tb = h.tb; hp.te = h.te; hp.qo = h.qo; hp.to = h.to; hp.pad_[0] = hp.pad_[1] = 0;
        p.hits_out[hb + a] = hp;
    }
}

// ---- frag_map_to_adj_list (seq_db.rs:876-944) --------------------------------------------------------------------------
struct AdjRow { uint32_t sid, bgn, end, ori; uint64_t h0, h1; uint32_t ok, pad_; };   // 40 bytes
static_assert(sizeof(AdjRow) == 40, "AdjRow layout");

// flatten the CSR into rows + a first sort key (sid, bgn); the remaining tuple fields are tie-breakers handled below
__global__ void adj_rows_kernel(const SortKey *ukeys, const uint64_t *offsets, const pgr_frag_sig *sigs, uint64_t n_sigs, uint64_t n_keys,
                                uint64_t min_count, const uint32_t *keeps, uint32_t n_keeps, uint32_t has_keeps, AdjRow *rows) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sigs) return;
    // key of signature i: largest kx with offsets[kx] <= i
    uint64_t lo = 0, hi = n_keys;
    while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (offsets[mid] <= i) lo = mid; else hi = mid; }
    const pgr_frag_sig sg = sigs[i];
    AdjRow r;
    r.sid = sg.sid; r.bgn = sg.bgn; r.end = sg.end; r.ori = sg.ori; r.h0 = ukeys[lo].k0; r.h1 = ukeys[lo].k1; r.pad_ = 0;
    bool ok = (offsets[lo + 1] - offsets[lo]) >= min_count;
    if (!ok && has_keeps) {
        uint32_t a = 0, z = n_keeps;  // keeps is sorted ascending by the host
        while (a < z) { const uint32_t mid = (a + z) >> 1; if (keeps[mid] < sg.sid) a = mid + 1; else z = mid; }
        ok = (a < n_keeps && keeps[a] == sg.sid);
    }
    r.ok = ok ? 1u : 0u;
    rows[i] = r;
}

// sort keys for the full-tuple order (sid, bgn, end, h0, h1, ori): three LSD rounds of the 128-bit sorter
//   round 1: (h1, ori)   round 2: (end, h0)   round 3: (sid, bgn)
__global__ void adj_keys_kernel(const AdjRow *rows, const uint32_t *idx, uint64_t n, int round, SortKey *keys) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const AdjRow &r = rows[idx ? idx[i] : i];
    SortKey k;
    if (round == 0) { k.k0 = r.h1; k.k1 = r.ori; }
    else if (round == 1) { k.k0 = r.end; k.k1 = r.h0; }
    else { k.k0 = r.sid; k.k1 = r.bgn; }
    keys[i] = k;
}

__global__ void adj_flag_kernel(const AdjRow *rows, const uint32_t *idx, uint64_t n, uint32_t *cnt) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = 0;
    if (i + 1 < n) {
        const AdjRow &v = rows[idx[i]], &w = rows[idx[i + 1]];
        if (v.ok && w.ok && v.sid == w.sid && v.end == w.bgn) c = 2;
    }
    cnt[i] = c;
}
__global__ void adj_emit_kernel(const AdjRow *rows, const uint32_t *idx, uint64_t n, const uint32_t *cnt, const uint64_t *off, pgr_adj_pair *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || cnt[i] == 0) return;
    const AdjRow &v = rows[idx[i]], &w = rows[idx[i + 1]];
    pgr_adj_pair a;
    a.sid = v.sid; a.pad_[0] = a.pad_[1] = 0;
    a.a0 = v.h0; a.a1 = v.h1; a.ori0 = (uint8_t)v.ori; a.b0 = w.h0; a.b1 = w.h1; a.ori1 = (uint8_t)w.ori;
    out[off[i]] = a;
    pgr_adj_pair r;
    r.sid = v.sid; r.pad_[0] = r.pad_[1] = 0;
    r.a0 = w.h0; r.a1 = w.h1; r.ori0 = (uint8_t)(1 - w.ori); r.b0 = v.h0; r.b1 = v.h1; r.ori1 = (uint8_t)(1 - v.ori);
    out[off[i] + 1] = r;
}

}  // namespace pgr
