// frags.cu — fragment compression of the FASTX path on the GPU: CompactSeqDB::seq_to_compressed's alignment branch
// (seq_db.rs:249-320) with shmmrutils::match_reads (shmmrutils.rs:57-223, the banded O(nD) variant), track_delta_point
// (:35-54) and deltas_to_aln_segs (seq_db.rs:113-156).
//
// The reference walks the sequences in order; an internal fragment (the bases between two adjacent shimmers, with the
// leading k-mer) longer than 128 bases is aligned against the earlier fragments of the same shimmer pair that are stored
// raw (`Fragment::Internal`), in insertion order, and becomes `AlnSegments` against the first that matches.  All fragments
// of one pair are one row of the finalized index (CSR, insertion order = (sequence, ordinal)), and rows are independent.
// The reference's decisions for a row — which entries stay Internal, which base each other entry aligns to — are
// reproduced with its alignment, band and traceback step by step, so the segments are identical (tests compare with
// oracle/frag_oracle.py, which is pinned to the reference's own .frg fixture).  Three kernels: (0a) every entry against
// the first entry of its row, one thread per entry — the common case, fully parallel and exact because the first entry is
// always raw and always the first candidate; (0b) the entries that failed there, sequentially per row; (1) the
// segments of the aligned entries written at their offsets, one thread per entry.
#include <algorithm>
#include <vector>

#include "index.cuh"

namespace pgr {

struct FragWork {
    const SortKey *ukeys; const uint64_t *offsets; const pgr_frag_sig *sigs; uint64_t n_keys;
    const uint8_t *seq; const uint64_t *seq_off;   // sequence store and per-sid byte offsets (sid -> offset; ~0 = absent)
    uint32_t n_sid, k;
    uint32_t d_cap;                                // scratch is sized for d_max <= d_cap
    uint32_t *scratch; uint64_t scratch_stride;    // per thread, in u32 words
    uint8_t *kind;                                 // [n_sigs] 0 = AlnSegments, 2 = Internal
    uint8_t *rc;                                   // [n_sigs]
    uint32_t *ref_sig;                             // [n_sigs] CSR index of the base fragment's signature
    uint32_t *n_segs;                              // [n_sigs]
    const uint64_t *seg_off;                       // [n_sigs+1] (pass 1)
    pgr_aln_seg *segs;                             // (pass 1)
    uint32_t *overflow;                            // set when a fragment needs d_max > d_cap
};

constexpr int FR_BAND = 32;                        // match_reads(.., bandwidth = 32) (seq_db.rs:285)
constexpr int FR_KPER = FR_BAND / 2 + 2;           // k values per d (band width <= 32, step 2) + slack

__device__ __forceinline__ uint8_t rc_base(uint8_t b) {   // fasta_io.rs:26-44
    switch (b) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
        default: return b;
    }
}

struct FragView {   // a fragment as match_reads sees it: seq[bgn-k .. end) of its sequence, reverse-complemented or not
    const uint8_t *p; uint32_t len; bool rc;
    __device__ __forceinline__ uint8_t at(uint32_t i) const { return rc ? rc_base(p[len - 1 - i]) : p[i]; }
};

// One alignment + traceback.  emit(type, a, b) receives the segments in the order deltas_to_aln_segs pushes them (i.e.
// REVERSED; the caller stores them back to front).  Returns false when match_reads returns None.
template <class Emit>
__device__ bool align_fragment(const FragWork &w, uint32_t *scr, const FragView &s0, const FragView &s1, Emit emit) {
    const uint32_t len0 = s0.len, len1 = s1.len;
    const uint32_t d_max = 32 + (uint32_t)(0.1 * (double)(len0 < len1 ? len0 : len1));
    if (d_max > w.d_cap) { *w.overflow = 1; return false; }
    // scratch: U[2*d_cap+3], V[2*d_cap+3] indexed by k + d_cap + 1; then per d: kmin, FR_KPER packed delta points
    uint32_t *U = scr, *V = scr + (2 * w.d_cap + 3);
    uint32_t *DP = V + (2 * w.d_cap + 3);
    const int kb = (int)w.d_cap + 1;
    for (int kk = -(int)d_max - 1; kk <= (int)d_max + 1; kk++) { U[kk + kb] = 0; V[kk + kb] = 0; }
    int k_min = 0, k_max = 0, best_m = -1;
    bool matched = false;
    uint32_t d_final = 0, end0 = 0, end1 = 0;
    int k_final = 0;
    for (uint32_t d = 0; d < d_max; d++) {
        if (k_max - k_min > FR_BAND) break;
        uint32_t *row = DP + (size_t)d * (FR_KPER + 1);
        row[0] = (uint32_t)k_min;
        for (int kk = k_min; kk <= k_max; kk += 2) {
            const uint32_t vn = V[kk - 1 + kb], vp = V[kk + 1 + kb];
            uint32_t x;
            int pre_k;
            if (kk == k_min || (kk != k_max && vn < vp)) { x = vp; pre_k = kk + 1; } else { x = vn + 1; pre_k = kk - 1; }
            uint32_t y = (uint32_t)((int)x - kk);
            row[1 + ((kk - k_min) >> 1)] = x | ((kk - pre_k) > 0 ? 0x80000000u : 0u);   // dk = +1 / -1
            while (x < len0 && y < len1 && s0.at(x) == s1.at(y)) { x++; y++; }
            U[kk + kb] = x + y; V[kk + kb] = x;
            if ((int)(x + y) > best_m) best_m = (int)(x + y);
            if (x >= len0 || y >= len1) { matched = true; d_final = d; k_final = kk; end0 = x; end1 = y; break; }
        }
        int k_max_new = k_min, k_min_new = k_max;
        for (int k2 = k_min; k2 <= k_max; k2 += 2)
            if ((int)U[k2 + kb] >= best_m - FR_BAND) { if (k2 < k_min_new) k_min_new = k2; if (k2 > k_max_new) k_max_new = k2; }
        k_max = k_max_new + 1; k_min = k_min_new - 1;
        if (matched) break;
    }
    if (!matched) return false;   // min_match_len = 0: a match is never discarded afterwards
    // deltas (track_delta_point: d_final .. 1, kept when bgn0 = 0 <= x <= end0) -> segments (deltas_to_aln_segs)
    // is the delta list empty?  (FullMatch needs that and equal lengths)
    bool any_delta = false;
    {
        uint32_t d = d_final; int kk = k_final;
        while (d > 0) {
            const uint32_t *row = DP + (size_t)d * (FR_KPER + 1);
            const uint32_t e = row[1 + ((kk - (int)row[0]) >> 1)];
            if ((e & 0x7FFFFFFFu) <= end0) { any_delta = true; break; }
            kk -= (e >> 31) ? 1 : -1;
            d--;
        }
    }
    if (!any_delta && len0 == len1) { emit(0u, 0u, 0u); return true; }
    uint32_t x = end0, y = end1;
    for (uint32_t yy = len1; yy > y; yy--) emit(2u, (uint32_t)s1.at(yy - 1), 0u);
    {
        uint32_t d = d_final; int kk = k_final;
        while (d > 0) {
            const uint32_t *row = DP + (size_t)d * (FR_KPER + 1);
            const uint32_t e = row[1 + ((kk - (int)row[0]) >> 1)];
            const uint32_t x1 = e & 0x7FFFFFFFu;
            const int dk = (e >> 31) ? 1 : -1;
            if (x1 <= end0) {
                const uint32_t y1 = (uint32_t)((int)x1 - kk);
                if (x1 < x) emit(1u, x1, x);
                x = x1; y = y1;
                if (dk > 0) x -= 1; else emit(2u, (uint32_t)s1.at(y - 1), 0u);
            }
            kk -= dk;
            d--;
        }
    }
    if (x != 0) emit(1u, 0u, x);
    return true;
}

__device__ __forceinline__ FragView frag_view(const FragWork &w, const pgr_frag_sig &sg, bool rc) {
    FragView f;
    f.len = sg.end - sg.bgn + w.k; f.rc = rc;
    f.p = w.seq + w.seq_off[sg.sid] + (sg.bgn - w.k);
    return f;
}

constexpr uint8_t FR_ALN = 0, FR_INTERNAL = 2, FR_PENDING = 3;

// Pass 0a — one thread per signature.  The first entry of a row is always raw, and it is the first candidate of every
// later entry of a LATER sequence: aligning against it needs nothing from the other entries, so all entries try it in
// parallel.  An entry that matches is final (the reference stops at the first match).  An entry of the same sequence as
// the first one has no candidate at all (candidates come from earlier sequences only) and stays raw.  Only an entry whose
// alignment against the first one fails depends on what became of the entries before it: it is left PENDING for pass 0b.
__global__ void frag_first_base_kernel(const FragWork w, uint32_t n_threads, uint64_t n_sigs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    uint32_t *scr = w.scratch + (size_t)t * w.scratch_stride;
    for (uint64_t j = t; j < n_sigs; j += n_threads) {
        uint64_t lo = 0, hi = w.n_keys;   // row of j: largest r with offsets[r] <= j
        while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (w.offsets[mid] <= j) lo = mid; else hi = mid; }
        const uint64_t r0 = w.offsets[lo];
        const pgr_frag_sig sj = w.sigs[j];
        uint8_t kind = FR_INTERNAL, rc = 0;
        uint32_t ref = 0, cnt = 0;
        if (j != r0 && sj.end - sj.bgn > 128) {
            const pgr_frag_sig s0 = w.sigs[r0];
            if (s0.sid < sj.sid) {
                const bool rv = sj.ori != s0.ori;
                if (align_fragment(w, scr, frag_view(w, s0, false), frag_view(w, sj, rv), [&](uint32_t, uint32_t, uint32_t) { cnt++; })) {
                    kind = FR_ALN; rc = rv ? 1 : 0; ref = (uint32_t)r0;
                } else {
                    kind = FR_PENDING; cnt = 0;
                }
            }
        }
        w.kind[j] = kind; w.rc[j] = rc; w.ref_sig[j] = ref; w.n_segs[j] = cnt;
    }
}

// Pass 0b — one thread per row: the pending entries, in insertion order, try the remaining raw entries before them that
// belong to earlier sequences (the first entry has been tried); an entry that matches nothing stays raw and is a
// candidate for the entries after it (seq_db.rs:258-318).
__global__ void frag_pending_kernel(const FragWork w, uint32_t n_threads) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    uint32_t *scr = w.scratch + (size_t)t * w.scratch_stride;
    for (uint64_t row = t; row < w.n_keys; row += n_threads) {
        const uint64_t r0 = w.offsets[row], r1 = w.offsets[row + 1];
        for (uint64_t j = r0 + 1; j < r1; j++) {
            if (w.kind[j] != FR_PENDING) continue;
            const pgr_frag_sig sj = w.sigs[j];
            uint8_t kind = FR_INTERNAL;
            for (uint64_t i = r0 + 1; i < j; i++) {
                const pgr_frag_sig si = w.sigs[i];
                if (si.sid >= sj.sid) break;              // the map a sequence sees holds earlier sequences only
                if (w.kind[i] != FR_INTERNAL) continue;   // only raw fragments serve as a base
                const bool rv = sj.ori != si.ori;
                uint32_t cnt = 0;
                if (align_fragment(w, scr, frag_view(w, si, false), frag_view(w, sj, rv), [&](uint32_t, uint32_t, uint32_t) { cnt++; })) {
                    kind = FR_ALN; w.rc[j] = rv ? 1 : 0; w.ref_sig[j] = (uint32_t)i; w.n_segs[j] = cnt;
                    break;
                }
            }
            w.kind[j] = kind;
        }
    }
}

// Pass 1 — one thread per aligned signature: the alignment once more, segments written back to front at their offsets
__global__ void frag_segments_kernel(const FragWork w, uint32_t n_threads, uint64_t n_sigs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    uint32_t *scr = w.scratch + (size_t)t * w.scratch_stride;
    for (uint64_t j = t; j < n_sigs; j += n_threads) {
        if (w.kind[j] != FR_ALN) continue;
        const pgr_frag_sig sj = w.sigs[j], si = w.sigs[w.ref_sig[j]];
        pgr_aln_seg *out = w.segs + w.seg_off[j];
        uint32_t pos = w.n_segs[j];
        align_fragment(w, scr, frag_view(w, si, false), frag_view(w, sj, w.rc[j] != 0), [&](uint32_t type, uint32_t a, uint32_t b) {
            pos--;
            pgr_aln_seg sg; sg.type = type; sg.a = a; sg.b = b;
            out[pos] = sg;
        });
    }
}

// per-signature results -> fragment records at their fragment id (the CSR order is by key, the file order by id: the
// permutation is a random scatter, done here rather than with cache-missing loops on the host).  Prefix / Suffix ids
// keep kind 0xFF and are filled in by the host, which walks the records sequentially.
__global__ void frag_record_kernel(const FragWork w, uint64_t n_sigs, pgr_fragment *rec, uint32_t n_frags, uint32_t *bad) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_sigs) return;
    const pgr_frag_sig sg = w.sigs[j];
    if (sg.frg_id >= n_frags) { *bad = 1; return; }
    pgr_fragment r;
    r.kind = w.kind[j]; r.reversed = w.rc[j]; r.pad_[0] = r.pad_[1] = 0;
    r.sid = sg.sid; r.bgn = sg.bgn - w.k; r.end = sg.end; r.len = sg.end - sg.bgn + w.k;
    r.ref_frag = 0; r.n_segs = 0; r.pad2_ = 0; r.seg_off = 0;
    if (r.kind == FR_ALN) { r.ref_frag = w.sigs[w.ref_sig[j]].frg_id; r.n_segs = w.n_segs[j]; r.seg_off = w.seg_off[j]; }
    rec[sg.frg_id] = r;
}

__global__ void frag_validate_kernel(const pgr_frag_sig *sigs, uint64_t n, const uint64_t *seq_len, uint32_t n_sid, uint32_t k, uint32_t *bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const pgr_frag_sig sg = sigs[i];
    if (sg.sid >= n_sid || seq_len[sg.sid] == 0 || sg.end > seq_len[sg.sid] || sg.bgn < k || sg.bgn >= sg.end) *bad = 1;
}

__global__ void max_span_kernel(const pgr_frag_sig *sigs, uint64_t n, uint32_t *mx) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicMax(mx, sigs[i].end - sigs[i].bgn);
}

}  // namespace pgr

using namespace pgr;

extern "C" {

int pgr_b200_index_compress_fragments(pgr_b200_index *idx, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens,
                                      pgr_fragment **frags, size_t *n_frags, pgr_aln_seg **segs, size_t *n_segs) {
    if (!idx || (n && (!sids || !seqs || !lens)) || !frags || !n_frags || !segs || !n_segs) { set_error("NULL argument"); return PGR_E_ARG; }
    if (idx->mode != 0) { set_error("fragment compression follows the FASTX fragment numbering (frg_id_mode 0)"); return PGR_E_ARG; }
    PGR_CUDA(cudaSetDevice(idx->ctx->device));
    PGR_TRY(pgr_b200_index_finalize(idx));
    cudaStream_t st = idx->ctx->stream;
    const uint64_t ns = idx->n_tuples, nk = idx->n_keys;
    const uint32_t k = idx->spec.k;
    // ---- sequence store on the device: sid -> offset ----
    uint32_t n_sid = 0;
    for (size_t i = 0; i < n; i++) n_sid = std::max<uint32_t>(n_sid, sids[i] + 1);
    std::vector<uint64_t> off(std::max<uint32_t>(n_sid, 1), ~0ull);
    std::vector<const uint8_t *> sp(n_sid, nullptr);
    std::vector<size_t> sl(n_sid, 0);
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) { off[sids[i]] = total; total += lens[i]; sp[sids[i]] = seqs[i]; sl[sids[i]] = lens[i]; }
    DevBuf d_seq, d_off, d_kind, d_rc, d_ref, d_ns, d_segoff, d_segs, d_scr, d_flag;
    auto release_all = [&]() { for (DevBuf *b : {&d_seq, &d_off, &d_kind, &d_rc, &d_ref, &d_ns, &d_segoff, &d_segs, &d_scr, &d_flag}) b->release(); };
    struct Guard { decltype(release_all) &f; ~Guard() { f(); } } guard{release_all};   // every exit path frees the device buffers (release is idempotent)
    PGR_TRY(d_seq.ensure(std::max<uint64_t>(total, 1)));
    PGR_TRY(d_off.ensure(off.size() * sizeof(uint64_t)));
    {
        // pageable sources go through the page-locked ring with the host pool doing the copies (PCIe rate instead of ~10 GB/s)
        pgr_b200_ctx *ctx = idx->ctx;
        const bool staged = total >= PACK_MIN_BYTES && !source_page_locked(seqs, lens, n) && (ctx->pack || (ctx->pack = pack_ring_acquire(ctx->device)));
        for (size_t i = 0; i < n; i++) {
            if (!lens[i]) continue;
            if (staged && lens[i] >= (1u << 20)) PGR_TRY(upload_raw_staged(ctx->pack, (uint8_t *)d_seq.p + off[sids[i]], seqs[i], lens[i], st));
            else PGR_CUDA(cudaMemcpyAsync((uint8_t *)d_seq.p + off[sids[i]], seqs[i], lens[i], cudaMemcpyHostToDevice, st));
        }
    }
    PGR_CUDA(cudaMemcpyAsync(d_off.p, off.data(), off.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    PGR_CUDA(cudaStreamSynchronize(st));
    {   // every signature must refer to a sequence that was passed, inside its bounds
        std::vector<uint64_t> lens64(std::max<uint32_t>(n_sid, 1), 0);
        for (uint32_t sI = 0; sI < n_sid; sI++) lens64[sI] = off[sI] == ~0ull ? 0 : sl[sI];
        DevBuf d_len;
        PGR_TRY(d_len.ensure(lens64.size() * sizeof(uint64_t)));
        PGR_TRY(d_flag.ensure(64));
        PGR_CUDA(cudaMemcpyAsync(d_len.p, lens64.data(), lens64.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        PGR_CUDA(cudaMemsetAsync(d_flag.p, 0, 64, st));
        if (ns) frag_validate_kernel<<<(unsigned)ceil_div<uint64_t>(ns, 256), 256, 0, st>>>(idx->sigs.as<pgr_frag_sig>(), ns, d_len.as<uint64_t>(), n_sid, k, d_flag.as<uint32_t>() + 2);
        uint32_t h_bad[3] = {0, 0, 0};
        PGR_CUDA(cudaMemcpyAsync(h_bad, d_flag.p, sizeof h_bad, cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaStreamSynchronize(st));
        d_len.release();
        if (h_bad[2]) { set_error("a fragment refers to a sequence that was not passed, or lies outside it"); return PGR_E_ARG; }
    }
    trace_mark("compress_fragments: sigs D2H + sequences H2D");
    // ---- device work ----
    const size_t nf = idx->n_frags;
    DevBuf d_rec;
    auto drop_rec = [&]() { d_rec.release(); };
    struct Guard2 { decltype(drop_rec) &f; ~Guard2() { f(); } } guard2{drop_rec};
    PGR_TRY(d_rec.ensure(std::max<size_t>(nf, 1) * sizeof(pgr_fragment)));
    PGR_CUDA(cudaMemsetAsync(d_rec.p, 0xFF, std::max<size_t>(nf, 1) * sizeof(pgr_fragment), st));   // kind 0xFF = not an internal fragment
    uint64_t tot_segs = 0;
    *frags = nullptr; *segs = nullptr;
    if (ns) {
        PGR_TRY(d_kind.ensure(ns)); PGR_TRY(d_rc.ensure(ns)); PGR_TRY(d_ref.ensure(ns * sizeof(uint32_t))); PGR_TRY(d_ns.ensure(ns * sizeof(uint32_t)));
        PGR_TRY(d_segoff.ensure((ns + 1) * sizeof(uint64_t)));
        PGR_CUDA(cudaMemsetAsync(d_flag.p, 0, 64, st));
        max_span_kernel<<<(unsigned)ceil_div<uint64_t>(ns, 256), 256, 0, st>>>(idx->sigs.as<pgr_frag_sig>(), ns, d_flag.as<uint32_t>() + 1);
        uint32_t h_flag[2];
        PGR_CUDA(cudaMemcpyAsync(h_flag, d_flag.p, sizeof h_flag, cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaStreamSynchronize(st));
        FragWork w;
        w.ukeys = idx->ukeys.as<SortKey>(); w.offsets = idx->offsets.as<uint64_t>(); w.sigs = idx->sigs.as<pgr_frag_sig>(); w.n_keys = nk;
        w.seq = d_seq.as<uint8_t>(); w.seq_off = d_off.as<uint64_t>(); w.n_sid = n_sid; w.k = k;
        w.d_cap = 32 + (uint32_t)(0.1 * (double)(h_flag[1] + k)) + 1;
        w.scratch_stride = 2ull * (2 * w.d_cap + 3) + (uint64_t)w.d_cap * (FR_KPER + 1);
        uint32_t n_threads = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(ns, 1), 65536);
        while (n_threads > 256 && (uint64_t)n_threads * w.scratch_stride * 4 > (2ull << 30)) n_threads /= 2;   // <= 2 GiB of scratch
        PGR_TRY(d_scr.ensure((uint64_t)n_threads * w.scratch_stride * sizeof(uint32_t)));
        w.scratch = d_scr.as<uint32_t>();
        w.kind = d_kind.as<uint8_t>(); w.rc = d_rc.as<uint8_t>(); w.ref_sig = d_ref.as<uint32_t>(); w.n_segs = d_ns.as<uint32_t>();
        w.seg_off = nullptr; w.segs = nullptr; w.overflow = d_flag.as<uint32_t>();
        const unsigned grid = ceil_div<uint32_t>(n_threads, 128);
        frag_first_base_kernel<<<grid, 128, 0, st>>>(w, n_threads, ns);
        frag_pending_kernel<<<grid, 128, 0, st>>>(w, n_threads);
        idx->launches += 3;
        PGR_CUDA(cudaGetLastError());
        trace_mark("compress_fragments: pass 0 (decide + count)");
        PGR_TRY(scan_u32(idx, d_ns.as<uint32_t>(), ns, d_segoff.as<uint64_t>(), &tot_segs));
        PGR_TRY(d_segs.ensure(std::max<uint64_t>(tot_segs, 1) * sizeof(pgr_aln_seg)));
        w.seg_off = d_segoff.as<uint64_t>(); w.segs = d_segs.as<pgr_aln_seg>();
        frag_segments_kernel<<<grid, 128, 0, st>>>(w, n_threads, ns);
        frag_record_kernel<<<(unsigned)ceil_div<uint64_t>(ns, 256), 256, 0, st>>>(w, ns, d_rec.as<pgr_fragment>(), (uint32_t)nf, d_flag.as<uint32_t>() + 2);
        idx->launches += 2;
        PGR_CUDA(cudaGetLastError());
        trace_mark("compress_fragments: pass 1 (segments) + records");
    }
    // ---- results to the host: records in fragment-id order, segments ----
    // large record sets come back through the page-locked ring into plain memory (download_staged); small ones into pool buffers
    const size_t rec_bytes = std::max<size_t>(nf, 1) * sizeof(pgr_fragment), seg_bytes = std::max<size_t>(tot_segs, 1) * sizeof(pgr_aln_seg);
    pgr_b200_ctx *rctx = idx->ctx;
    const bool staged_out = rec_bytes + seg_bytes >= (64u << 20) && (rctx->pack || (rctx->pack = pack_ring_acquire(rctx->device)));
    *frags = (pgr_fragment *)(staged_out ? malloc(rec_bytes) : result_alloc(rec_bytes));
    *segs = (pgr_aln_seg *)(staged_out ? malloc(seg_bytes) : result_alloc(seg_bytes));
    auto drop = [&](int rc_) { result_free(*frags); result_free(*segs); *frags = nullptr; *segs = nullptr; return rc_; };   // error exits release the outputs
    if (!*frags || !*segs) { set_error("out of host memory"); return drop(PGR_E_ARG); }
    {
        uint32_t h_flag[3] = {0, 0, 0};
        cudaError_t ce = cudaSuccess;
        if (staged_out) {
            int rc2 = download_staged(rctx->pack, (uint8_t *)*frags, (const uint8_t *)d_rec.p, rec_bytes, st);
            if (rc2 == PGR_OK && tot_segs) rc2 = download_staged(rctx->pack, (uint8_t *)*segs, (const uint8_t *)d_segs.p, tot_segs * sizeof(pgr_aln_seg), st);
            if (rc2 != PGR_OK) return drop(rc2);
        } else {
            ce = cudaMemcpyAsync(*frags, d_rec.p, rec_bytes, cudaMemcpyDeviceToHost, st);
            if (ce == cudaSuccess && tot_segs) ce = cudaMemcpyAsync(*segs, d_segs.p, tot_segs * sizeof(pgr_aln_seg), cudaMemcpyDeviceToHost, st);
        }
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_flag, d_flag.p, sizeof h_flag, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) { set_error("copying the fragment records failed: %s", cudaGetErrorString(ce)); return drop(PGR_E_CUDA); }
        if (h_flag[0]) { set_error("fragment longer than the alignment scratch was sized for"); return drop(PGR_E_LIMIT); }
        if (h_flag[2]) { set_error("fragment id out of range"); return drop(PGR_E_ARG); }
    }
    trace_mark("compress_fragments: results D2H");
    // ---- Prefix / Suffix records (seq_db.rs:203-231, :326-347): a sequential walk over the records in id order ----
    pgr_fragment *R = *frags;
    size_t f = 0;
    auto put = [&](uint8_t kind, uint32_t sid, uint32_t bgn, uint32_t end) {
        pgr_fragment &r = R[f++];
        memset(&r, 0, sizeof r);
        r.kind = kind; r.sid = sid; r.bgn = bgn; r.end = end; r.len = end - bgn;
    };
    for (uint32_t s = 0; s < n_sid; s++) {
        if (off[s] == ~0ull) continue;
        if (f >= nf) { set_error("more fragments than the index counted"); return drop(PGR_E_ARG); }
        const uint32_t L = (uint32_t)sl[s];
        if (R[f].kind != 0xFF) { set_error("fragment numbering does not match the sequences passed (sid %u)", s); return drop(PGR_E_ARG); }
        if (f + 1 < nf && R[f + 1].kind != 0xFF && R[f + 1].sid == s) {
            put(1, s, 0, R[f + 1].bgn + k);                                              // Prefix(seq[..pos0 + 1])
            uint32_t last_end = 0;
            while (f < nf && R[f].kind != 0xFF && R[f].sid == s) { last_end = R[f].end; f++; }   // the internal fragments are in place
            if (f >= nf || R[f].kind != 0xFF) { set_error("fragment numbering does not match the sequences passed (sid %u)", s); return drop(PGR_E_ARG); }
            put(3, s, last_end, L);                                                      // Suffix(seq[pos_last + 1..])
        } else {
            // no pair: 0 shimmers -> Prefix(whole), Suffix(empty); 1 shimmer -> split after it.  Rare: recompute the shimmers.
            pgr_mm128 *mm = nullptr;
            size_t nm = 0;
            { const int rc_ = pgr_b200_sequence_to_shmmrs(s, sp[s], sl[s], &idx->spec, 0, &mm, &nm); if (rc_ != PGR_OK) return drop(rc_); }
            const uint32_t cut = nm ? (((uint32_t)(mm[0].y & 0xFFFFFFFFu) >> 1) + 1) : L;
            pgr_b200_free(mm);
            if (f + 1 >= nf || R[f + 1].kind != 0xFF) { set_error("fragment numbering does not match the sequences passed (sid %u)", s); return drop(PGR_E_ARG); }
            put(1, s, 0, cut);
            put(3, s, cut, L);
        }
    }
    if (f != nf) { set_error("fragment count mismatch: assembled %zu, index counted %zu", f, nf); return drop(PGR_E_ARG); }
    *n_frags = nf;
    *n_segs = tot_segs;
    trace_mark("compress_fragments: prefix / suffix records");
    return PGR_OK;
}

}  // extern "C"
