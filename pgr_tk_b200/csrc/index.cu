// index.cu — ShmmrFragMap (CompactSeqDB.frag_map, seq_db.rs:72-100) as a device-resident CSR:
//   tuples (insertion order) --stable radix sort by (h0,h1)--> sigs[] + unique keys[] + offsets[]
// plus .mdb I/O (seq_db.rs:1291-1407).  Query, chaining and adjacency live in query.cu.
#include "index.cuh"

#include <memory>

#include "hostpack.hpp"

#include <algorithm>
#include <cstring>

using namespace pgr;

namespace pgr {

int index_reserve_tuples(pgr_b200_index *idx, uint64_t need) {
    if (need * sizeof(FragTuple) <= idx->tuples.cap) return PGR_OK;
    DevBuf nb;
    PGR_TRY(nb.ensure(std::max<uint64_t>(need, idx->n_tuples * 2) * sizeof(FragTuple)));
    if (idx->n_tuples)
        PGR_CUDA(cudaMemcpyAsync(nb.p, idx->tuples.p, idx->n_tuples * sizeof(FragTuple), cudaMemcpyDeviceToDevice, idx->ctx->stream));
    PGR_CUDA(cudaStreamSynchronize(idx->ctx->stream));
    idx->tuples.release();
    idx->tuples = nb;
    return PGR_OK;
}

// tuples of the sequences [c0, c0+cn) whose final shimmers are the ctx's current device result
static int emit_tuples(pgr_b200_index *idx, const uint32_t *sids, size_t cn, size_t n_mm, bool query_mode, FragTuple *dst_override,
                       uint64_t *n_pairs_out) {
    pgr_b200_ctx *ctx = idx->ctx;
    cudaStream_t st = ctx->stream;
    std::vector<uint64_t> off(cn + 1);
    PGR_CUDA(cudaMemcpyAsync(off.data(), ctx->d_result_off, (cn + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    PGR_CUDA(cudaStreamSynchronize(st));
    std::vector<uint64_t> pair_off(cn + 1);
    std::vector<uint32_t> frg_base(cn);
    uint64_t np = 0;
    uint32_t frags = idx->n_frags;
    for (size_t i = 0; i < cn; i++) {
        const uint64_t ns = off[i + 1] - off[i];
        pair_off[i] = np;
        if (idx->mode == 0 && !query_mode) {
            // seq_db.rs:203-231 (empty: two fragments), :224-231 prefix, :326-340 one per pair, :342-347 suffix
            frg_base[i] = frags + 1;
            frags += (ns == 0) ? 2u : (uint32_t)(ns + 1);
        } else {
            frg_base[i] = 0;  // seq_db.rs:402-407: per-sequence pair ordinal
        }
        np += ns ? ns - 1 : 0;
    }
    pair_off[cn] = np;
    if (!query_mode) idx->n_frags = frags;
    *n_pairs_out = np;
    const uint32_t ord0 = query_mode ? 0u : idx->ord_base + idx->ord_in_batch;   // insertion ordinal of the chunk's first sequence
    if (!query_mode) idx->ord_in_batch += (uint32_t)cn;
    FragTuple *dst = dst_override;
    if (!dst) {
        PGR_TRY(index_reserve_tuples(idx, idx->n_tuples + np));
        dst = idx->tuples.as<FragTuple>() + idx->n_tuples;
    }
    if (np == 0 || n_mm == 0) return PGR_OK;
    PGR_TRY(idx->d_sid.ensure(cn * sizeof(uint32_t)));
    PGR_TRY(idx->d_pair_off.ensure((cn + 1) * sizeof(uint64_t)));
    PGR_TRY(idx->d_frg_base.ensure(cn * sizeof(uint32_t)));
    PGR_CUDA(cudaMemcpyAsync(idx->d_sid.p, sids, cn * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PGR_CUDA(cudaMemcpyAsync(idx->d_pair_off.p, pair_off.data(), (cn + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    PGR_CUDA(cudaMemcpyAsync(idx->d_frg_base.p, frg_base.data(), cn * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PairParams p;
    p.mm = ctx->d_result; p.mm_off = ctx->d_result_off; p.n_seq = (uint32_t)cn; p.sid = idx->d_sid.as<uint32_t>();
    p.pair_off = idx->d_pair_off.as<uint64_t>(); p.frg_base = idx->d_frg_base.as<uint32_t>(); p.out = dst; p.n_mm = n_mm;
    p.query_mode = query_mode ? 1u : 0u;
    p.ord_base = ord0;
    pair_tuples_kernel<<<(uint32_t)ceil_div<uint64_t>(n_mm, 256), 256, 0, st>>>(p);
    PGR_CUDA(cudaGetLastError());
    PGR_CUDA(cudaStreamSynchronize(st));  // host vectors above go out of scope
    if (!dst_override) idx->n_tuples += np;
    return PGR_OK;
}

// shimmers of a host batch -> tuples appended to `dst` semantics of emit_tuples; used by add_batch and by queries
int index_batch_tuples(pgr_b200_index *idx, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens,
                       bool query_mode, DevBuf *qbuf, uint64_t *n_pairs_total, std::vector<uint64_t> *pairs_per_seq) {
    pgr_b200_ctx *ctx = idx->ctx;
    uint64_t total = 0;
    if (pairs_per_seq) pairs_per_seq->assign(n, 0);
    std::vector<uint32_t> ord(n);
    for (size_t i = 0; i < n; i++) ord[i] = (uint32_t)i;
    if (query_mode) {
        // queries: tuples go to a scratch buffer; its size is not known before the shimmers are: grow as chunks arrive
    }
    const int rc = run_chunked(ctx, n, ord.data(), seqs, lens, idx->spec, 0, [&](size_t c0, size_t cn, size_t ns) -> int {
        uint64_t np = 0;
        FragTuple *dst = nullptr;
        std::vector<uint64_t> off;
        if (query_mode) {
            // upper bound for this chunk: one pair per shimmer
            const uint64_t need = (total + ns + 1) * sizeof(FragTuple);
            if (need > qbuf->cap) {
                DevBuf nb;
                PGR_TRY(nb.ensure(std::max<uint64_t>(need, qbuf->cap * 2)));
                if (total) PGR_CUDA(cudaMemcpyAsync(nb.p, qbuf->p, total * sizeof(FragTuple), cudaMemcpyDeviceToDevice, ctx->stream));
                PGR_CUDA(cudaStreamSynchronize(ctx->stream));
                qbuf->release();
                *qbuf = nb;
            }
            dst = qbuf->as<FragTuple>() + total;
        }
        if (pairs_per_seq) {
            off.resize(cn + 1);
            PGR_CUDA(cudaMemcpyAsync(off.data(), ctx->d_result_off, (cn + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
            PGR_CUDA(cudaStreamSynchronize(ctx->stream));
            for (size_t i = 0; i < cn; i++) { const uint64_t k = off[i + 1] - off[i]; (*pairs_per_seq)[c0 + i] = k ? k - 1 : 0; }
        }
        PGR_TRY(emit_tuples(idx, sids + c0, cn, ns, query_mode, dst, &np));
        total += np;
        return PGR_OK;
    });
    if (n_pairs_total) *n_pairs_total = total;
    return rc;
}

// stable LSD radix sort of (keys, idx) by the 112-bit key; result in keysA/idxA.  One-sweep form (sort_kernels.cuh): all digit
// histograms from one read, constant digits skipped without a launch, one kernel per remaining pass.
// PGR_B200_SORT_3KERNEL=1 selects the first-generation three-kernel passes (A/B aid; also taken for n >= 2^30).
static int index_sort_3kernel(pgr_b200_index *idx, uint64_t n, int first_pass, int last_pass);
int index_sort(pgr_b200_index *idx, uint64_t n, int first_pass, int last_pass) {
    static const bool old_sort = getenv("PGR_B200_SORT_3KERNEL") != nullptr;
    if (old_sort || n >= (1ull << 30) || n == 0) return index_sort_3kernel(idx, n, first_pass, last_pass);
    cudaStream_t st = idx->ctx->stream;
    const uint32_t n_tiles = (uint32_t)ceil_div<uint64_t>(n, OS_TILE);
    PGR_TRY(idx->hist.ensure((size_t)OS_PASSES * 256 * 4 + OS_PASSES * 4 + 64 + (size_t)n_tiles * 256 * 4));
    PGR_TRY(idx->keysB.ensure(n * sizeof(SortKey)));
    PGR_TRY(idx->idxB.ensure(n * sizeof(uint32_t)));
    uint32_t *hist = idx->hist.as<uint32_t>(), *skip = hist + OS_PASSES * 256, *ticket = skip + OS_PASSES, *status = ticket + 2;
    PGR_CUDA(cudaMemsetAsync(hist, 0, (size_t)OS_PASSES * 256 * 4 + OS_PASSES * 4 + 8, st));
    const uint32_t hg = (uint32_t)std::min<uint64_t>(ceil_div<uint64_t>(n, 256 * 16), (uint64_t)idx->ctx->n_sm * 8);
    os_hist_kernel<<<std::max(1u, hg), 256, 0, st>>>(idx->keysA.as<SortKey>(), n, first_pass, last_pass, hist);
    os_scan_kernel<<<OS_PASSES, 256, 0, st>>>(hist, n, skip);
    idx->launches += 2;
    PGR_CUDA(cudaGetLastError());
    PGR_TRY(idx->ctx->ensure_ctl(64));
    uint32_t *h_skip = (uint32_t *)idx->ctx->h_ctl;
    PGR_CUDA(cudaMemcpyAsync(h_skip, skip, OS_PASSES * 4, cudaMemcpyDeviceToHost, st));
    PGR_CUDA(cudaStreamSynchronize(st));
    static bool attr_set = false;
    if (!attr_set) { PGR_CUDA(cudaFuncSetAttribute(os_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OsSmem))); attr_set = true; }
    SortKey *ka = idx->keysA.as<SortKey>(), *kb = idx->keysB.as<SortKey>();
    uint32_t *ia = idx->idxA.as<uint32_t>(), *ib = idx->idxB.as<uint32_t>();
    for (int pass = first_pass; pass <= last_pass; pass++) {
        if (h_skip[pass]) continue;
        PGR_CUDA(cudaMemsetAsync(ticket, 0, 8 + (size_t)n_tiles * 256 * 4, st));
        os_pass_kernel<<<n_tiles, OS_NT, sizeof(OsSmem), st>>>(ka, ia, n, pass, hist + pass * 256, status, ticket, kb, ib);
        idx->launches += 1;
        std::swap(ka, kb); std::swap(ia, ib);
        trace_mark("index_sort: pass");
    }
    PGR_CUDA(cudaGetLastError());
    if (ka != idx->keysA.as<SortKey>()) { std::swap(idx->keysA, idx->keysB); std::swap(idx->idxA, idx->idxB); }
    return PGR_OK;
}

static int index_sort_3kernel(pgr_b200_index *idx, uint64_t n, int first_pass, int last_pass) {
    cudaStream_t st = idx->ctx->stream;
    const uint32_t n_seg = (uint32_t)ceil_div<uint64_t>(n, RS_SEG);
    const uint32_t grid = ceil_div<uint32_t>(n_seg, RS_WARPS);
    PGR_TRY(idx->hist.ensure((size_t)n_seg * 256 * sizeof(uint32_t) + 64));
    PGR_TRY(idx->keysB.ensure(n * sizeof(SortKey)));
    PGR_TRY(idx->idxB.ensure(n * sizeof(uint32_t)));
    uint32_t *skip = idx->hist.as<uint32_t>() + (size_t)n_seg * 256;
    SortKey *ka = idx->keysA.as<SortKey>(), *kb = idx->keysB.as<SortKey>();
    uint32_t *ia = idx->idxA.as<uint32_t>(), *ib = idx->idxB.as<uint32_t>();
    PGR_TRY(idx->ctx->ensure_ctl(64));
    for (int pass = first_pass; pass <= last_pass; pass++) {
        rs_hist_kernel<<<grid, RS_NT, 0, st>>>(ka, n, pass, idx->hist.as<uint32_t>(), n_seg);
        rs_scan_kernel<<<1, 1024, 0, st>>>(idx->hist.as<uint32_t>(), (uint64_t)n_seg * 256, n, n_seg, skip);
        rs_scatter_kernel<<<grid, RS_NT, 0, st>>>(ka, ia, n, pass, idx->hist.as<uint32_t>(), n_seg, kb, ib, skip);
        PGR_CUDA(cudaGetLastError());
        PGR_CUDA(cudaMemcpyAsync(idx->ctx->h_ctl, skip, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaStreamSynchronize(st));
        idx->launches += 3;
        if (*(uint32_t *)idx->ctx->h_ctl == 0) { std::swap(ka, kb); std::swap(ia, ib); }
        trace_mark("index_sort: pass");
    }
    if (ka != idx->keysA.as<SortKey>()) { std::swap(idx->keysA, idx->keysB); std::swap(idx->idxA, idx->idxB); }
    return PGR_OK;
}

__global__ void iota_kernel(uint32_t *p, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}
__global__ void set_u64_kernel(uint64_t *p, uint64_t v) { *p = v; }
__global__ void add_frg_base_kernel(FragTuple *t, uint64_t n, uint32_t base) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) t[i].frg_id += base;
}

// destination part of every tuple: number of splitters <= h0
__global__ void dest_keys_kernel(const FragTuple *t, uint64_t n, const uint64_t *splitters, uint32_t n_split, SortKey *keys,
                                 unsigned long long *counts) {
    __shared__ unsigned int local[256];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) local[i] = 0;
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint64_t h0 = t[i].h0;
        uint32_t lo = 0, hi = n_split;  // first splitter > h0
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (splitters[mid] <= h0) lo = mid + 1; else hi = mid; }
        SortKey k; k.k0 = lo; k.k1 = 0;
        keys[i] = k;
        atomicAdd(&local[lo], 1u);
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d <= n_split; d += blockDim.x) if (local[d]) atomicAdd(&counts[d], (unsigned long long)local[d]);
}
__global__ void gather_tuples_kernel(const FragTuple *in, const uint32_t *idx, uint64_t n, FragTuple *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[idx[i]];
}

// the .mdb byte stream of a canonical CSR (seq_db.rs:1291-1326): "mdb", spec as 5 x u32, n_keys, then per key
// (h0, h1, len, len x 17-byte FragmentSignature)
int write_mdb_file(const pgr_shmmr_spec &spec, uint64_t nk, const uint64_t *keys, const uint64_t *offs, const pgr_frag_sig *sigs, const char *path) {
    const uint64_t ns = nk ? offs[nk] : 0;
    const size_t total = 31 + nk * 24 + ns * 17;
    std::unique_ptr<uint8_t[]> buf(new uint8_t[total]);   // no zero fill: every byte is written below
    {
        uint8_t *w = buf.get();
        auto p32 = [&](uint32_t v) { memcpy(w, &v, 4); w += 4; };
        *w++ = 'm'; *w++ = 'd'; *w++ = 'b';
        p32(spec.w); p32(spec.k); p32(spec.r); p32(spec.min_span); p32(spec.sketch ? 1u : 0u);
        memcpy(w, &nk, 8);
    }
    // key i starts at 31 + 24 i + 17 offs[i]: the records are formatted by the host pool, a range of keys per item
    const uint64_t per = 1u << 14;
    parallel_for((size_t)ceil_div<uint64_t>(nk, per), [&](size_t it) {
        const uint64_t k0 = it * per, k1 = std::min<uint64_t>(nk, k0 + per);
        uint8_t *w = buf.get() + 31 + k0 * 24 + offs[k0] * 17;
        for (uint64_t i = k0; i < k1; i++) {
            const uint64_t cnt = offs[i + 1] - offs[i];
            memcpy(w, &keys[2 * i], 16); memcpy(w + 16, &cnt, 8); w += 24;
            for (uint64_t j = offs[i]; j < offs[i + 1]; j++) {
                const pgr_frag_sig &sg = sigs[j];
                memcpy(w, &sg.frg_id, 4); memcpy(w + 4, &sg.sid, 4); memcpy(w + 8, &sg.bgn, 4); memcpy(w + 12, &sg.end, 4); w[16] = sg.ori;
                w += 17;
            }
        }
    });
    FILE *f = fopen(path, "wb");
    if (!f) { set_error("cannot create %s", path); return PGR_E_IO; }
    const size_t wr = fwrite(buf.get(), 1, total, f);
    const int cl = fclose(f);
    if (wr != total || cl != 0) { set_error("short write to %s", path); return PGR_E_IO; }
    return PGR_OK;
}

}  // namespace pgr

extern "C" {

pgr_b200_index *pgr_b200_index_new(const pgr_shmmr_spec *spec, int frg_id_mode, int device) {
    if (check_spec(spec) != PGR_OK) return nullptr;
    if (frg_id_mode != 0 && frg_id_mode != 1) { set_error("frg_id_mode must be 0 (FASTX) or 1 (AGC)"); return nullptr; }
    pgr_b200_ctx *ctx = pgr_b200_ctx_new(device < 0 ? default_device() : device);
    if (!ctx) return nullptr;
    pgr_b200_index *idx = new pgr_b200_index();
    idx->spec = *spec; idx->mode = frg_id_mode; idx->ctx = ctx;
    return idx;
}

void pgr_b200_index_free(pgr_b200_index *idx) {
    if (!idx) return;
    cudaSetDevice(idx->ctx->device);
    cudaStreamSynchronize(idx->ctx->stream);
    DevBuf *bufs[] = {&idx->tuples, &idx->ukeys, &idx->offsets, &idx->sigs, &idx->keysA, &idx->keysB, &idx->idxA, &idx->idxB, &idx->hist,
                      &idx->head, &idx->block_sum, &idx->block_prefix, &idx->d_sid, &idx->d_pair_off, &idx->d_frg_base, &idx->qtuples,
                      &idx->q_hit_begin, &idx->q_hit_count, &idx->scratch0, &idx->scratch1, &idx->scratch2, &idx->scratch3,
                      &idx->sid_count, &idx->hitsA, &idx->hitsB, &idx->seg_keys, &idx->seg_off, &idx->chain_f, &idx->chain_u, &idx->chain_b,
                      &idx->chain_seg, &idx->asm_prefix, &idx->asm_has, &idx->asm_out, &idx->asm_out2, &idx->sendbuf};
    for (auto b : bufs) b->release();
    if (idx->d2h_stream) cudaStreamDestroy(idx->d2h_stream);
    pgr_b200_ctx_free(idx->ctx);
    delete idx;
}

int pgr_b200_index_get_spec(const pgr_b200_index *idx, pgr_shmmr_spec *spec) {
    if (!idx || !spec) { set_error("NULL argument"); return PGR_E_ARG; }
    *spec = idx->spec;
    return PGR_OK;
}

int pgr_b200_index_add_batch(pgr_b200_index *idx, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens) {
    if (!idx || (n && (!sids || !seqs || !lens))) { set_error("NULL argument"); return PGR_E_ARG; }
    if (idx->staged) { set_error("a staged batch is pending: call pgr_b200_index_commit_batch first"); return PGR_E_ARG; }
    if (idx->gathered) { set_error("an index gathered from a multi-GPU build is read-only"); return PGR_E_ARG; }
    idx->finalized = false;
    idx->ord_in_batch = 0;
    const uint64_t t0 = idx->n_tuples;
    const uint32_t f0 = idx->n_frags;
    trace_mark("index_add_batch: begin");
    const int rc = index_batch_tuples(idx, n, sids, seqs, lens, false, nullptr, nullptr, nullptr);
    trace_mark("index_add_batch: shimmers + tuples");
    if (rc != PGR_OK) { idx->n_tuples = t0; idx->n_frags = f0; }   // a failed batch leaves the index as it was
    return rc;
}

// stage_batch for sequences that already live in device memory (layout rules of pgr_b200_ctx_set_device_seqs)
int pgr_b200_index_stage_device(pgr_b200_index *idx, const uint8_t *dev_base, size_t n, const uint32_t *sids, const uint64_t *offs,
                                const uint64_t *lens, uint64_t *n_frags_in_batch) {
    if (!idx || (n && (!sids || !dev_base || !offs || !lens))) { set_error("NULL argument"); return PGR_E_ARG; }
    if (idx->staged) { set_error("a staged batch is pending: call pgr_b200_index_commit_batch first"); return PGR_E_ARG; }
    pgr_b200_ctx *ctx = idx->ctx;
    idx->finalized = false;
    idx->staged_prev_frags = idx->n_frags;
    idx->staged_t0 = idx->n_tuples;
    idx->n_frags = 0;
    idx->ord_in_batch = 0;
    auto fail = [&](int rc) { idx->n_frags = idx->staged_prev_frags; idx->n_tuples = idx->staged_t0; return rc; };
    int rc = pgr_b200_ctx_set_device_seqs(ctx, dev_base, n, nullptr, offs, lens);
    if (rc != PGR_OK) return fail(rc);
    ctx->timer.reset();
    memset(ctx->counters, 0, sizeof ctx->counters);
    ctx->r0 = 0; ctx->rn = n;
    size_t ns = 0;
    if (n) {
        if ((rc = shmmrs_range(ctx, idx->spec, 0, &ns)) != PGR_OK) return fail(rc);
        uint64_t np = 0;
        if ((rc = emit_tuples(idx, sids, n, ns, false, nullptr, &np)) != PGR_OK) return fail(rc);
    }
    idx->staged_frags = idx->n_frags;
    if (n_frags_in_batch) *n_frags_in_batch = (idx->mode == 0) ? idx->staged_frags : 0;
    idx->staged = true;
    return PGR_OK;
}

// stage: the chunked, copy/compute-overlapped path of add_batch with fragment ids counted from 0; commit shifts the ids
// of the staged tuples by the number of fragments that precede this batch globally (known only after the ranks have
// exchanged their totals)
int pgr_b200_index_stage_batch(pgr_b200_index *idx, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens,
                               uint64_t *n_frags_in_batch) {
    if (!idx || (n && (!sids || !seqs || !lens))) { set_error("NULL argument"); return PGR_E_ARG; }
    if (idx->staged) { set_error("a staged batch is pending: call pgr_b200_index_commit_batch first"); return PGR_E_ARG; }
    idx->finalized = false;
    idx->staged_prev_frags = idx->n_frags;
    idx->staged_t0 = idx->n_tuples;
    idx->n_frags = 0;
    idx->ord_in_batch = 0;
    const int rc = index_batch_tuples(idx, n, sids, seqs, lens, false, nullptr, nullptr, nullptr);
    if (rc != PGR_OK) { idx->n_frags = idx->staged_prev_frags; idx->n_tuples = idx->staged_t0; return rc; }
    idx->staged_frags = idx->n_frags;
    if (n_frags_in_batch) *n_frags_in_batch = (idx->mode == 0) ? idx->staged_frags : 0;
    idx->staged = true;
    return PGR_OK;
}

int pgr_b200_index_commit_batch(pgr_b200_index *idx, uint32_t frag_base) {
    if (!idx || !idx->staged) { set_error("no staged batch"); return PGR_E_ARG; }
    PGR_CUDA(cudaSetDevice(idx->ctx->device));
    const uint64_t nt = idx->n_tuples - idx->staged_t0;
    if (idx->mode == 0 && frag_base && nt) {
        add_frg_base_kernel<<<(uint32_t)ceil_div<uint64_t>(nt, 256), 256, 0, idx->ctx->stream>>>(idx->tuples.as<FragTuple>() + idx->staged_t0, nt, frag_base);
        idx->launches += 1;
        PGR_CUDA(cudaGetLastError());
        PGR_CUDA(cudaStreamSynchronize(idx->ctx->stream));
    }
    idx->n_frags = (idx->mode == 0) ? frag_base + idx->staged_frags : 0;
    idx->staged = false;
    return PGR_OK;
}

int pgr_b200_index_finalize(pgr_b200_index *idx) {
    if (!idx) { set_error("idx is NULL"); return PGR_E_ARG; }
    if (idx->finalized) return PGR_OK;
    idx->sid_count_valid = false;
    pgr_b200_ctx *ctx = idx->ctx;
    PGR_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t n = idx->n_tuples;
    trace_mark("index_finalize: begin");
    if (n >= 0xFFFFFFF0ull) { set_error("more than 2^32 tuples on one device"); return PGR_E_LIMIT; }
    PGR_TRY(idx->offsets.ensure(sizeof(uint64_t) * (n + 2)));
    if (n == 0) {
        idx->n_keys = 0;
        PGR_CUDA(cudaMemsetAsync(idx->offsets.p, 0, sizeof(uint64_t), st));
        idx->finalized = true;
        return PGR_OK;
    }
    PGR_TRY(idx->keysA.ensure(n * sizeof(SortKey)));
    PGR_TRY(idx->idxA.ensure(n * sizeof(uint32_t)));
    trace_mark("index_finalize: alloc");
    const uint32_t g = (uint32_t)ceil_div<uint64_t>(n, 256);
    iota_kernel<<<g, 256, 0, st>>>(idx->idxA.as<uint32_t>(), n);
    if (idx->ord_sort) {
        // tuples of interleaved blocks (multi-GPU merge of several batches): minor key = the sequences' insertion ordinal
        tuple_ord_keys_kernel<<<g, 256, 0, st>>>(idx->tuples.as<FragTuple>(), n, idx->keysA.as<SortKey>());
        PGR_TRY(index_sort(idx, n, 0, 3));
        tuple_keys_perm_kernel<<<g, 256, 0, st>>>(idx->tuples.as<FragTuple>(), idx->idxA.as<uint32_t>(), n, idx->keysA.as<SortKey>());
        idx->launches += 1;
    } else {
        tuple_keys_kernel<<<g, 256, 0, st>>>(idx->tuples.as<FragTuple>(), n, idx->keysA.as<SortKey>());
    }
    idx->launches += 2;
    PGR_TRY(index_sort(idx, n, 0, 13));
    PGR_TRY(idx->sigs.ensure(n * sizeof(pgr_frag_sig)));
    PGR_TRY(idx->head.ensure(n));
    csr_gather_kernel<<<g, 256, 0, st>>>(idx->tuples.as<FragTuple>(), idx->keysA.as<SortKey>(), idx->idxA.as<uint32_t>(), n,
                                         idx->sigs.as<pgr_frag_sig>(), idx->head.as<uint8_t>());
    const uint32_t nb = (uint32_t)ceil_div<uint64_t>(n, CS_BLK);
    PGR_TRY(idx->block_sum.ensure(nb * sizeof(uint32_t)));
    PGR_TRY(idx->block_prefix.ensure((nb + 1) * sizeof(uint64_t)));
    csr_count_kernel<<<nb, CS_NT, 0, st>>>(idx->head.as<uint8_t>(), n, idx->block_sum.as<uint32_t>());
    block_scan_kernel<<<1, 1024, 0, st>>>(idx->block_sum.as<uint32_t>(), idx->block_prefix.as<uint64_t>(), nb);
    PGR_TRY(ctx->ensure_ctl(64));
    PGR_CUDA(cudaMemcpyAsync(ctx->h_ctl, idx->block_prefix.as<uint64_t>() + nb, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    PGR_CUDA(cudaStreamSynchronize(st));
    idx->n_keys = *(uint64_t *)ctx->h_ctl;
    PGR_TRY(idx->ukeys.ensure(idx->n_keys * sizeof(SortKey)));
    csr_write_kernel<<<nb, CS_NT, 0, st>>>(idx->head.as<uint8_t>(), idx->keysA.as<SortKey>(), n, idx->block_prefix.as<uint64_t>(),
                                           idx->ukeys.as<SortKey>(), idx->offsets.as<uint64_t>());
    set_u64_kernel<<<1, 1, 0, st>>>(idx->offsets.as<uint64_t>() + idx->n_keys, n);
    idx->launches += 5;
    PGR_CUDA(cudaGetLastError());
    PGR_CUDA(cudaStreamSynchronize(st));
    idx->finalized = true;
    trace_mark("index_finalize: sort + CSR");
    return PGR_OK;
}

int pgr_b200_index_counts(pgr_b200_index *idx, size_t *n_keys, size_t *n_sigs, uint32_t *n_frags) {
    if (!idx) { set_error("idx is NULL"); return PGR_E_ARG; }
    if (n_keys) { PGR_TRY(pgr_b200_index_finalize(idx)); *n_keys = idx->n_keys; }
    if (n_sigs) *n_sigs = idx->n_tuples;
    if (n_frags) *n_frags = idx->n_frags;
    return PGR_OK;
}

int pgr_b200_index_export_csr(pgr_b200_index *idx, uint64_t *keys, uint64_t *offsets, pgr_frag_sig *sigs) {
    if (!idx || !keys || !offsets || !sigs) { set_error("NULL argument"); return PGR_E_ARG; }
    PGR_TRY(pgr_b200_index_finalize(idx));
    PGR_CUDA(cudaSetDevice(idx->ctx->device));
    cudaStream_t st = idx->ctx->stream;
    if (idx->n_keys) PGR_CUDA(cudaMemcpyAsync(keys, idx->ukeys.p, idx->n_keys * sizeof(SortKey), cudaMemcpyDeviceToHost, st));
    PGR_CUDA(cudaMemcpyAsync(offsets, idx->offsets.p, (idx->n_keys + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    if (idx->n_tuples) PGR_CUDA(cudaMemcpyAsync(sigs, idx->sigs.p, idx->n_tuples * sizeof(pgr_frag_sig), cudaMemcpyDeviceToHost, st));
    PGR_CUDA(cudaStreamSynchronize(st));
    return PGR_OK;
}

int pgr_b200_index_tuples_device(pgr_b200_index *idx, void **dev_tuples, size_t *n_tuples) {
    if (!idx) { set_error("idx is NULL"); return PGR_E_ARG; }
    if (dev_tuples) *dev_tuples = idx->tuples.p;
    if (n_tuples) *n_tuples = idx->n_tuples;
    return PGR_OK;
}

int pgr_b200_index_set_tuples_device(pgr_b200_index *idx, const void *dev_tuples, size_t n_tuples) {
    if (!idx || (n_tuples && !dev_tuples)) { set_error("NULL argument"); return PGR_E_ARG; }
    PGR_CUDA(cudaSetDevice(idx->ctx->device));
    idx->finalized = false;
    if (dev_tuples != idx->tuples.p) {
        idx->n_tuples = 0;
        PGR_TRY(index_reserve_tuples(idx, n_tuples));
        if (n_tuples) PGR_CUDA(cudaMemcpyAsync(idx->tuples.p, dev_tuples, n_tuples * sizeof(FragTuple), cudaMemcpyDeviceToDevice, idx->ctx->stream));
        PGR_CUDA(cudaStreamSynchronize(idx->ctx->stream));
    }
    idx->n_tuples = n_tuples;
    return PGR_OK;
}

int pgr_b200_index_partition(pgr_b200_index *idx, size_t n_parts, const uint64_t *splitters, uint64_t *counts) {
    if (!idx || !counts || n_parts == 0 || n_parts > 256 || (n_parts > 1 && !splitters)) { set_error("bad argument"); return PGR_E_ARG; }
    pgr_b200_ctx *ctx = idx->ctx;
    PGR_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t n = idx->n_tuples;
    for (size_t i = 0; i < n_parts; i++) counts[i] = 0;
    if (n == 0) return PGR_OK;
    idx->finalized = false;
    PGR_TRY(idx->keysA.ensure(n * sizeof(SortKey)));
    PGR_TRY(idx->idxA.ensure(n * sizeof(uint32_t)));
    PGR_TRY(idx->scratch0.ensure(256 * sizeof(uint64_t) * 2));
    uint64_t *d_split = idx->scratch0.as<uint64_t>();
    unsigned long long *d_counts = (unsigned long long *)(d_split + 256);
    if (n_parts > 1) PGR_CUDA(cudaMemcpyAsync(d_split, splitters, (n_parts - 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    PGR_CUDA(cudaMemsetAsync(d_counts, 0, 256 * sizeof(uint64_t), st));
    const uint32_t g = (uint32_t)ceil_div<uint64_t>(n, 256);
    dest_keys_kernel<<<g, 256, 0, st>>>(idx->tuples.as<FragTuple>(), n, d_split, (uint32_t)(n_parts - 1), idx->keysA.as<SortKey>(), d_counts);
    iota_kernel<<<g, 256, 0, st>>>(idx->idxA.as<uint32_t>(), n);
    idx->launches += 2;
    PGR_TRY(index_sort(idx, n, 7, 7));
    DevBuf nb;
    PGR_TRY(nb.ensure(std::max<uint64_t>(n * sizeof(FragTuple), idx->tuples.cap)));
    gather_tuples_kernel<<<g, 256, 0, st>>>(idx->tuples.as<FragTuple>(), idx->idxA.as<uint32_t>(), n, nb.as<FragTuple>());
    idx->launches += 1;
    PGR_CUDA(cudaGetLastError());
    PGR_CUDA(cudaMemcpyAsync(counts, d_counts, n_parts * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    PGR_CUDA(cudaStreamSynchronize(st));
    idx->tuples.release();
    idx->tuples = nb;
    return PGR_OK;
}

// seq_db.rs:1291-1326 with keys ascending (canonical form; the reference writes FxHashMap iteration order)
int pgr_b200_index_write_mdb(pgr_b200_index *idx, const char *path) {
    if (!idx || !path) { set_error("NULL argument"); return PGR_E_ARG; }
    PGR_TRY(pgr_b200_index_finalize(idx));
    const uint64_t nk = idx->n_keys, ns = idx->n_tuples;
    std::vector<uint64_t> keys(2 * nk), offs(nk + 1);
    if (nk == 0) offs[0] = 0;
    uint64_t dk = 0; pgr_frag_sig ds;
    // the signatures (20 B each, the bulk of the map) come back into a page-locked pool buffer: PCIe rate, no zero fill
    pgr_frag_sig *sigs = ns ? (pgr_frag_sig *)result_alloc(ns * sizeof(pgr_frag_sig)) : &ds;
    if (!sigs) { set_error("out of host memory"); return PGR_E_ARG; }
    int rc = pgr_b200_index_export_csr(idx, nk ? keys.data() : &dk, offs.data(), sigs);
    if (rc == PGR_OK) rc = write_mdb_file(idx->spec, nk, keys.data(), offs.data(), sigs, path);
    if (ns) result_free(sigs);
    return rc;
}

// seq_db.rs:1328-1407: any key order is accepted; per-key vectors keep their file order
pgr_b200_index *pgr_b200_index_read_mdb(const char *path, int device) {
    if (!path) { set_error("path is NULL"); return nullptr; }
    FILE *f = fopen(path, "rb");
    if (!f) { set_error("cannot open %s", path); return nullptr; }
    std::vector<uint8_t> buf;
    {   // one read of the whole file when its size is known (a regular file), chunks otherwise
        long sz = -1;
        if (fseek(f, 0, SEEK_END) == 0) { sz = ftell(f); rewind(f); }
        if (sz > 0) {
            buf.resize((size_t)sz);
            const size_t got = fread(buf.data(), 1, buf.size(), f);
            buf.resize(got);
        }
        uint8_t tmp[1 << 16];
        size_t got;
        while ((got = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    }
    fclose(f);
    if (buf.size() < 31 || memcmp(buf.data(), "mdb", 3) != 0) { set_error("%s is not an .mdb file", path); return nullptr; }
    size_t c = 3;
    auto r32 = [&]() { uint32_t v; memcpy(&v, &buf[c], 4); c += 4; return v; };
    auto r64 = [&]() { uint64_t v; memcpy(&v, &buf[c], 8); c += 8; return v; };
    pgr_shmmr_spec spec;
    spec.w = r32(); spec.k = r32(); spec.r = r32(); spec.min_span = r32(); spec.sketch = r32() & 1u;
    const uint64_t nk = r64();
    if (nk > (buf.size() - c) / 24) { set_error("%s is truncated", path); return nullptr; }
    // pass 1 (sequential, headers only): where every key's records start and how many signatures precede it
    std::vector<uint64_t> key_at(nk), sig_before(nk + 1);
    uint64_t n_sig = 0;
    for (uint64_t i = 0; i < nk; i++) {
        if (c + 24 > buf.size()) { set_error("%s is truncated", path); return nullptr; }
        key_at[i] = c;
        uint64_t vl; memcpy(&vl, &buf[c + 16], 8);
        c += 24;
        if (vl > (buf.size() - c) / 17) { set_error("%s is truncated", path); return nullptr; }   // no overflow: compare counts
        sig_before[i] = n_sig;
        n_sig += vl; c += 17 * vl;
    }
    sig_before[nk] = n_sig;
    if (n_sig >= 0xFFFFFFF0ull) { set_error("%s holds more than 2^32 signatures", path); return nullptr; }
    // pass 2 (host pool): the records of a range of keys -> tuples at their final places
    std::unique_ptr<FragTuple[]> tuples(new FragTuple[std::max<uint64_t>(1, n_sig)]);
    const uint64_t per = 1u << 13;
    std::vector<uint64_t> max_of((size_t)ceil_div<uint64_t>(std::max<uint64_t>(nk, 1), per), 0);
    parallel_for(max_of.size(), [&](size_t it) {
        const uint64_t k0 = it * per, k1 = std::min<uint64_t>(nk, k0 + per);
        uint64_t mx = 0;
        for (uint64_t i = k0; i < k1; i++) {
            const uint8_t *q = &buf[key_at[i]];
            uint64_t h0, h1;
            memcpy(&h0, q, 8); memcpy(&h1, q + 8, 8);
            q += 24;
            FragTuple *t = tuples.get() + sig_before[i];
            for (uint64_t j = sig_before[i]; j < sig_before[i + 1]; j++, t++, q += 17) {
                t->h0 = h0; t->h1 = h1;
                memcpy(&t->frg_id, q, 4); memcpy(&t->sid, q + 4, 4); memcpy(&t->bgn, q + 8, 4); memcpy(&t->end, q + 12, 4);
                t->ori = q[16]; t->ord = 0;
                mx = std::max<uint64_t>(mx, (uint64_t)t->frg_id + 1);
            }
        }
        max_of[it] = mx;
    });
    uint64_t max_frg = 0;
    for (uint64_t m : max_of) max_frg = std::max(max_frg, m);
    pgr_b200_index *idx = pgr_b200_index_new(&spec, 0, device);
    if (!idx) return nullptr;
    if (index_reserve_tuples(idx, n_sig) != PGR_OK) { pgr_b200_index_free(idx); return nullptr; }
    if (n_sig) {
        // through the page-locked ring (the tuple array is plain memory): PCIe rate instead of the driver's staging
        pgr_b200_ctx *ctx = idx->ctx;
        const size_t bytes = n_sig * sizeof(FragTuple);
        int rc = PGR_OK;
        if (bytes >= PACK_MIN_BYTES && (ctx->pack || (ctx->pack = pack_ring_acquire(ctx->device)))) {
            rc = upload_raw_staged(ctx->pack, (uint8_t *)idx->tuples.p, (const uint8_t *)tuples.get(), bytes, ctx->stream);
            if (rc == PGR_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = PGR_E_CUDA;
        } else if (cudaMemcpy(idx->tuples.p, tuples.get(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) rc = PGR_E_CUDA;
        if (rc != PGR_OK) { set_error("H2D of the .mdb records failed"); pgr_b200_index_free(idx); return nullptr; }
    }
    idx->n_tuples = n_sig;
    idx->n_frags = (uint32_t)std::min<uint64_t>(max_frg, 0xFFFFFFFFull);   // fragment ids in use (no prefix/suffix records in an .mdb)
    idx->from_mdb = true;
    if (pgr_b200_index_finalize(idx) != PGR_OK) { pgr_b200_index_free(idx); return nullptr; }
    return idx;
}

}  // extern "C"
