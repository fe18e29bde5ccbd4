// query.cu — raw_query_fragment (seq_db.rs:1200-1228), query_fragment_to_hps (aln.rs:147-242 via ext.rs:252-282),
// sparse_aln (aln.rs:12-142) and frag_map_to_adj_list (seq_db.rs:876-944) against the device-resident index.
#include <algorithm>
#include <cstring>

#include "index.cuh"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <memory>
#include "query_kernels.cuh"

using namespace pgr;

namespace {

int scan_u32(pgr_b200_index *idx, const uint32_t *v, uint64_t n, uint64_t *out /* n+1 */, uint64_t *total) {
    cudaStream_t st = idx->ctx->stream;
    if (n == 0) {
        PGR_CUDA(cudaMemsetAsync(out, 0, sizeof(uint64_t), st));
        if (total) *total = 0;
        return PGR_OK;
    }
    const uint32_t nb = (uint32_t)ceil_div<uint64_t>(n + 1, SC_BLK);  // +1: slot n (the total) must be owned by a block
    PGR_TRY(idx->block_sum.ensure(nb * sizeof(uint32_t)));
    PGR_TRY(idx->block_prefix.ensure((nb + 1) * sizeof(uint64_t)));
    scan_reduce_kernel<<<nb, SC_NT, 0, st>>>(v, n, idx->block_sum.as<uint32_t>());
    block_scan_kernel<<<1, 1024, 0, st>>>(idx->block_sum.as<uint32_t>(), idx->block_prefix.as<uint64_t>(), nb);
    scan_apply_kernel<<<nb, SC_NT, 0, st>>>(v, n, idx->block_prefix.as<uint64_t>(), out);
    idx->launches += 3;
    PGR_CUDA(cudaGetLastError());
    if (total) {
        PGR_TRY(idx->ctx->ensure_ctl(64));
        PGR_CUDA(cudaMemcpyAsync(idx->ctx->h_ctl, out + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaStreamSynchronize(st));
        *total = *(uint64_t *)idx->ctx->h_ctl;
    }
    return PGR_OK;
}

// per-signature same-sid counts of the index (computed once per finalize)
int ensure_sid_count(pgr_b200_index *idx) {
    PGR_TRY(pgr_b200_index_finalize(idx));
    if (idx->sid_count_valid) return PGR_OK;
    PGR_TRY(idx->sid_count.ensure(std::max<uint64_t>(1, idx->n_tuples) * sizeof(uint32_t)));
    if (idx->n_keys) {
        sid_count_kernel<<<(uint32_t)ceil_div<uint64_t>(idx->n_keys, 128), 128, 0, idx->ctx->stream>>>(
            idx->sigs.as<pgr_frag_sig>(), idx->offsets.as<uint64_t>(), idx->n_keys, idx->sid_count.as<uint32_t>());
        idx->launches += 1;
        PGR_CUDA(cudaGetLastError());
    }
    idx->sid_count_valid = true;
    return PGR_OK;
}

// query shimmers -> query tuples (device, idx->qtuples) + pair offsets per query (host) + lookup
int query_pairs_and_lookup(pgr_b200_index *idx, size_t n_q, const uint8_t *const *seqs, const size_t *lens, uint64_t *n_qp_out,
                           std::vector<uint64_t> *qp_off_out) {
    PGR_TRY(pgr_b200_index_finalize(idx));
    cudaStream_t st = idx->ctx->stream;
    std::vector<uint32_t> qids(n_q);
    for (size_t i = 0; i < n_q; i++) qids[i] = (uint32_t)i;
    std::vector<uint64_t> per_seq;
    uint64_t n_qp = 0;
    PGR_TRY(index_batch_tuples(idx, n_q, qids.data(), seqs, lens, true, &idx->qtuples, &n_qp, &per_seq));
    qp_off_out->assign(n_q + 1, 0);
    for (size_t i = 0; i < n_q; i++) (*qp_off_out)[i + 1] = (*qp_off_out)[i] + per_seq[i];
    *n_qp_out = n_qp;
    PGR_TRY(idx->q_hit_begin.ensure(std::max<uint64_t>(1, n_qp) * sizeof(uint64_t)));
    PGR_TRY(idx->q_hit_count.ensure(std::max<uint64_t>(1, n_qp) * sizeof(uint32_t)));
    if (n_qp) {
        lookup_kernel<<<(uint32_t)ceil_div<uint64_t>(n_qp, 256), 256, 0, st>>>(idx->qtuples.as<FragTuple>(), n_qp, idx->ukeys.as<SortKey>(),
                                                                              idx->offsets.as<uint64_t>(), idx->n_keys,
                                                                              idx->q_hit_begin.as<uint64_t>(), idx->q_hit_count.as<uint32_t>());
        idx->launches += 1;
        PGR_CUDA(cudaGetLastError());
    }
    return PGR_OK;
}

template <class T>
T *host_dup(const std::vector<T> &v) {
    T *p = (T *)malloc(std::max<size_t>(1, v.size()) * sizeof(T));
    if (p && !v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

}  // namespace

extern "C" {

// replaces seq_db::raw_query_fragment(frag_map, query, spec) -> Vec<FragmentHit>
int pgr_b200_raw_query(pgr_b200_index *idx, const uint8_t *seq, size_t len, pgr_query_pair **pairs, size_t *n_pairs, uint64_t **hit_off,
                       pgr_frag_sig **hits) {
    if (!idx || !pairs || !n_pairs || !hit_off || !hits) { set_error("NULL argument"); return PGR_E_ARG; }
    PGR_CUDA(cudaSetDevice(idx->ctx->device));
    cudaStream_t st = idx->ctx->stream;
    const uint8_t *sp[1] = {seq};
    const size_t ln[1] = {len};
    uint64_t n_qp = 0;
    std::vector<uint64_t> qp_off;
    PGR_TRY(query_pairs_and_lookup(idx, 1, sp, ln, &n_qp, &qp_off));
    std::vector<FragTuple> qt(n_qp);
    std::vector<uint64_t> hb(n_qp);
    std::vector<uint32_t> hc(n_qp);
    if (n_qp) {
        PGR_CUDA(cudaMemcpyAsync(qt.data(), idx->qtuples.p, n_qp * sizeof(FragTuple), cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaMemcpyAsync(hb.data(), idx->q_hit_begin.p, n_qp * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaMemcpyAsync(hc.data(), idx->q_hit_count.p, n_qp * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaStreamSynchronize(st));
    }
    std::vector<pgr_query_pair> qp(n_qp);
    std::vector<uint64_t> off(n_qp + 1, 0);
    for (uint64_t i = 0; i < n_qp; i++) {
        memset(&qp[i], 0, sizeof qp[i]);
        qp[i].h0 = qt[i].h0; qp[i].h1 = qt[i].h1; qp[i].bgn = qt[i].bgn; qp[i].end = qt[i].end; qp[i].ori = (uint8_t)qt[i].ori;
        off[i + 1] = off[i] + hc[i];
    }
    std::vector<pgr_frag_sig> hs(off[n_qp]);
    for (uint64_t i = 0; i < n_qp; i++)
        if (hc[i]) PGR_CUDA(cudaMemcpyAsync(hs.data() + off[i], idx->sigs.as<pgr_frag_sig>() + hb[i], hc[i] * sizeof(pgr_frag_sig), cudaMemcpyDeviceToHost, st));
    PGR_CUDA(cudaStreamSynchronize(st));
    *pairs = host_dup(qp);
    *n_pairs = n_qp;
    *hit_off = host_dup(off);
    *hits = host_dup(hs);
    return PGR_OK;
}

// ---- .mdb-resident look-ups (seq_db.rs:1230-1269, :1409-1471) ---------------------------------------------------------
}  // extern "C"

struct pgr_b200_mdb_map {
    pgr_shmmr_spec spec;
    int fd = -1;
    const uint8_t *base = nullptr;    // the memory-mapped file
    size_t bytes = 0;
    // keys sorted ascending with the file position and length of their signature vectors (ShmmrIndexFileLocation)
    std::vector<uint64_t> h0, h1, at;
    std::vector<uint32_t> cnt;
    size_t n_sigs = 0;
};

extern "C" {

pgr_b200_mdb_map *pgr_b200_mdb_map_open(const char *path) {
    if (!path) { set_error("path is NULL"); return nullptr; }
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { set_error("cannot open %s", path); return nullptr; }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size < 31) { close(fd); set_error("%s is not an .mdb file", path); return nullptr; }
    void *mp = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (mp == MAP_FAILED) { close(fd); set_error("mmap of %s failed", path); return nullptr; }
    std::unique_ptr<pgr_b200_mdb_map> m(new pgr_b200_mdb_map());
    m->fd = fd; m->base = (const uint8_t *)mp; m->bytes = (size_t)sb.st_size;
    auto fail = [&](const char *why) -> pgr_b200_mdb_map * { set_error("%s: %s", path, why); munmap(mp, m->bytes); close(fd); return nullptr; };
    const uint8_t *b = m->base;
    if (memcmp(b, "mdb", 3) != 0) return fail("not an .mdb file");
    size_t c = 3;
    auto r32 = [&]() { uint32_t v; memcpy(&v, b + c, 4); c += 4; return v; };
    m->spec.w = r32(); m->spec.k = r32(); m->spec.r = r32(); m->spec.min_span = r32(); m->spec.sketch = r32() & 1u;
    uint64_t nk; memcpy(&nk, b + c, 8); c += 8;
    if (nk > (m->bytes - c) / 24) return fail("truncated");
    // header pass: the 24-byte key records only; the signatures are never touched here
    std::vector<uint64_t> k0(nk), k1(nk), at(nk);
    std::vector<uint32_t> cn(nk);
    for (uint64_t i = 0; i < nk; i++) {
        if (c + 24 > m->bytes) return fail("truncated");
        uint64_t vl;
        memcpy(&k0[i], b + c, 8); memcpy(&k1[i], b + c + 8, 8); memcpy(&vl, b + c + 16, 8);
        c += 24;
        if (vl > (m->bytes - c) / 17 || vl > 0xFFFFFFFFull) return fail("truncated");
        at[i] = c; cn[i] = (uint32_t)vl;
        c += 17 * vl;
        m->n_sigs += vl;
    }
    // the file may list its keys in any order (the reference writes hash-map order): sort the table by key
    std::vector<uint32_t> ord(nk);
    for (uint64_t i = 0; i < nk; i++) ord[i] = (uint32_t)i;
    bool sorted = true;
    for (uint64_t i = 1; i < nk && sorted; i++) sorted = k0[i - 1] < k0[i] || (k0[i - 1] == k0[i] && k1[i - 1] < k1[i]);
    if (!sorted) std::sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) { return k0[x] != k0[y] ? k0[x] < k0[y] : k1[x] < k1[y]; });
    m->h0.resize(nk); m->h1.resize(nk); m->at.resize(nk); m->cnt.resize(nk);
    for (uint64_t i = 0; i < nk; i++) { const uint32_t j = ord[i]; m->h0[i] = k0[j]; m->h1[i] = k1[j]; m->at[i] = at[j]; m->cnt[i] = cn[j]; }
    return m.release();
}

void pgr_b200_mdb_map_close(pgr_b200_mdb_map *m) {
    if (!m) return;
    if (m->base) munmap((void *)m->base, m->bytes);
    if (m->fd >= 0) close(m->fd);
    delete m;
}

int pgr_b200_mdb_map_info(const pgr_b200_mdb_map *m, pgr_shmmr_spec *spec, size_t *n_keys, size_t *n_sigs) {
    if (!m) { set_error("NULL argument"); return PGR_E_ARG; }
    if (spec) *spec = m->spec;
    if (n_keys) *n_keys = m->h0.size();
    if (n_sigs) *n_sigs = m->n_sigs;
    return PGR_OK;
}

// raw_query_fragment_from_mmap_midx: shimmers on the device, pairs with the strict '<' rule (seq_db.rs:1243-1253), table look-up,
// signatures decoded from the 17-byte records of the map (:1473-1504)
int pgr_b200_raw_query_mmap(pgr_b200_mdb_map *m, const uint8_t *seq, size_t len, pgr_query_pair **pairs, size_t *n_pairs, uint64_t **hit_off,
                            pgr_frag_sig **hits) {
    if (!m || !pairs || !n_pairs || !hit_off || !hits) { set_error("NULL argument"); return PGR_E_ARG; }
    pgr_mm128 *mm = nullptr;
    size_t n_mm = 0;
    PGR_TRY(pgr_b200_sequence_to_shmmrs(0, seq, len, &m->spec, 0, &mm, &n_mm));
    const size_t np = n_mm > 1 ? n_mm - 1 : 0;
    std::vector<pgr_query_pair> qp(np);
    std::vector<uint64_t> off(np + 1, 0);
    std::vector<size_t> slot(np, (size_t)-1);
    const size_t nk = m->h0.size();
    for (size_t i = 0; i < np; i++) {
        const uint64_t s0 = mm[i].x >> 8, s1 = mm[i + 1].x >> 8;
        const bool fwd = s0 < s1;
        pgr_query_pair &q = qp[i];
        memset(&q, 0, sizeof q);
        q.h0 = fwd ? s0 : s1; q.h1 = fwd ? s1 : s0; q.ori = fwd ? 0 : 1;
        q.bgn = ((uint32_t)(mm[i].y & 0xFFFFFFFFu) >> 1) + 1;
        q.end = ((uint32_t)(mm[i + 1].y & 0xFFFFFFFFu) >> 1) + 1;
        size_t lo = 0, hi = nk;
        while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (m->h0[mid] < q.h0 || (m->h0[mid] == q.h0 && m->h1[mid] < q.h1)) lo = mid + 1; else hi = mid; }
        if (lo < nk && m->h0[lo] == q.h0 && m->h1[lo] == q.h1) { slot[i] = lo; off[i + 1] = off[i] + m->cnt[lo]; } else off[i + 1] = off[i];
    }
    pgr_b200_free(mm);
    std::vector<pgr_frag_sig> hs(off[np]);
    for (size_t i = 0; i < np; i++) {
        if (slot[i] == (size_t)-1) continue;
        const uint8_t *r = m->base + m->at[slot[i]];
        pgr_frag_sig *o = hs.data() + off[i];
        for (uint32_t j = 0; j < m->cnt[slot[i]]; j++, r += 17, o++) {
            memset(o, 0, sizeof *o);
            memcpy(&o->frg_id, r, 4); memcpy(&o->sid, r + 4, 4); memcpy(&o->bgn, r + 8, 4); memcpy(&o->end, r + 12, 4); o->ori = r[16];
        }
    }
    *pairs = host_dup(qp);
    *n_pairs = np;
    *hit_off = host_dup(off);
    *hits = host_dup(hs);
    return PGR_OK;
}

// query_fragment_to_hps_from_mmap_file for a batch: sub-index of the hit keys, then the ordinary batch
int pgr_b200_query_batch_mmap(pgr_b200_mdb_map *m, int device, size_t n_q, const uint8_t *const *seqs, const size_t *lens,
                              const pgr_query_params *prm, pgr_query_result **out) {
    if (!m || !prm || !out || (n_q && (!seqs || !lens))) { set_error("NULL argument"); return PGR_E_ARG; }
    // 1. shimmers of every query (device), the distinct keys of their pairs (strict '<' rule as in the query path)
    std::vector<uint32_t> rids(n_q);
    for (size_t q = 0; q < n_q; q++) rids[q] = (uint32_t)q;
    std::vector<size_t> offs(n_q + 1, 0);
    pgr_mm128 *mm = nullptr;
    PGR_TRY(pgr_b200_shmmrs_batch(n_q, rids.data(), seqs, lens, &m->spec, 0, &mm, offs.data()));   // on the calling thread's default device
    std::vector<std::pair<uint64_t, uint64_t>> keys;
    for (size_t q = 0; q < n_q; q++)
        for (size_t i = offs[q]; i + 1 < offs[q + 1]; i++) {
            const uint64_t s0 = mm[i].x >> 8, s1 = mm[i + 1].x >> 8;
            keys.emplace_back(std::min(s0, s1), std::max(s0, s1));
        }
    pgr_b200_free(mm);
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    // 2. their signature vectors out of the map -> tuples (file order inside a key, as read_mdb keeps it)
    const size_t nk = m->h0.size();
    std::vector<FragTuple> tuples;
    uint64_t max_frg = 0;
    for (const auto &k : keys) {
        size_t lo = 0, hi = nk;
        while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (m->h0[mid] < k.first || (m->h0[mid] == k.first && m->h1[mid] < k.second)) lo = mid + 1; else hi = mid; }
        if (lo >= nk || m->h0[lo] != k.first || m->h1[lo] != k.second) continue;
        const uint8_t *r = m->base + m->at[lo];
        for (uint32_t j = 0; j < m->cnt[lo]; j++, r += 17) {
            FragTuple t;
            t.h0 = k.first; t.h1 = k.second;
            memcpy(&t.frg_id, r, 4); memcpy(&t.sid, r + 4, 4); memcpy(&t.bgn, r + 8, 4); memcpy(&t.end, r + 12, 4);
            t.ori = r[16]; t.ord = 0;
            max_frg = std::max<uint64_t>(max_frg, (uint64_t)t.frg_id + 1);
            tuples.push_back(t);
        }
    }
    // 3. temporary index of those keys, the ordinary batch on it
    pgr_b200_index *idx = pgr_b200_index_new(&m->spec, 0, device);
    if (!idx) return PGR_E_NO_DEVICE;
    int rc = index_reserve_tuples(idx, std::max<size_t>(1, tuples.size()));
    if (rc == PGR_OK && !tuples.empty() &&
        cudaMemcpy(idx->tuples.p, tuples.data(), tuples.size() * sizeof(FragTuple), cudaMemcpyHostToDevice) != cudaSuccess) { set_error("H2D of the hit keys failed"); rc = PGR_E_CUDA; }
    if (rc == PGR_OK) {
        idx->n_tuples = tuples.size();
        idx->n_frags = (uint32_t)std::min<uint64_t>(max_frg, 0xFFFFFFFFull);
        idx->from_mdb = true;
        rc = pgr_b200_index_finalize(idx);
    }
    if (rc == PGR_OK) rc = pgr_b200_query_batch(idx, n_q, seqs, lens, prm, out);
    pgr_b200_index_free(idx);
    return rc;
}

void pgr_b200_query_result_free(pgr_query_result *r) {
    if (!r) return;
    result_free(r->q_target_off); result_free(r->target_sid); result_free(r->target_chain_off); result_free(r->chain_score);
    result_free(r->chain_hit_off); result_free(r->hits);
    free(r);
}

}  // extern "C"

// device half of query_fragment_to_hps for queries [0, n_q) of `seqs`: shimmers, pairs, look-up, filters, expansion, per-query
// sort, chaining and the nested result arrays, left in `asm_out` (QueryGroupOut points into it).  The offsets it writes are
// already global: chain_base / hit_base = chains / hits of the groups before this one.
struct QueryGroupOut {
    uint64_t n_targets = 0, n_chains = 0, n_hits = 0;
    const uint32_t *target_sid = nullptr, *target_qid = nullptr;
    const uint64_t *target_chain_off = nullptr, *chain_hit_off = nullptr;
    const float *chain_score = nullptr;
    const pgr_hit_pair *hits = nullptr;
};
static int query_group(pgr_b200_index *idx, size_t n_q, const uint8_t *const *seqs, const size_t *lens, const pgr_query_params *prm, DevBuf &asm_out,
                       uint64_t chain_base, uint64_t hit_base, QueryGroupOut *out) {
    pgr_b200_ctx *ctx = idx->ctx;
    cudaStream_t st = ctx->stream;
    *out = QueryGroupOut();
    uint64_t n_qp = 0;
    std::vector<uint64_t> qp_off;
    PGR_TRY(query_pairs_and_lookup(idx, n_q, seqs, lens, &n_qp, &qp_off));
    trace_mark("query_batch: shimmers+pairs+lookup");

    QueryFilter f;
    f.max_count = prm->max_count < 0 ? 128u : (uint32_t)prm->max_count;
    f.max_count_query = prm->max_count_query < 0 ? 128u : (uint32_t)prm->max_count_query;
    f.max_count_target = prm->max_count_target < 0 ? 128u : (uint32_t)prm->max_count_target;
    const uint32_t max_span = prm->max_aln_span < 0 ? 8u : (uint32_t)prm->max_aln_span;

    uint64_t n_hits = 0;
    if (n_qp) {
        // per-pair statistics, filters, expansion
        PGR_TRY(idx->scratch0.ensure((n_q + 1) * sizeof(uint64_t)));
        PGR_CUDA(cudaMemcpyAsync(idx->scratch0.p, qp_off.data(), (n_q + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        PGR_TRY(idx->scratch1.ensure(n_qp * sizeof(uint32_t) * 2));
        uint32_t *qcount = idx->scratch1.as<uint32_t>(), *nh = qcount + n_qp;
        const uint32_t g = (uint32_t)ceil_div<uint64_t>(n_qp, 256);
        qpair_count_kernel<<<g, 256, 0, st>>>(idx->qtuples.as<FragTuple>(), n_qp, idx->scratch0.as<uint64_t>(), qcount);
        hit_count_kernel<<<g, 256, 0, st>>>(n_qp, qcount, idx->q_hit_begin.as<uint64_t>(), idx->q_hit_count.as<uint32_t>(),
                                            idx->sid_count.as<uint32_t>(), f, nh);
        idx->launches += 2;
        PGR_TRY(idx->scratch2.ensure((n_qp + 1) * sizeof(uint64_t)));
        PGR_TRY(scan_u32(idx, nh, n_qp, idx->scratch2.as<uint64_t>(), &n_hits));
        if (n_hits >= 0xFFFFFFF0ull) { set_error("more than 2^32 hit pairs in one batch"); return PGR_E_LIMIT; }
        if (n_hits) {
            PGR_TRY(idx->hitsA.ensure(n_hits * sizeof(HitRec)));
            PGR_TRY(idx->hitsB.ensure(n_hits * sizeof(HitRec)));
            PGR_TRY(idx->keysA.ensure(n_hits * sizeof(SortKey)));
            PGR_TRY(idx->idxA.ensure(n_hits * sizeof(uint32_t)));
            hit_expand_kernel<<<g, 256, 0, st>>>(idx->qtuples.as<FragTuple>(), n_qp, qcount, idx->q_hit_begin.as<uint64_t>(),
                                                 idx->q_hit_count.as<uint32_t>(), idx->sid_count.as<uint32_t>(), idx->sigs.as<pgr_frag_sig>(), f,
                                                 idx->scratch2.as<uint64_t>(), idx->hitsA.as<HitRec>(), idx->keysA.as<SortKey>());
            const uint32_t gh = (uint32_t)ceil_div<uint64_t>(n_hits, 256);
            idx->launches += 1;
            trace_mark("query_batch: filters+expand");
            PGR_TRY(idx->head.ensure(n_hits));
            // The hits come out query by query (pairs of a query are contiguous, expansion keeps their order): what is left is
            // a stable sort by target sid inside every query.  One CTA per query does it in shared memory; a batch with a
            // query of more than QS_CAP hits takes the global radix sort instead.
            PGR_TRY(idx->scratch3.ensure((n_q + 1) * sizeof(uint64_t)));
            uint64_t *hq_off = idx->scratch3.as<uint64_t>();
            hit_query_offsets_kernel<<<(uint32_t)ceil_div<uint64_t>(n_q + 1, 256), 256, 0, st>>>(idx->scratch0.as<uint64_t>(), idx->scratch2.as<uint64_t>(), n_q, hq_off);
            std::vector<uint64_t> h_hq(n_q + 1);
            PGR_CUDA(cudaMemcpyAsync(h_hq.data(), hq_off, (n_q + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            PGR_CUDA(cudaStreamSynchronize(st));
            uint64_t max_q_hits = 0;
            for (size_t q = 0; q < n_q; q++) max_q_hits = std::max(max_q_hits, h_hq[q + 1] - h_hq[q]);
            static const bool force_global_sort = getenv("PGR_B200_QUERY_GLOBAL_SORT") != nullptr;   // A/B aid
            if (max_q_hits <= QS_CAP && !force_global_sort) {
                uint32_t m = 1;
                while (m < max_q_hits) m <<= 1;
                static bool attr_set = false;
                if (!attr_set) { PGR_CUDA(cudaFuncSetAttribute(query_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, QS_CAP * 8)); attr_set = true; }
                query_sort_kernel<<<(uint32_t)n_q, QS_NT, (size_t)m * 8, st>>>(idx->hitsA.as<HitRec>(), hq_off, idx->hitsB.as<HitRec>(), idx->head.as<uint8_t>());
                hit_keys_kernel<<<gh, 256, 0, st>>>(idx->hitsB.as<HitRec>(), n_hits, idx->keysA.as<SortKey>());
                idx->launches += 3;
                PGR_CUDA(cudaGetLastError());
            } else {
                iota_kernel<<<gh, 256, 0, st>>>(idx->idxA.as<uint32_t>(), n_hits);
                // stable sort by (qid, sid): bytes 0..3 of k1 (sid) then bytes 0..3 of k0 (qid)
                PGR_TRY(index_sort(idx, n_hits, 0, 3));
                PGR_TRY(index_sort(idx, n_hits, 7, 10));
                hit_gather_kernel<<<gh, 256, 0, st>>>(idx->hitsA.as<HitRec>(), idx->keysA.as<SortKey>(), idx->idxA.as<uint32_t>(), n_hits,
                                                      idx->hitsB.as<HitRec>(), idx->head.as<uint8_t>());
                idx->launches += 2;
            }
            // segments = runs of equal (qid, sid)
            const uint32_t nb = (uint32_t)ceil_div<uint64_t>(n_hits, CS_BLK);
            PGR_TRY(idx->block_sum.ensure(nb * sizeof(uint32_t)));
            PGR_TRY(idx->block_prefix.ensure((nb + 1) * sizeof(uint64_t)));
            csr_count_kernel<<<nb, CS_NT, 0, st>>>(idx->head.as<uint8_t>(), n_hits, idx->block_sum.as<uint32_t>());
            block_scan_kernel<<<1, 1024, 0, st>>>(idx->block_sum.as<uint32_t>(), idx->block_prefix.as<uint64_t>(), nb);
            PGR_TRY(ctx->ensure_ctl(64));
            PGR_CUDA(cudaMemcpyAsync(ctx->h_ctl, idx->block_prefix.as<uint64_t>() + nb, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            PGR_CUDA(cudaStreamSynchronize(st));
            const uint64_t n_seg = *(uint64_t *)ctx->h_ctl;
            PGR_TRY(idx->seg_keys.ensure(n_seg * sizeof(SortKey)));
            PGR_TRY(idx->seg_off.ensure((n_seg + 1) * sizeof(uint64_t)));
            csr_write_kernel<<<nb, CS_NT, 0, st>>>(idx->head.as<uint8_t>(), idx->keysA.as<SortKey>(), n_hits, idx->block_prefix.as<uint64_t>(),
                                                   idx->seg_keys.as<SortKey>(), idx->seg_off.as<uint64_t>());
            set_u64_kernel<<<1, 1, 0, st>>>(idx->seg_off.as<uint64_t>() + n_seg, n_hits);
            idx->launches += 5;
            trace_mark("query_batch: sort+segments");
            // chaining
            PGR_TRY(idx->chain_f.ensure(n_hits * sizeof(float) * 2));
            PGR_TRY(idx->chain_u.ensure(n_hits * sizeof(uint32_t) * 5));
            PGR_TRY(idx->chain_b.ensure(n_hits * 2));
            PGR_TRY(idx->chain_seg.ensure(n_seg * sizeof(uint32_t) * 3));
            ChainParams cp;
            cp.hits = idx->hitsB.as<HitRec>(); cp.seg_off = idx->seg_off.as<uint64_t>(); cp.n_seg = n_seg;
            cp.max_span = max_span; cp.penalty = prm->penalty; cp.has_gap = prm->max_gap >= 0 ? 1u : 0u;
            cp.max_gap = prm->max_gap >= 0 ? (float)(uint32_t)prm->max_gap : 0.0f; cp.oriented = prm->oriented ? 1u : 0u;
            cp.v_s = idx->chain_f.as<float>(); cp.out_score = cp.v_s + n_hits;
            uint32_t *u = idx->chain_u.as<uint32_t>();
            cp.best_pre = (int32_t *)u; cp.cls_first = u + n_hits; cp.cls_last = u + 2 * n_hits; cp.order = u + 3 * n_hits; cp.out_idx = u + 4 * n_hits;
            cp.visited = idx->chain_b.as<uint8_t>(); cp.out_start = cp.visited + n_hits;
            cp.seg_n_out = idx->chain_seg.as<uint32_t>(); cp.seg_n_chains = cp.seg_n_out + n_seg; cp.seg_err = cp.seg_n_out + 2 * n_seg;
            chain_kernel<<<(uint32_t)ceil_div<uint64_t>(n_seg, 64), 64, 0, st>>>(cp);
            idx->launches += 1;
            PGR_CUDA(cudaGetLastError());
            trace_mark("query_batch: chain kernel");
            // nested result arrays on the device, then one D2H per array into (pinned) result buffers
            uint64_t tot_hits = 0, tot_chains = 0, tot_targets = 0;
            PGR_TRY(idx->asm_prefix.ensure(3 * (n_seg + 1) * sizeof(uint64_t)));
            PGR_TRY(idx->asm_has.ensure(n_seg * sizeof(uint32_t)));
            uint64_t *hit_prefix = idx->asm_prefix.as<uint64_t>(), *chain_prefix = hit_prefix + (n_seg + 1), *target_prefix = chain_prefix + (n_seg + 1);
            seg_has_kernel<<<(uint32_t)ceil_div<uint64_t>(n_seg, 256), 256, 0, st>>>(cp.seg_n_out, n_seg, idx->asm_has.as<uint32_t>());
            PGR_TRY(scan_u32(idx, cp.seg_n_out, n_seg, hit_prefix, &tot_hits));
            PGR_TRY(scan_u32(idx, cp.seg_n_chains, n_seg, chain_prefix, &tot_chains));
            PGR_TRY(scan_u32(idx, idx->asm_has.as<uint32_t>(), n_seg, target_prefix, &tot_targets));
            // reference behaviour check: a segment whose scores are all <= 0 would never terminate in aln.rs:108-131
            {
                std::vector<uint32_t> err(n_seg);
                PGR_CUDA(cudaMemcpyAsync(err.data(), cp.seg_err, n_seg * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                PGR_CUDA(cudaStreamSynchronize(st));
                for (uint64_t s = 0; s < n_seg; s++)
                    if (err[s]) { set_error("sparse_aln: all scores <= 0 (the reference would not terminate)"); return PGR_E_ASSERT; }
            }
            PGR_TRY(asm_out.ensure(tot_targets * 16 + (tot_targets + 1) * 8 + tot_chains * 4 + (tot_chains + 1) * 8 + tot_hits * sizeof(pgr_hit_pair) + 256));
            uint8_t *ob = asm_out.as<uint8_t>();
            AssembleParams ap;
            ap.hits = idx->hitsB.as<HitRec>(); ap.seg_off = idx->seg_off.as<uint64_t>(); ap.seg_keys = idx->seg_keys.as<SortKey>(); ap.n_seg = n_seg;
            ap.out_idx = cp.out_idx; ap.out_start = cp.out_start; ap.out_score = cp.out_score; ap.seg_n_out = cp.seg_n_out;
            ap.hit_prefix = hit_prefix; ap.chain_prefix = chain_prefix; ap.target_prefix = target_prefix;
            ap.target_chain_off = (uint64_t *)ob; ob += (tot_targets + 1) * 8;
            ap.chain_hit_off = (uint64_t *)ob; ob += (tot_chains + 1) * 8;
            ap.hits_out = (pgr_hit_pair *)ob; ob += ((tot_hits * sizeof(pgr_hit_pair) + 7) & ~(size_t)7);
            ap.target_sid = (uint32_t *)ob; ob += tot_targets * 4;
            ap.target_qid = (uint32_t *)ob; ob += tot_targets * 4;
            ap.chain_score = (float *)ob;
            ap.chain_base = chain_base; ap.hit_base = hit_base;
            assemble_kernel<<<(uint32_t)ceil_div<uint64_t>(n_seg, 128), 128, 0, st>>>(ap);
            set_u64_kernel<<<1, 1, 0, st>>>(ap.target_chain_off + tot_targets, chain_base + tot_chains);
            set_u64_kernel<<<1, 1, 0, st>>>(ap.chain_hit_off + tot_chains, hit_base + tot_hits);
            idx->launches += 4;
            PGR_CUDA(cudaGetLastError());
            trace_mark("query_batch: device assembly");
            out->n_targets = tot_targets; out->n_chains = tot_chains; out->n_hits = tot_hits;
            out->target_sid = ap.target_sid; out->target_qid = ap.target_qid; out->target_chain_off = ap.target_chain_off;
            out->chain_score = ap.chain_score; out->chain_hit_off = ap.chain_hit_off; out->hits = ap.hits_out;
        }
    }
    return PGR_OK;
}

// grow-only host array for the batch result (page-locked pool buffers): append(n) returns where the next n items go
namespace {
template <class T>
struct HostArr {
    T *p = nullptr; size_t n = 0, cap = 0;
    // false when out of memory.  `sync` is called before the contents move to a larger buffer (copies may be in flight).
    template <class F> bool reserve(size_t need, F sync, size_t hint = 0) {
        if (need <= cap) return true;
        const size_t ncap = std::max(std::max(need, hint), cap + cap / 2);   // the hint only sizes an allocation that has to happen anyway
        T *np = (T *)result_alloc(std::max<size_t>(1, ncap) * sizeof(T));
        if (!np) return false;
        if (p) { sync(); memcpy(np, p, n * sizeof(T)); result_free(p); }
        p = np; cap = ncap;
        return true;
    }
};
}  // namespace

extern "C" {

// replaces SeqIndexDB::query_fragment_to_hps (ext.rs:252-282) for a batch of queries.  A large batch is cut into groups of
// queries: the device-to-host copy of a group's result (the hit pairs are the bulk: 20 B each) runs on its own stream while
// the next group is computed, into one set of host arrays whose offsets are global from the start.
int pgr_b200_query_batch(pgr_b200_index *idx, size_t n_q, const uint8_t *const *seqs, const size_t *lens, const pgr_query_params *prm,
                         pgr_query_result **out) {
    if (!idx || !prm || !out || (n_q && (!seqs || !lens))) { set_error("NULL argument"); return PGR_E_ARG; }
    PGR_CUDA(cudaSetDevice(idx->ctx->device));
    pgr_b200_ctx *ctx = idx->ctx;
    cudaStream_t st = ctx->stream;
    trace_mark("query_batch: begin");
    PGR_TRY(ensure_sid_count(idx));
    trace_mark("query_batch: sid_count");
    uint64_t total_len = 0;
    for (size_t q = 0; q < n_q; q++) total_len += lens[q];
    static const int forced_groups = getenv("PGR_B200_QUERY_GROUPS") ? atoi(getenv("PGR_B200_QUERY_GROUPS")) : 0;   // A/B aid
    // Default: one group.  Measured on config 4 (10 000 x 20 kb queries, 573 MB of result; profiles/r2_sort_query_ab.txt): the copy
    // of a group's result does hide behind the next group's kernels, but four groups of 2 500 queries cost more in per-group work
    // (upload not overlapped inside a group, shorter chain launches) than the 8.6 ms they hide: 34.4 ms with one group, ~36 ms with four.
    size_t n_groups = forced_groups > 0 ? (size_t)forced_groups : 1;
    n_groups = std::max<size_t>(1, std::min(n_groups, std::max<size_t>(1, n_q)));
    std::vector<size_t> cut(n_groups + 1, n_q);
    cut[0] = 0;
    {
        uint64_t acc = 0;
        size_t g = 1;
        for (size_t q = 0; q < n_q && g < n_groups; q++) {
            acc += lens[q];
            while (g < n_groups && acc * n_groups >= total_len * g) cut[g++] = q + 1;
        }
    }
    if (!idx->d2h_stream) PGR_CUDA(cudaStreamCreateWithFlags(&idx->d2h_stream, cudaStreamNonBlocking));
    cudaStream_t cs = idx->d2h_stream;
    cudaEvent_t ev_ready[2], ev_copied[2];
    for (int i = 0; i < 2; i++) { cudaEventCreateWithFlags(&ev_ready[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&ev_copied[i], cudaEventDisableTiming); }
    bool copied_pending[2] = {false, false};
    HostArr<uint32_t> h_tsid, h_tqid;
    HostArr<uint64_t> h_tco, h_cho;
    HostArr<float> h_score;
    HostArr<pgr_hit_pair> h_hits;
    std::vector<std::pair<size_t, size_t>> group_targets;   // (first target, count) per group, for q_target_off
    auto sync_copies = [&]() { cudaStreamSynchronize(cs); };
    int rc = PGR_OK;
    auto cleanup = [&]() {
        cudaStreamSynchronize(cs);
        for (int i = 0; i < 2; i++) { cudaEventDestroy(ev_ready[i]); cudaEventDestroy(ev_copied[i]); }
    };
    auto fail = [&](int r) { cleanup(); result_free(h_tsid.p); result_free(h_tqid.p); result_free(h_tco.p); result_free(h_cho.p); result_free(h_score.p); result_free(h_hits.p); return r; };
    for (size_t g = 0; g < n_groups; g++) {
        const int slot = (int)(g & 1);
        DevBuf &asm_out = slot ? idx->asm_out2 : idx->asm_out;
        if (copied_pending[slot]) { cudaEventSynchronize(ev_copied[slot]); copied_pending[slot] = false; }   // the copy out of this buffer is over
        QueryGroupOut go;
        const size_t q0 = cut[g], nq = cut[g + 1] - cut[g];
        if (nq == 0) { group_targets.push_back({h_tsid.n, 0}); continue; }
        rc = query_group(idx, nq, seqs + q0, lens + q0, prm, asm_out, h_score.n, h_hits.n, &go);
        if (rc != PGR_OK) return fail(rc);
        group_targets.push_back({h_tsid.n, (size_t)go.n_targets});
        if (go.n_targets == 0 && go.n_chains == 0 && go.n_hits == 0) continue;
        // capacity: after the first group with results, extrapolate to the whole batch (+30 %)
        const double scale = (double)total_len / (double)std::max<uint64_t>(1, [&] { uint64_t a = 0; for (size_t q = 0; q < cut[g + 1]; q++) a += lens[q]; return a; }()) * 1.3;
        const bool ok = h_tsid.reserve(h_tsid.n + go.n_targets, sync_copies, (size_t)((h_tsid.n + go.n_targets) * scale)) &&
                        h_tqid.reserve(h_tqid.n + go.n_targets, sync_copies, (size_t)((h_tqid.n + go.n_targets) * scale)) &&
                        h_tco.reserve(h_tco.n + go.n_targets + 1, sync_copies, (size_t)((h_tco.n + go.n_targets) * scale) + 1) &&
                        h_score.reserve(h_score.n + go.n_chains, sync_copies, (size_t)((h_score.n + go.n_chains) * scale)) &&
                        h_cho.reserve(h_cho.n + go.n_chains + 1, sync_copies, (size_t)((h_cho.n + go.n_chains) * scale) + 1) &&
                        h_hits.reserve(h_hits.n + go.n_hits, sync_copies, (size_t)((h_hits.n + go.n_hits) * scale));
        if (!ok) { set_error("out of host memory"); return fail(PGR_E_ARG); }
        cudaEventRecord(ev_ready[slot], st);
        cudaStreamWaitEvent(cs, ev_ready[slot], 0);
        cudaError_t ce = cudaSuccess;
        if (go.n_targets) {
            ce = cudaMemcpyAsync(h_tsid.p + h_tsid.n, go.target_sid, go.n_targets * 4, cudaMemcpyDeviceToHost, cs);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_tqid.p + h_tqid.n, go.target_qid, go.n_targets * 4, cudaMemcpyDeviceToHost, cs);
        }
        // the offset arrays carry one closing entry per group; the next group overwrites it with its own first entry (equal value)
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_tco.p + h_tco.n, go.target_chain_off, (go.n_targets + 1) * 8, cudaMemcpyDeviceToHost, cs);
        if (ce == cudaSuccess && go.n_chains) ce = cudaMemcpyAsync(h_score.p + h_score.n, go.chain_score, go.n_chains * 4, cudaMemcpyDeviceToHost, cs);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_cho.p + h_cho.n, go.chain_hit_off, (go.n_chains + 1) * 8, cudaMemcpyDeviceToHost, cs);
        if (ce == cudaSuccess && go.n_hits) ce = cudaMemcpyAsync(h_hits.p + h_hits.n, go.hits, go.n_hits * sizeof(pgr_hit_pair), cudaMemcpyDeviceToHost, cs);
        if (ce != cudaSuccess) { set_error("D2H of the query result failed: %s", cudaGetErrorString(ce)); return fail(PGR_E_CUDA); }
        cudaEventRecord(ev_copied[slot], cs);
        copied_pending[slot] = true;
        h_tsid.n += go.n_targets; h_tqid.n += go.n_targets; h_tco.n += go.n_targets;
        h_score.n += go.n_chains; h_cho.n += go.n_chains; h_hits.n += go.n_hits;
    }
    cleanup();
    trace_mark("query_batch: host assembly");
    const size_t n_targets = h_tsid.n, n_chains = h_score.n, n_hits_out = h_hits.n;
    // closing entries (also for an empty result)
    if (!h_tco.reserve(n_targets + 1, sync_copies) || !h_cho.reserve(n_chains + 1, sync_copies) || !h_tsid.reserve(1, sync_copies) ||
        !h_score.reserve(1, sync_copies) || !h_hits.reserve(1, sync_copies)) { set_error("out of host memory"); return fail(PGR_E_ARG); }
    h_tco.p[n_targets] = n_chains; h_cho.p[n_chains] = n_hits_out;
    // targets are sorted by query inside a group, groups are in query order: q_target_off[q] = number of targets of queries < q
    std::vector<uint64_t> q_target_off(n_q + 1, 0);
    for (size_t g = 0; g < n_groups; g++) {
        const size_t q0 = cut[g], nq = cut[g + 1] - cut[g];
        size_t ti = group_targets[g].first;
        const size_t te = ti + group_targets[g].second;
        for (size_t q = 0; q < nq; q++) {
            q_target_off[q0 + q] = ti;
            while (ti < te && h_tqid.p[ti] == q) ti++;
        }
    }
    q_target_off[n_q] = n_targets;
    result_free(h_tqid.p);
    pgr_query_result *r = (pgr_query_result *)calloc(1, sizeof(pgr_query_result));
    r->n_queries = n_q; r->n_targets = n_targets; r->n_chains = n_chains; r->n_hits = n_hits_out;
    r->q_target_off = (uint64_t *)result_alloc((n_q + 1) * 8);
    memcpy(r->q_target_off, q_target_off.data(), (n_q + 1) * 8);
    r->target_sid = h_tsid.p; r->target_chain_off = h_tco.p; r->chain_score = h_score.p;
    r->chain_hit_off = h_cho.p; r->hits = h_hits.p;
    *out = r;
    return PGR_OK;
}

// replaces aln::sparse_aln(&mut sp_hits, max_span, penalty, max_gap, oriented) on one hit list.  hits is sorted in place
// (stable, by q_bgn, aln.rs:21).
int pgr_b200_sparse_aln(pgr_hit_pair *hits, size_t n, uint32_t max_span, float penalty, int64_t max_gap, int oriented, size_t *n_chains,
                        uint64_t **chain_off, float **scores, pgr_hit_pair **chain_hits) {
    if (!hits || !n_chains || !chain_off || !scores || !chain_hits) { set_error("NULL argument"); return PGR_E_ARG; }
    if (n < 2) { set_error("assert!(sp_hits.len() > 1) violated (aln.rs:24)"); return PGR_E_ASSERT; }
    pgr_b200_ctx *ctx = tls_ctx();
    if (!ctx) return PGR_E_NO_DEVICE;
    PGR_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    std::stable_sort(hits, hits + n, [](const pgr_hit_pair &a, const pgr_hit_pair &b) { return a.qb < b.qb; });
    std::vector<HitRec> hr(n);
    for (size_t i = 0; i < n; i++) {
        memset(&hr[i], 0, sizeof(HitRec));
        hr[i].qb = hits[i].qb; hr[i].qe = hits[i].qe; hr[i].qo = hits[i].qo; hr[i].tb = hits[i].tb; hr[i].te = hits[i].te; hr[i].to = hits[i].to;
    }
    DevBuf d_hits, d_f, d_u, d_b, d_seg;
    int rc = PGR_OK;
    auto cleanup = [&]() { d_hits.release(); d_f.release(); d_u.release(); d_b.release(); d_seg.release(); };
    if ((rc = d_hits.ensure(n * sizeof(HitRec))) || (rc = d_f.ensure(n * sizeof(float) * 2)) || (rc = d_u.ensure(n * sizeof(uint32_t) * 5)) ||
        (rc = d_b.ensure(n * 2)) || (rc = d_seg.ensure(64))) { cleanup(); return rc; }
    const uint64_t seg_off_h[2] = {0, n};
    uint64_t *d_off = d_seg.as<uint64_t>();
    uint32_t *d_meta = (uint32_t *)(d_off + 2);
    cudaMemcpyAsync(d_hits.p, hr.data(), n * sizeof(HitRec), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_off, seg_off_h, sizeof seg_off_h, cudaMemcpyHostToDevice, st);
    ChainParams cp;
    cp.hits = d_hits.as<HitRec>(); cp.seg_off = d_off; cp.n_seg = 1; cp.max_span = max_span; cp.penalty = penalty;
    cp.has_gap = max_gap >= 0 ? 1u : 0u; cp.max_gap = max_gap >= 0 ? (float)(uint32_t)max_gap : 0.0f; cp.oriented = oriented ? 1u : 0u;
    cp.v_s = d_f.as<float>(); cp.out_score = cp.v_s + n;
    uint32_t *u = d_u.as<uint32_t>();
    cp.best_pre = (int32_t *)u; cp.cls_first = u + n; cp.cls_last = u + 2 * n; cp.order = u + 3 * n; cp.out_idx = u + 4 * n;
    cp.visited = d_b.as<uint8_t>(); cp.out_start = cp.visited + n;
    cp.seg_n_out = d_meta; cp.seg_n_chains = d_meta + 1; cp.seg_err = d_meta + 2;
    chain_kernel<<<1, 64, 0, st>>>(cp);
    std::vector<uint32_t> out_idx(n);
    std::vector<uint8_t> out_start(n);
    std::vector<float> out_score(n);
    uint32_t meta[3] = {0, 0, 0};
    cudaMemcpyAsync(out_idx.data(), cp.out_idx, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(out_start.data(), cp.out_start, n, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(out_score.data(), cp.out_score, n * sizeof(float), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(meta, d_meta, sizeof meta, cudaMemcpyDeviceToHost, st);
    const cudaError_t e = cudaStreamSynchronize(st);
    cleanup();
    if (e != cudaSuccess) { set_error("sparse_aln kernel failed: %s", cudaGetErrorString(e)); return PGR_E_CUDA; }
    if (meta[2]) { set_error("sparse_aln: all scores <= 0 (the reference would not terminate)"); return PGR_E_ASSERT; }
    std::vector<uint64_t> off;
    std::vector<float> sc;
    std::vector<pgr_hit_pair> ch;
    for (uint32_t a = 0; a < meta[0]; a++) {
        if (out_start[a]) { off.push_back(ch.size()); sc.push_back(out_score[a]); }
        ch.push_back(hits[out_idx[a]]);
    }
    off.push_back(ch.size());
    *n_chains = sc.size();
    *chain_off = host_dup(off);
    *scores = host_dup(sc);
    *chain_hits = host_dup(ch);
    return PGR_OK;
}

// replaces seq_db::frag_map_to_adj_list(&frag_map, min_count, keeps) -> AdjList
int pgr_b200_adj_list(pgr_b200_index *idx, size_t min_count, const uint32_t *keeps, size_t n_keeps, int has_keeps, pgr_adj_pair **out,
                      size_t *n_out) {
    if (!idx || !out || !n_out) { set_error("NULL argument"); return PGR_E_ARG; }
    PGR_CUDA(cudaSetDevice(idx->ctx->device));
    PGR_TRY(pgr_b200_index_finalize(idx));
    cudaStream_t st = idx->ctx->stream;
    const uint64_t n = idx->n_tuples;
    *out = (pgr_adj_pair *)result_alloc(sizeof(pgr_adj_pair));
    *n_out = 0;
    if (n < 2) return PGR_OK;  // seq_db.rs:889-891
    trace_mark("adj_list: begin (after finalize)");
    std::vector<uint32_t> ks(keeps, keeps + (has_keeps ? n_keeps : 0));
    std::sort(ks.begin(), ks.end());
    PGR_TRY(idx->scratch0.ensure(std::max<size_t>(1, ks.size()) * sizeof(uint32_t)));
    if (!ks.empty()) PGR_CUDA(cudaMemcpyAsync(idx->scratch0.p, ks.data(), ks.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PGR_TRY(idx->hitsA.ensure(n * sizeof(AdjRow)));
    PGR_TRY(idx->keysA.ensure(n * sizeof(SortKey)));
    PGR_TRY(idx->idxA.ensure(n * sizeof(uint32_t)));
    const uint32_t g = (uint32_t)ceil_div<uint64_t>(n, 256);
    AdjRow *rows = idx->hitsA.as<AdjRow>();
    adj_rows_kernel<<<g, 256, 0, st>>>(idx->ukeys.as<SortKey>(), idx->offsets.as<uint64_t>(), idx->sigs.as<pgr_frag_sig>(), n, idx->n_keys,
                                       (uint64_t)min_count, idx->scratch0.as<uint32_t>(), (uint32_t)ks.size(), has_keeps ? 1u : 0u, rows);
    iota_kernel<<<g, 256, 0, st>>>(idx->idxA.as<uint32_t>(), n);
    idx->launches += 2;
    for (int round = 0; round < 3; round++) {   // out.par_sort() on the full tuple (seq_db.rs:892)
        adj_keys_kernel<<<g, 256, 0, st>>>(rows, idx->idxA.as<uint32_t>(), n, round, idx->keysA.as<SortKey>());
        idx->launches += 1;
        PGR_TRY(index_sort(idx, n, 0, 13));
    }
    trace_mark("adj_list: rows + 3 sort rounds");
    PGR_TRY(idx->scratch1.ensure(n * sizeof(uint32_t)));
    PGR_TRY(idx->scratch2.ensure((n + 1) * sizeof(uint64_t)));
    adj_flag_kernel<<<g, 256, 0, st>>>(rows, idx->idxA.as<uint32_t>(), n, idx->scratch1.as<uint32_t>());
    uint64_t total = 0;
    PGR_TRY(scan_u32(idx, idx->scratch1.as<uint32_t>(), n, idx->scratch2.as<uint64_t>(), &total));
    idx->launches += 1;
    if (total == 0) return PGR_OK;
    PGR_TRY(idx->scratch3.ensure(total * sizeof(pgr_adj_pair)));
    adj_emit_kernel<<<g, 256, 0, st>>>(rows, idx->idxA.as<uint32_t>(), n, idx->scratch1.as<uint32_t>(), idx->scratch2.as<uint64_t>(),
                                       idx->scratch3.as<pgr_adj_pair>());
    idx->launches += 1;
    PGR_CUDA(cudaGetLastError());
    result_free(*out);
    *out = (pgr_adj_pair *)result_alloc(total * sizeof(pgr_adj_pair));
    if (!*out) { set_error("out of host memory"); return PGR_E_ARG; }
    PGR_CUDA(cudaMemcpyAsync(*out, idx->scratch3.p, total * sizeof(pgr_adj_pair), cudaMemcpyDeviceToHost, st));
    PGR_CUDA(cudaStreamSynchronize(st));
    *n_out = total;
    trace_mark("adj_list: join + D2H");
    return PGR_OK;
}

// replaces seq_db::generate_smp_adj_list_for_seq(&seq, sid, &frag_map, &spec, min_count) -> AdjList for a batch of sequences
int pgr_b200_smp_adj_list_for_seqs(pgr_b200_index *idx, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens,
                                   const uint64_t *min_counts, pgr_adj_pair **out, size_t *n_out) {
    if (!idx || !out || !n_out || (n && (!sids || !seqs || !lens || !min_counts))) { set_error("NULL argument"); return PGR_E_ARG; }
    PGR_CUDA(cudaSetDevice(idx->ctx->device));
    cudaStream_t st = idx->ctx->stream;
    *out = (pgr_adj_pair *)result_alloc(sizeof(pgr_adj_pair));
    *n_out = 0;
    uint64_t n_qp = 0;
    std::vector<uint64_t> qp_off;
    PGR_TRY(query_pairs_and_lookup(idx, n, seqs, lens, &n_qp, &qp_off));   // pairs with the strict '<' rule (seq_db.rs:962-966), key counts
    if (n_qp < 2) return PGR_OK;
    PGR_TRY(idx->scratch0.ensure((n + 1) * sizeof(uint64_t) + n * sizeof(uint64_t) + n * sizeof(uint32_t)));
    uint64_t *d_qoff = idx->scratch0.as<uint64_t>(), *d_mc = d_qoff + (n + 1);
    uint32_t *d_sids = (uint32_t *)(d_mc + n);
    PGR_CUDA(cudaMemcpyAsync(d_qoff, qp_off.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    PGR_CUDA(cudaMemcpyAsync(d_mc, min_counts, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    PGR_CUDA(cudaMemcpyAsync(d_sids, sids, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PGR_TRY(idx->scratch1.ensure(n_qp * sizeof(uint32_t)));
    PGR_TRY(idx->scratch2.ensure((n_qp + 1) * sizeof(uint64_t)));
    const uint32_t g = (uint32_t)ceil_div<uint64_t>(n_qp, 256);
    smp_adj_flag_kernel<<<g, 256, 0, st>>>(idx->qtuples.as<FragTuple>(), n_qp, d_qoff, idx->q_hit_count.as<uint32_t>(), d_mc, idx->scratch1.as<uint32_t>());
    uint64_t total = 0;
    PGR_TRY(scan_u32(idx, idx->scratch1.as<uint32_t>(), n_qp, idx->scratch2.as<uint64_t>(), &total));
    idx->launches += 1;
    if (total == 0) return PGR_OK;
    PGR_TRY(idx->scratch3.ensure(2 * total * sizeof(pgr_adj_pair)));
    smp_adj_emit_kernel<<<g, 256, 0, st>>>(idx->qtuples.as<FragTuple>(), n_qp, idx->scratch1.as<uint32_t>(), idx->scratch2.as<uint64_t>(), d_sids,
                                           idx->scratch3.as<pgr_adj_pair>());
    idx->launches += 1;
    PGR_CUDA(cudaGetLastError());
    result_free(*out);
    *out = (pgr_adj_pair *)result_alloc(2 * total * sizeof(pgr_adj_pair));
    if (!*out) { set_error("out of host memory"); return PGR_E_ARG; }
    PGR_CUDA(cudaMemcpyAsync(*out, idx->scratch3.p, 2 * total * sizeof(pgr_adj_pair), cudaMemcpyDeviceToHost, st));
    PGR_CUDA(cudaStreamSynchronize(st));
    *n_out = 2 * total;
    return PGR_OK;
}

}  // extern "C"
