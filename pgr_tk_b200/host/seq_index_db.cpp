#include "seq_index_db.hpp"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace pgrb200 {

// fasta_io.rs:46-172 (the parsing itself: fastx_ingest.cpp)
bool read_fastx(const std::string &path, std::vector<SeqRec> &out, std::string &err) {
    ParsedFile pf;
    parse_fastx_file(path, INGEST_PLAIN, nullptr, pf);
    if (!pf.ok) { err = pf.err; return false; }
    for (size_t i = 0; i < pf.seqs.size(); i++) {
        SeqRec r;
        r.id = pf.ids[i];
        r.seq.assign(pf.seqs[i].p, pf.seqs[i].p + pf.seqs[i].len);
        out.push_back(std::move(r));
    }
    return true;
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

SeqIndexDB::~SeqIndexDB() { reset(); }

void SeqIndexDB::reset() {
    if (idx_) { pgr_b200_index_free(idx_); idx_ = nullptr; }
    if (midx_) { pgr_b200_mindex_free(midx_); midx_ = nullptr; }
    if (map_) { pgr_b200_mdb_map_close(map_); map_ = nullptr; }
    seqs_.clear();
    seq_data_.clear();
    file_bufs_.clear();
    owned_seqs_.clear();
    fastx_backend_ = false;
}

int SeqIndexDB::load_from_fastx(const std::string &path, uint32_t w, uint32_t k, uint32_t r, uint32_t min_span) {
    spec_.w = w; spec_.k = k; spec_.r = r; spec_.min_span = min_span; spec_.sketch = 0;
    reset();
    idx_ = pgr_b200_index_new(&spec_, 0 /* FASTX fragment numbering */, -1);
    if (!idx_) { err_ = pgr_b200_last_error(); return PGR_E_NO_DEVICE; }
    fastx_backend_ = true;
    return load_seqs_from_fastx(path);
}

int SeqIndexDB::append_from_fastx(const std::string &path) {
    if (!idx_ || !fastx_backend_) { err_ = "Only DB created with load_from_fastx() can add data from another fastx file"; return PGR_E_ARG; }
    return load_seqs_from_fastx(path);
}

// seq_db.rs:471-525: sid continues from seqs.len(); records are handed to the GPU in the reference's batches of <=129
int SeqIndexDB::load_seqs_from_fastx(const std::string &path) {
    std::vector<std::unique_ptr<ParsedFile>> one;
    one.emplace_back(new ParsedFile());
    parse_fastx_file(path, INGEST_PLAIN, nullptr, *one[0]);
    if (!one[0]->ok) { err_ = one[0]->err; return PGR_E_IO; }
    return add_parsed(one);
}

// the sequences of the parsed files (in file order) become the next sequence ids; ONE GPU call for all of them: the batch
// boundary (129 records in the reference) only sets its parallel granularity, fragment ids are a running counter across batches
int SeqIndexDB::add_parsed(std::vector<std::unique_ptr<ParsedFile>> &files) {
    uint32_t sid = (uint32_t)seqs_.size();
    std::vector<uint32_t> sids;
    std::vector<const uint8_t *> ptrs;
    std::vector<size_t> lens;
    for (auto &f : files) {
        for (size_t i = 0; i < f->seqs.size(); i++) {
            sids.push_back(sid);
            ptrs.push_back(f->gpu_seqs.empty() ? f->seqs[i].p : f->gpu_seqs[i].p);   // the page-locked copy when there is one
            lens.push_back(f->seqs[i].len);
            CompactSeq cs;
            cs.id = sid; cs.len = f->seqs[i].len; cs.name = f->ids[i]; cs.source = f->path;
            seqs_.push_back(std::move(cs));
            if (keep_seqs_) seq_data_.push_back(f->seqs[i]);
            sid++;
        }
        timing_.bases += f->bases; timing_.reader_read_s += f->read_s; timing_.reader_parse_s += f->parse_s; timing_.reader_pin_s += f->pin_s;
    }
    const double t0 = now_s();
    const int rc = midx_ ? pgr_b200_mindex_add_batch(midx_, sids.size(), sids.data(), ptrs.data(), lens.data())
                         : pgr_b200_index_add_batch(idx_, sids.size(), sids.data(), ptrs.data(), lens.data());
    timing_.gpu_index_s += now_s() - t0;
    if (rc != PGR_OK) err_ = pgr_b200_last_error();
    for (auto &f : files) { f->gpu_buf.reset(); if (keep_seqs_) file_bufs_.push_back(std::move(f->buf)); }
    return rc;
}

int SeqIndexDB::load_from_fastx_list(const std::vector<std::string> &paths, uint32_t w, uint32_t k, uint32_t r, uint32_t min_span, int n_readers, int n_gpus,
                                     const std::vector<int> &devices) {
    spec_.w = w; spec_.k = k; spec_.r = r; spec_.min_span = min_span; spec_.sketch = 0;
    reset();
    timing_ = Timing();
    if (paths.empty()) { err_ = "empty file list"; return PGR_E_ARG; }
    if (!devices.empty()) n_gpus = (int)devices.size();
    fastx_backend_ = true;
    // The library's packed transport reads the caller's bytes with its own host threads and moves pageable memory as fast as
    // page-locked memory, so the files are parsed in plain memory: reserving page-locked slots cost ~1 s per GB here, more
    // than the whole index build of config 3.  PGR_B200_INGEST_PINNED=1 restores the page-locked slots (A/B:
    // profiles/r2_cli_make_frgdb.txt).
    static const bool pinned_slots = getenv("PGR_B200_INGEST_PINNED") != nullptr;
    FastxPipeline pipe(paths, n_readers, !pinned_slots ? INGEST_PLAIN : keep_seqs_ ? INGEST_KEEP : INGEST_PINNED);
    // the readers are already at work on the first files while the CUDA context comes up (0.3 - 2 s depending on the box)
    const double t_init0 = now_s();
    if (n_gpus > 1) {
        midx_ = devices.empty() ? pgr_b200_mindex_new(&spec_, 0, n_gpus) : pgr_b200_mindex_new_devices(&spec_, 0, n_gpus, devices.data());
        if (!midx_) { err_ = pgr_b200_last_error(); return PGR_E_NO_DEVICE; }
    } else {
        idx_ = pgr_b200_index_new(&spec_, 0, -1);
        if (!idx_) { err_ = pgr_b200_last_error(); return PGR_E_NO_DEVICE; }
    }
    timing_.device_init_s = now_s() - t_init0;
    // a batch = the files that are parsed by now (at least one); with several GPUs at least a few files per call so that every
    // GPU has a block of the batch to work on
    const size_t max_batch = (size_t)std::max(4, 2 * std::max(1, n_gpus));   // page-locked slots are scarce: window = readers + 4
    size_t i = 0;
    while (i < paths.size()) {
        std::vector<std::unique_ptr<ParsedFile>> batch;
        const double t0 = now_s();
        do {
            batch.push_back(pipe.take(i));
            i++;
            if (!batch.back()->ok) { err_ = batch.back()->err; return PGR_E_IO; }
        } while (i < paths.size() && batch.size() < max_batch && (pipe.ready(i) || (n_gpus > 1 && batch.size() < (size_t)n_gpus)));
        timing_.wait_parse_s += now_s() - t0;
        const int rc = add_parsed(batch);
        if (rc != PGR_OK) return rc;
    }
    if (midx_) {
        const double t0 = now_s();
        if (pgr_b200_mindex_finalize(midx_) != PGR_OK) { err_ = pgr_b200_last_error(); return PGR_E_CUDA; }
        idx_ = pgr_b200_mindex_gather(midx_, 0);
        if (!idx_) { err_ = pgr_b200_last_error(); return PGR_E_CUDA; }
        pgr_b200_mindex_free(midx_);
        midx_ = nullptr;
        timing_.merge_s += now_s() - t0;
    } else {
        const double t0 = now_s();
        if (pgr_b200_index_finalize(idx_) != PGR_OK) { err_ = pgr_b200_last_error(); return PGR_E_CUDA; }
        timing_.merge_s += now_s() - t0;
    }
    return PGR_OK;
}

int SeqIndexDB::load_from_seq_list(const std::vector<SeqRec> &seq_list, const std::string &source, uint32_t w, uint32_t k, uint32_t r, uint32_t min_span) {
    spec_.w = w; spec_.k = k; spec_.r = r; spec_.min_span = min_span; spec_.sketch = 0;
    reset();
    idx_ = pgr_b200_index_new(&spec_, 0, -1);
    if (!idx_) { err_ = pgr_b200_last_error(); return PGR_E_NO_DEVICE; }
    fastx_backend_ = true;
    std::vector<SeqRec> recs = seq_list;
    return add_records(recs, source);
}

int SeqIndexDB::add_records(std::vector<SeqRec> &recs, const std::string &source) {
    uint32_t sid = (uint32_t)seqs_.size();
    std::vector<uint32_t> sids;
    std::vector<const uint8_t *> ptrs;
    std::vector<size_t> lens;
    for (auto &r : recs) {
        sids.push_back(sid);
        ptrs.push_back(r.seq.data());
        lens.push_back(r.seq.size());
        CompactSeq cs;
        cs.id = sid; cs.len = r.seq.size(); cs.name = r.id; cs.source = source;
        seqs_.push_back(std::move(cs));
        sid++;
    }
    // one GPU call per file: the batch boundary (129 records) only sets the reference's parallel granularity, it has
    // no effect on results (fragment ids are a running counter across batches)
    const int rc = pgr_b200_index_add_batch(idx_, recs.size(), sids.data(), ptrs.data(), lens.data());
    if (rc != PGR_OK) err_ = pgr_b200_last_error();
    if (keep_seqs_) for (auto &r : recs) {
        owned_seqs_.push_back(std::move(r.seq));
        SeqSpan sp; sp.p = owned_seqs_.back().data(); sp.len = owned_seqs_.back().size();
        seq_data_.push_back(sp);
    }
    return rc;
}

int SeqIndexDB::load_from_index_files(const std::string &prefix, bool mdb_resident) {
    reset();
    if (mdb_resident) {
        map_ = pgr_b200_mdb_map_open((prefix + ".mdb").c_str());
        if (!map_) { err_ = pgr_b200_last_error(); return PGR_E_IO; }
        pgr_b200_mdb_map_info(map_, &spec_, nullptr, nullptr);
    } else {
        idx_ = pgr_b200_index_read_mdb((prefix + ".mdb").c_str(), -1);
        if (!idx_) { err_ = pgr_b200_last_error(); return PGR_E_IO; }
        pgr_b200_index_get_spec(idx_, &spec_);
    }
    FILE *f = fopen((prefix + ".midx").c_str(), "rb");
    if (!f) { err_ = "cannot open " + prefix + ".midx"; return PGR_E_IO; }
    char *line = nullptr;
    size_t cap = 0;
    ssize_t n;
    while ((n = getline(&line, &cap, f)) > 0) {     // sid \t len \t ctg_name \t source (seq_db.rs:795-807)
        std::string l(line, (size_t)n);
        while (!l.empty() && (l.back() == '\n' || l.back() == '\r')) l.pop_back();
        std::vector<std::string> fld;
        size_t a = 0;
        for (;;) { const size_t b = l.find('\t', a); fld.push_back(l.substr(a, b == std::string::npos ? b : b - a)); if (b == std::string::npos) break; a = b + 1; }
        if (fld.size() < 4) { err_ = "malformed .midx line"; free(line); fclose(f); return PGR_E_IO; }
        CompactSeq cs;
        cs.id = (uint32_t)strtoul(fld[0].c_str(), nullptr, 10); cs.len = strtoull(fld[1].c_str(), nullptr, 10); cs.name = fld[2]; cs.source = fld[3];
        seqs_.push_back(std::move(cs));
    }
    free(line);
    fclose(f);
    return PGR_OK;
}

bool SeqIndexDB::get_sub_seq_by_id(uint32_t sid, size_t bgn, size_t end, std::vector<uint8_t> &out) {
    if (sid < seq_data_.size()) {
        if (bgn > end || end > seq_data_[sid].len) return false;
        out.assign(seq_data_[sid].p + bgn, seq_data_[sid].p + end);
        return true;
    }
    if (!frag_store_.loaded()) return false;
    std::vector<uint8_t> whole;                       // frag_file_io.rs:182-229 fetches chunk-wise; the result is seq[bgn..end)
    if (!frag_store_.get_seq_by_id(sid, whole, err_) || bgn > end || end > whole.size()) return false;
    out.assign(whole.begin() + (ptrdiff_t)bgn, whole.begin() + (ptrdiff_t)end);
    return true;
}

int SeqIndexDB::load_from_frg_index(const std::string &prefix) {
    const int rc = load_from_index_files(prefix);
    if (rc != PGR_OK) return rc;
    if (!frag_store_.load(prefix, spec_.k, err_)) return PGR_E_IO;
    return PGR_OK;
}

// ---- FragStore -----------------------------------------------------------------------------------------------------------
namespace {
struct BinReader {
    const uint8_t *p, *e;
    bool ok = true;
    uint8_t u8() { if (p >= e) { ok = false; return 0; } return *p++; }
    uint64_t varint() {
        const uint8_t t = u8();
        if (t < 251) return t;
        const int n = t == 251 ? 2 : t == 252 ? 4 : t == 253 ? 8 : 16;
        if (e - p < n) { ok = false; return 0; }
        uint64_t v = 0;
        for (int i = 0; i < n && i < 8; i++) v |= (uint64_t)p[i] << (8 * i);
        p += n;
        return v;
    }
    // an element count: every element takes at least one byte, so a count above the remaining bytes is a corrupt file (checked
    // BEFORE anything is allocated from it)
    uint64_t count() { const uint64_t n = varint(); if (!ok || n > (uint64_t)(e - p)) { ok = false; return 0; } return n; }
    bool bytes(std::vector<uint8_t> &o) { const uint64_t n = varint(); if (!ok || (uint64_t)(e - p) < n) { ok = false; return false; } o.assign(p, p + n); p += n; return true; }
    bool str(std::string &o) { const uint64_t n = varint(); if (!ok || (uint64_t)(e - p) < n) { ok = false; return false; } o.assign((const char *)p, n); p += n; return true; }
};
bool slurp_plain(const std::string &path, std::vector<uint8_t> &buf) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    uint8_t tmp[1 << 16];
    size_t got;
    while ((got = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    fclose(f);
    return true;
}
bool inflate_raw(const uint8_t *in, size_t n, std::vector<uint8_t> &out) {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<Bytef *>(in); zs.avail_in = (uInt)n;
    out.clear();
    uint8_t tmp[1 << 16];
    int rc;
    do {
        zs.next_out = tmp; zs.avail_out = sizeof tmp;
        rc = inflate(&zs, Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END) { inflateEnd(&zs); return false; }
        out.insert(out.end(), tmp, tmp + (sizeof tmp - zs.avail_out));
    } while (rc != Z_STREAM_END);
    inflateEnd(&zs);
    return true;
}
uint8_t rc_base(uint8_t b) {   // fasta_io.rs:26-44
    switch (b) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
        default: return b;
    }
}
}  // namespace

bool FragStore::load(const std::string &prefix, uint32_t k, std::string &err) {
    k_ = k;
    std::vector<uint8_t> sd;
    if (!slurp_plain(prefix + ".sdx", sd) || sd.size() < 7 || memcmp(sd.data(), "SDX:0.5", 7) != 0) { err = "sdx file open / version error"; return false; }
    frg_.clear();
    if (!slurp_plain(prefix + ".frg", frg_) || frg_.size() < 7 || memcmp(frg_.data(), "FRG:0.5", 7) != 0) { err = "frg file open / version error"; return false; }
    BinReader r{sd.data() + 7, sd.data() + sd.size()};
    chunk_size_ = r.varint();
    const uint64_t na = r.count();
    addr_.clear();
    for (uint64_t i = 0; i < na && r.ok; i++) { Addr a; a.off = r.varint(); a.len = r.varint(); a.bases = r.varint(); addr_.push_back(a); }
    const uint64_t nsq = r.count();
    seqs_.clear();
    for (uint64_t i = 0; i < nsq && r.ok; i++) {
        SeqEntry s;
        s.has_source = r.u8() != 0;
        if (s.has_source) r.str(s.source);
        r.str(s.name);
        s.id = (uint32_t)r.varint(); s.frag_first = (uint32_t)r.varint(); s.frag_count = (uint32_t)r.varint(); s.len = r.varint();
        seqs_.push_back(std::move(s));
    }
    if (!r.ok || chunk_size_ == 0) { err = "read sdx file error"; return false; }
    cache_.clear();
    return true;
}

const FragStore::Fragment *FragStore::fragment(uint32_t id, std::string &err) {
    const uint64_t c = id / chunk_size_;
    if (c >= addr_.size()) { err = "fragment id out of range"; return nullptr; }
    for (auto &e : cache_) if (e.first == c) { const size_t i = id % chunk_size_; return i < e.second.size() ? &e.second[i] : nullptr; }
    const Addr &a = addr_[c];
    if (7 + a.off + a.len > frg_.size()) { err = "frg chunk out of range"; return nullptr; }
    std::vector<uint8_t> raw;
    if (!inflate_raw(frg_.data() + 7 + a.off, a.len, raw)) { err = "frg chunk inflate error"; return nullptr; }
    BinReader r{raw.data(), raw.data() + raw.size()};
    std::vector<Fragment> fr((size_t)r.count());
    for (auto &f : fr) {
        f.kind = (uint8_t)r.varint();
        if (f.kind == 0) {
            f.ref = (uint32_t)r.varint(); f.reversed = r.u8() != 0; f.len = (uint32_t)r.varint();
            f.segs.resize((size_t)r.count());
            for (auto &s : f.segs) {
                s.type = (uint32_t)r.varint(); s.a = s.b = 0;
                if (s.type == 1) { s.a = (uint32_t)r.varint(); s.b = (uint32_t)r.varint(); } else if (s.type == 2) s.a = r.u8();
            }
        } else {
            r.bytes(f.bases);
        }
        if (!r.ok) { err = "frg chunk decode error"; return nullptr; }
    }
    if (cache_.size() >= 8) cache_.erase(cache_.begin());
    cache_.emplace_back(c, std::move(fr));
    const size_t i = id % chunk_size_;
    return i < cache_.back().second.size() ? &cache_.back().second[i] : nullptr;
}

// seq_db.rs:685-735 reconstruct_seq_from_frags
bool FragStore::get_seq_by_id(uint32_t sid, std::vector<uint8_t> &out, std::string &err) {
    if (sid >= seqs_.size()) { err = "sequence id out of range"; return false; }
    out.clear();
    const SeqEntry &s = seqs_[sid];
    for (uint32_t id = s.frag_first; id < s.frag_first + s.frag_count; id++) {
        const Fragment *fp = fragment(id, err);
        if (!fp) return false;
        const Fragment f = *fp;                                   // copy: fetching the base may evict the chunk
        if (f.kind == 1 || f.kind == 3) out.insert(out.end(), f.bases.begin(), f.bases.end());
        else if (f.kind == 2) { if (f.bases.size() < k_) { err = "short internal fragment"; return false; } out.insert(out.end(), f.bases.begin() + k_, f.bases.end()); }
        else {
            const Fragment *bp = fragment(f.ref, err);
            if (!bp || bp->kind != 2) { err = "base fragment is not raw"; return false; }
            const std::vector<uint8_t> &base = bp->bases;
            std::vector<uint8_t> seq;
            for (const Seg &g : f.segs) {                         // seq_db.rs:158-174
                if (g.type == 0) seq.insert(seq.end(), base.begin(), base.end());
                else if (g.type == 1) { if (g.a > g.b || g.b > base.size()) { err = "bad match segment"; return false; } seq.insert(seq.end(), base.begin() + g.a, base.begin() + g.b); }
                else seq.push_back((uint8_t)g.a);
            }
            if (f.reversed) { std::reverse(seq.begin(), seq.end()); for (auto &b : seq) b = rc_base(b); }
            if (seq.size() < k_) { err = "short aligned fragment"; return false; }
            out.insert(out.end(), seq.begin() + k_, seq.end());
        }
    }
    return true;
}

int SeqIndexDB::query_fragment_to_hps(const std::vector<SeqRec> &queries, const pgr_query_params &params, pgr_query_result **out) {
    if (!idx_ && !map_) { err_ = "no index"; return PGR_E_ARG; }
    std::vector<const uint8_t *> ptrs;
    std::vector<size_t> lens;
    for (auto &q : queries) { ptrs.push_back(q.seq.data()); lens.push_back(q.seq.size()); }
    const int rc = map_ ? pgr_b200_query_batch_mmap(map_, -1, queries.size(), ptrs.data(), lens.data(), &params, out)
                        : pgr_b200_query_batch(idx_, queries.size(), ptrs.data(), lens.data(), &params, out);
    if (rc != PGR_OK) err_ = pgr_b200_last_error();
    return rc;
}

// ---- bincode 2 `config::standard()`: little endian, variable-length integers ----------------------------------------
static void put_varint(std::vector<uint8_t> &o, uint64_t v) {
    if (v < 251) { o.push_back((uint8_t)v); return; }
    int n; uint8_t tag;
    if (v < (1ull << 16)) { n = 2; tag = 251; } else if (v < (1ull << 32)) { n = 4; tag = 252; } else { n = 8; tag = 253; }
    o.push_back(tag);
    for (int i = 0; i < n; i++) o.push_back((uint8_t)(v >> (8 * i)));
}
static void put_bytes(std::vector<uint8_t> &o, const uint8_t *p, size_t n) { put_varint(o, n); o.insert(o.end(), p, p + n); }
static void put_string(std::vector<uint8_t> &o, const std::string &s) { put_bytes(o, (const uint8_t *)s.data(), s.size()); }

// raw deflate (the reference uses flate2's DeflateEncoder, Compression::default(); the streams differ byte-wise between
// deflate implementations, their inflated content is what is compared)
static bool deflate_raw(const std::vector<uint8_t> &in, std::vector<uint8_t> &out) {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out.resize(deflateBound(&zs, in.size()));
    zs.next_in = const_cast<Bytef *>(in.data()); zs.avail_in = (uInt)in.size();
    zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
    const int rc = deflate(&zs, Z_FINISH);
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return rc == Z_STREAM_END;
}

// seq_db.rs:814-873 on the fragment records of pgr_b200_index_compress_fragments (tests feed the oracle's records, which
// have the same layout, to run this writer without a device)
// chunks of `chunk_size` fragments are encoded (bincode 2, standard configuration) and deflated independently — by a pool of
// host threads, like the reference's par_iter over the chunks (seq_db.rs:829-853) — and written in chunk order
int write_frag_store(const std::string &prefix, size_t chunk_size, uint32_t k, const pgr_fragment *frags, size_t nf, const pgr_aln_seg *segs,
                     const std::vector<CompactSeq> &seqs, const std::vector<SeqSpan> &seq_data, std::string &err, int n_threads) {
    if (chunk_size == 0 || seq_data.size() != seqs.size()) { err = "write_frag_store: bad arguments"; return PGR_E_ARG; }
    std::vector<std::pair<uint32_t, uint32_t>> range(seqs.size(), {0, 0});   // per sequence (first fragment, count)
    for (size_t i = 0; i < nf; i++) {
        const pgr_fragment &f = frags[i];
        if (f.sid >= seqs.size() || f.end > seq_data[f.sid].len || f.bgn > f.end) { err = "fragment record out of range"; return PGR_E_ARG; }
        auto &rg = range[f.sid];
        if (rg.second == 0) rg.first = (uint32_t)i;
        rg.second++;
    }
    const size_t n_chunks = (nf + chunk_size - 1) / chunk_size;
    std::vector<std::vector<uint8_t>> z(n_chunks);
    std::vector<uint32_t> total_len(n_chunks, 0);
    std::atomic<size_t> next{0};
    std::atomic<bool> failed{false};
    auto work = [&]() {
        std::vector<uint8_t> w;
        for (;;) {
            const size_t c = next.fetch_add(1);
            if (c >= n_chunks || failed.load()) return;
            const size_t c0 = c * chunk_size, c1 = std::min(nf, c0 + chunk_size);
            w.clear();
            put_varint(w, c1 - c0);
            uint32_t tl = 0;
            for (size_t i = c0; i < c1; i++) {
                const pgr_fragment &f = frags[i];
                put_varint(w, f.kind);
                if (f.kind == 0) {                               // AlnSegments((ref, reversed, len, Vec<AlnSegment>))
                    put_varint(w, f.ref_frag); w.push_back(f.reversed ? 1 : 0); put_varint(w, f.len); put_varint(w, f.n_segs);
                    for (uint32_t s = 0; s < f.n_segs; s++) {
                        const pgr_aln_seg &g = segs[f.seg_off + s];
                        put_varint(w, g.type);
                        if (g.type == 1) { put_varint(w, g.a); put_varint(w, g.b); } else if (g.type == 2) w.push_back((uint8_t)g.a);
                    }
                    tl += f.len - k;
                } else {
                    put_bytes(w, seq_data[f.sid].p + f.bgn, f.end - f.bgn);
                    tl += f.kind == 2 ? f.len - k : f.len;
                }
            }
            total_len[c] = tl;
            if (!deflate_raw(w, z[c])) failed.store(true);
        }
    };
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = (int)std::min<size_t>((size_t)n_threads, std::max<size_t>(1, n_chunks));
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    if (failed.load()) { err = "deflate failed"; return PGR_E_IO; }
    FILE *frg = fopen((prefix + ".frg").c_str(), "wb"), *sdx = fopen((prefix + ".sdx").c_str(), "wb");
    if (!frg || !sdx) { err = "frag file creating fail"; if (frg) fclose(frg); if (sdx) fclose(sdx); return PGR_E_IO; }
    fwrite("FRG:0.5", 1, 7, frg);
    fwrite("SDX:0.5", 1, 7, sdx);
    std::vector<uint8_t> sd;
    put_varint(sd, chunk_size);
    put_varint(sd, n_chunks);
    uint64_t offset = 0;
    for (size_t c = 0; c < n_chunks; c++) {
        fwrite(z[c].data(), 1, z[c].size(), frg);
        put_varint(sd, offset); put_varint(sd, z[c].size()); put_varint(sd, total_len[c]);
        offset += z[c].size();
    }
    put_varint(sd, seqs.size());
    for (size_t i = 0; i < seqs.size(); i++) {               // CompactSeq {source: Option<String>, name, id, seq_frag_range, len}
        sd.push_back(1); put_string(sd, seqs[i].source);
        put_string(sd, seqs[i].name);
        put_varint(sd, seqs[i].id);
        put_varint(sd, range[i].first); put_varint(sd, range[i].second);
        put_varint(sd, seqs[i].len);
    }
    fwrite(sd.data(), 1, sd.size(), sdx);
    const bool bad = ferror(frg) || ferror(sdx);
    if (fclose(frg) != 0 || fclose(sdx) != 0 || bad) { err = "frag file writing error"; return PGR_E_IO; }
    return PGR_OK;
}

int SeqIndexDB::write_to_frag_files(const std::string &prefix, size_t chunk_size) {
    if (!idx_) { err_ = "no index"; return PGR_E_ARG; }
    if (seq_data_.size() != seqs_.size()) { err_ = "write_to_frag_files needs the sequences (keep_sequences(true) before loading)"; return PGR_E_ARG; }
    std::vector<uint32_t> sids;
    std::vector<const uint8_t *> ptrs;
    std::vector<size_t> lens;
    for (size_t i = 0; i < seqs_.size(); i++) { sids.push_back(seqs_[i].id); ptrs.push_back(seq_data_[i].p); lens.push_back(seq_data_[i].len); }
    pgr_fragment *frags = nullptr;
    pgr_aln_seg *segs = nullptr;
    size_t nf = 0, nsg = 0;
    double t0 = now_s();
    int rc = pgr_b200_index_compress_fragments(idx_, seqs_.size(), sids.data(), ptrs.data(), lens.data(), &frags, &nf, &segs, &nsg);
    timing_.frag_gpu_s += now_s() - t0;
    if (rc != PGR_OK) { err_ = pgr_b200_last_error(); return rc; }
    t0 = now_s();
    rc = write_frag_store(prefix, chunk_size, spec_.k, frags, nf, segs, seqs_, seq_data_, err_);
    timing_.frag_encode_s += now_s() - t0;
    pgr_b200_free(frags);
    pgr_b200_free(segs);
    return rc;
}

int SeqIndexDB::write_shmmr_map_index(const std::string &prefix) {
    if (!idx_) { err_ = "no index"; return PGR_E_ARG; }
    const double t0 = now_s();
    int rc = pgr_b200_index_write_mdb(idx_, (prefix + ".mdb").c_str());
    timing_.mdb_write_s += now_s() - t0;
    if (rc != PGR_OK) { err_ = pgr_b200_last_error(); return rc; }
    FILE *f = fopen((prefix + ".midx").c_str(), "wb");
    if (!f) { err_ = "file create error"; return PGR_E_IO; }
    for (auto &s : seqs_) fprintf(f, "%u\t%llu\t%s\t%s\n", s.id, (unsigned long long)s.len, s.name.c_str(), s.source.c_str());
    fclose(f);
    return PGR_OK;
}

}  // namespace pgrb200
