#include "seq_index_db.hpp"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace pgrb200 {

// read a whole file, transparently inflating gzip (seq_db.rs:420-454 sniffs the magic bytes 0x1f 0x8b)
static bool slurp(const std::string &path, std::vector<uint8_t> &buf, std::string &err) {
    gzFile f = gzopen(path.c_str(), "rb");   // zlib reads plain files as-is
    if (!f) { err = "cannot open " + path; return false; }
    uint8_t tmp[1 << 16];
    int got;
    while ((got = gzread(f, tmp, sizeof tmp)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    const bool ok = got == 0;
    gzclose(f);
    if (!ok) err = "read error on " + path;
    return ok;
}

// fasta_io.rs:46-172.  The reader's constructor consumes the first byte of the file ('>' or '@').
bool read_fastx(const std::string &path, std::vector<SeqRec> &out, std::string &err) {
    std::vector<uint8_t> b;
    if (!slurp(path, b, err)) return false;
    if (b.empty()) { err = "empty file: " + path; return false; }       // fasta_io.rs:58-63
    const bool fastq = b[0] == '@';
    size_t p = 1;
    const size_t N = b.size();
    auto take_id = [&](size_t line_end) {
        std::string id;
        for (size_t i = p; i < line_end; i++) {
            const uint8_t c = b[i];
            if (c == ' ') break;                                         // read_until(b' ')
            if (c != '\n' && c != '\r') id.push_back((char)c);
        }
        return id;
    };
    auto line_end_from = [&](size_t q) { while (q < N && b[q] != '\n') q++; return q < N ? q + 1 : N; };
    if (!fastq) {
        for (;;) {
            if (p >= N) break;                                           // read_until returned 0 -> None (fasta_io.rs:90-93)
            const size_t le = line_end_from(p);
            SeqRec rec;
            rec.id = take_id(le);
            p = le;
            while (p < N && b[p] != '>') { const uint8_t c = b[p]; if (c != '\n' && c != '\r') rec.seq.push_back(c); p++; }
            if (p < N) p++;                                              // the '>' of the next record is consumed
            out.push_back(std::move(rec));
        }
    } else {
        for (;;) {                                                       // fasta_io.rs:120-165
            const size_t le = line_end_from(p);
            SeqRec rec;
            rec.id = take_id(le);
            p = le;
            const size_t se = line_end_from(p);
            for (size_t i = p; i < se; i++) if (b[i] != '\n' && b[i] != '\r') rec.seq.push_back(b[i]);
            p = se;
            while (p < N && b[p] != '+') p++;                            // read_until(b'+')
            if (p < N) p++;
            p = line_end_from(p);                                        // rest of the '+' line
            p = line_end_from(p);                                        // quality line
            size_t q = p;
            while (q < N && b[q] != '@') q++;                            // read_until(b'@')
            const size_t consumed = (q < N ? q + 1 : N) - p;
            p = q < N ? q + 1 : N;
            out.push_back(std::move(rec));
            if (consumed == 0) break;                                    // res == Some(0) -> None
        }
    }
    return true;
}

SeqIndexDB::~SeqIndexDB() { if (idx_) pgr_b200_index_free(idx_); }

int SeqIndexDB::load_from_fastx(const std::string &path, uint32_t w, uint32_t k, uint32_t r, uint32_t min_span) {
    spec_.w = w; spec_.k = k; spec_.r = r; spec_.min_span = min_span; spec_.sketch = 0;
    if (idx_) { pgr_b200_index_free(idx_); idx_ = nullptr; }
    seqs_.clear();
    seq_data_.clear();
    idx_ = pgr_b200_index_new(&spec_, 0 /* FASTX fragment numbering */, -1);
    if (!idx_) { err_ = pgr_b200_last_error(); return PGR_E_NO_DEVICE; }
    return load_seqs_from_fastx(path);
}

int SeqIndexDB::append_from_fastx(const std::string &path) {
    if (!idx_) { err_ = "Only DB created with load_from_fastx() can add data from another fastx file"; return PGR_E_ARG; }
    return load_seqs_from_fastx(path);
}

// seq_db.rs:471-525: sid continues from seqs.len(); records are handed to the GPU in the reference's batches of <=129
int SeqIndexDB::load_seqs_from_fastx(const std::string &path) {
    std::vector<SeqRec> recs;
    if (!read_fastx(path, recs, err_)) return PGR_E_IO;
    uint32_t sid = (uint32_t)seqs_.size();
    std::vector<uint32_t> sids;
    std::vector<const uint8_t *> ptrs;
    std::vector<size_t> lens;
    for (auto &r : recs) {
        sids.push_back(sid);
        ptrs.push_back(r.seq.data());
        lens.push_back(r.seq.size());
        CompactSeq cs;
        cs.id = sid; cs.len = r.seq.size(); cs.name = r.id; cs.source = path;
        seqs_.push_back(std::move(cs));
        sid++;
    }
    // one GPU call per file: the batch boundary (129 records) only sets the reference's parallel granularity, it has
    // no effect on results (fragment ids are a running counter across batches)
    const int rc = pgr_b200_index_add_batch(idx_, recs.size(), sids.data(), ptrs.data(), lens.data());
    if (rc != PGR_OK) err_ = pgr_b200_last_error();
    if (keep_seqs_) for (auto &r : recs) seq_data_.push_back(std::move(r.seq));
    return rc;
}

int SeqIndexDB::load_from_index_files(const std::string &prefix) {
    if (idx_) { pgr_b200_index_free(idx_); idx_ = nullptr; }
    seqs_.clear();
    seq_data_.clear();
    idx_ = pgr_b200_index_read_mdb((prefix + ".mdb").c_str(), -1);
    if (!idx_) { err_ = pgr_b200_last_error(); return PGR_E_IO; }
    pgr_b200_index_get_spec(idx_, &spec_);
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

bool SeqIndexDB::get_sub_seq_by_id(uint32_t sid, size_t bgn, size_t end, std::vector<uint8_t> &out) const {
    if (sid >= seq_data_.size() || bgn > end || end > seq_data_[sid].size()) return false;
    out.assign(seq_data_[sid].begin() + (ptrdiff_t)bgn, seq_data_[sid].begin() + (ptrdiff_t)end);
    return true;
}

int SeqIndexDB::query_fragment_to_hps(const std::vector<SeqRec> &queries, const pgr_query_params &params, pgr_query_result **out) {
    if (!idx_) { err_ = "no index"; return PGR_E_ARG; }
    std::vector<const uint8_t *> ptrs;
    std::vector<size_t> lens;
    for (auto &q : queries) { ptrs.push_back(q.seq.data()); lens.push_back(q.seq.size()); }
    const int rc = pgr_b200_query_batch(idx_, queries.size(), ptrs.data(), lens.data(), &params, out);
    if (rc != PGR_OK) err_ = pgr_b200_last_error();
    return rc;
}

int SeqIndexDB::write_shmmr_map_index(const std::string &prefix) {
    if (!idx_) { err_ = "no index"; return PGR_E_ARG; }
    int rc = pgr_b200_index_write_mdb(idx_, (prefix + ".mdb").c_str());
    if (rc != PGR_OK) { err_ = pgr_b200_last_error(); return rc; }
    FILE *f = fopen((prefix + ".midx").c_str(), "wb");
    if (!f) { err_ = "file create error"; return PGR_E_IO; }
    for (auto &s : seqs_) fprintf(f, "%u\t%llu\t%s\t%s\n", s.id, (unsigned long long)s.len, s.name.c_str(), s.source.c_str());
    fclose(f);
    return PGR_OK;
}

}  // namespace pgrb200
