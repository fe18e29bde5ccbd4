#include "fastx_ingest.hpp"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <chrono>
#include <cstdlib>
#include <cstring>

#include "../../include/pgr_b200.h"

namespace pgrb200 {

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

FileBuf::~FileBuf() {
    if (pinned) pgr_b200_host_unregister(p);
    free(p);
}

static bool grow(FileBuf &b, size_t need) {
    if (need <= b.cap) return true;
    size_t ncap = std::max(need, b.cap * 2);
    ncap = (ncap + 4095) & ~(size_t)4095;
    void *np = nullptr;
    if (posix_memalign(&np, 4096, ncap) != 0) return false;
    if (b.size) memcpy(np, b.p, b.size);
    free(b.p);
    b.p = (uint8_t *)np; b.cap = ncap;
    return true;
}

// whole file into a page-aligned buffer; gzip members are inflated (zlib reads plain files as-is, but read(2) is faster)
static bool read_whole(const std::string &path, FileBuf &b, std::string &err) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) { err = "cannot open " + path; return false; }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); err = "cannot stat " + path; return false; }
    unsigned char magic[2] = {0, 0};
    const ssize_t got = pread(fd, magic, 2, 0);
    const bool gz = got == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    if (!gz) {
        if (!grow(b, (size_t)st.st_size + 64)) { close(fd); err = "out of memory reading " + path; return false; }
        size_t off = 0;
        while (off < (size_t)st.st_size) {
            const ssize_t r = read(fd, b.p + off, (size_t)st.st_size - off);
            if (r < 0) { close(fd); err = "read error on " + path; return false; }
            if (r == 0) break;
            off += (size_t)r;
        }
        b.size = off;
        close(fd);
        return true;
    }
    close(fd);
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    gzbuffer(f, 1 << 20);
    if (!grow(b, (size_t)st.st_size * 4 + (1 << 20))) { gzclose(f); err = "out of memory reading " + path; return false; }
    for (;;) {
        if (b.cap - b.size < (1u << 20) && !grow(b, b.cap + (b.cap >> 1) + (1 << 20))) { gzclose(f); err = "out of memory reading " + path; return false; }
        const int r = gzread(f, b.p + b.size, (unsigned)std::min<size_t>(b.cap - b.size, 1u << 30));
        if (r < 0) { gzclose(f); err = "read error on " + path; return false; }
        if (r == 0) break;
        b.size += (size_t)r;
    }
    gzclose(f);
    return true;
}

// id of a header line [p, line_end): up to the first ' ', minus '\n' '\r' (fasta_io.rs:79-86, :126-133)
static std::string take_id(const uint8_t *p, const uint8_t *line_end) {
    const uint8_t *sp = (const uint8_t *)memchr(p, ' ', (size_t)(line_end - p));
    const uint8_t *e = sp ? sp : line_end;
    std::string id;
    id.reserve((size_t)(e - p));
    for (const uint8_t *q = p; q < e; q++) if (*q != '\n' && *q != '\r') id.push_back((char)*q);
    return id;
}

// FASTA (fasta_io.rs:65-118): the reader's constructor has consumed the first byte; a record is the header line, then every
// byte up to the next '>' without '\n' and '\r'.  Compacts the sequence bytes towards `wp` (wp <= read position).
static void parse_fasta(FileBuf &b, ParsedFile &out) {
    uint8_t *const base = b.p;
    const uint8_t *const end = b.p + b.size;
    uint8_t *wp = base;
    const uint8_t *p = base + 1;
    while (p < end) {                                                    // read_until returned 0 -> None (fasta_io.rs:90-93)
        const uint8_t *nl = (const uint8_t *)memchr(p, '\n', (size_t)(end - p));
        const uint8_t *le = nl ? nl + 1 : end;
        out.ids.push_back(take_id(p, le));
        p = le;
        const uint8_t *gt = (const uint8_t *)memchr(p, '>', (size_t)(end - p));
        const uint8_t *se = gt ? gt : end;
        uint8_t *s0 = wp;
        while (p < se) {                                                 // one line at a time
            const uint8_t *e = (const uint8_t *)memchr(p, '\n', (size_t)(se - p));
            const uint8_t *ln_end = e ? e : se;
            size_t n = (size_t)(ln_end - p);
            if (n && memchr(p, '\r', n)) {                               // '\r' anywhere in the line is dropped (usually only at its end)
                for (const uint8_t *q = p; q < ln_end; q++) if (*q != '\r') *wp++ = *q;
            } else {
                if (wp != p) memmove(wp, p, n);
                wp += n;
            }
            p = e ? e + 1 : se;
        }
        SeqSpan sp; sp.p = s0; sp.len = (size_t)(wp - s0);
        out.seqs.push_back(sp);
        out.bases += sp.len;
        if (p < end) p++;                                                // the '>' of the next record is consumed
    }
}

// FASTQ (fasta_io.rs:120-165): id line, ONE sequence line, then read_until('+'), two read_until('\n'), read_until('@'); when
// that last call consumes nothing (the file ends right after the quality line) the reference returns None WITHOUT yielding
// the record it has just parsed — reproduced here (the last record of a FASTQ that ends in '\n' is dropped).
static void parse_fastq(FileBuf &b, ParsedFile &out) {
    uint8_t *const base = b.p;
    const uint8_t *const end = b.p + b.size;
    uint8_t *wp = base;
    const uint8_t *p = base + 1;
    auto line_end_from = [&](const uint8_t *q) { const uint8_t *nl = q < end ? (const uint8_t *)memchr(q, '\n', (size_t)(end - q)) : nullptr; return nl ? nl + 1 : end; };
    for (;;) {
        const uint8_t *le = line_end_from(p);
        std::string id = take_id(p, le);
        p = le;
        const uint8_t *se = line_end_from(p);
        uint8_t *s0 = wp;
        for (const uint8_t *q = p; q < se; q++) if (*q != '\n' && *q != '\r') *wp++ = *q;
        p = se;
        const uint8_t *plus = p < end ? (const uint8_t *)memchr(p, '+', (size_t)(end - p)) : nullptr;   // read_until(b'+')
        p = plus ? plus + 1 : end;
        p = line_end_from(p);                                            // rest of the '+' line
        p = line_end_from(p);                                            // quality line
        const uint8_t *at = p < end ? (const uint8_t *)memchr(p, '@', (size_t)(end - p)) : nullptr;     // read_until(b'@')
        const size_t consumed = (size_t)((at ? at + 1 : end) - p);
        p = at ? at + 1 : end;
        if (consumed == 0) { wp = s0; break; }                           // res == Some(0) -> None: the record is not yielded
        out.ids.push_back(std::move(id));
        SeqSpan sp; sp.p = s0; sp.len = (size_t)(wp - s0);
        out.seqs.push_back(sp);
        out.bases += sp.len;
    }
}

void parse_fastx_file(const std::string &path, bool pin, ParsedFile &out) {
    out.path = path;
    out.ok = false;
    out.buf.reset(new FileBuf());
    double t0 = now_s();
    if (!read_whole(path, *out.buf, out.err)) return;
    double t1 = now_s();
    out.read_s = t1 - t0;
    if (out.buf->size == 0) { out.err = "empty file: " + path; return; }  // fasta_io.rs:58-63
    if (out.buf->p[0] == '@') parse_fastq(*out.buf, out); else parse_fasta(*out.buf, out);
    double t2 = now_s();
    out.parse_s = t2 - t1;
    if (pin && out.bases) {
        // page-lock what holds sequence bytes now (the compacted front of the buffer)
        const uint8_t *last = out.seqs.empty() ? out.buf->p : out.seqs.back().p + out.seqs.back().len;
        const size_t bytes = ((size_t)(last - out.buf->p) + 4095) & ~(size_t)4095;
        if (pgr_b200_host_register(out.buf->p, std::min(bytes, out.buf->cap)) == PGR_OK) out.buf->pinned = true;   // failure only costs speed
    }
    out.pin_s = now_s() - t2;
    out.ok = true;
}

FastxPipeline::FastxPipeline(std::vector<std::string> paths, int n_readers, bool pin, size_t window)
    : paths_(std::move(paths)), pin_(pin), window_(window ? window : (size_t)std::max(2, n_readers) * 2) {
    slot_.resize(paths_.size());
    done_.assign(paths_.size(), 0);
    n_readers = std::max(1, std::min<int>(n_readers, (int)std::max<size_t>(1, paths_.size())));
    for (int i = 0; i < n_readers; i++) threads_.emplace_back([this] { worker(); });
}

FastxPipeline::~FastxPipeline() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto &t : threads_) t.join();
}

void FastxPipeline::worker() {
    for (;;) {
        size_t i;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || next_claim_ >= paths_.size() || next_claim_ < taken_ + window_; });
            if (stop_ || next_claim_ >= paths_.size()) return;
            i = next_claim_++;
        }
        std::unique_ptr<ParsedFile> pf(new ParsedFile());
        parse_fastx_file(paths_[i], pin_, *pf);
        {
            std::lock_guard<std::mutex> lk(mu_);
            slot_[i] = std::move(pf);
            done_[i] = 1;
        }
        cv_.notify_all();
    }
}

bool FastxPipeline::ready(size_t i) {
    std::lock_guard<std::mutex> lk(mu_);
    return i < done_.size() && done_[i];
}

std::unique_ptr<ParsedFile> FastxPipeline::take(size_t i) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return done_[i] != 0; });
    std::unique_ptr<ParsedFile> r = std::move(slot_[i]);
    taken_ = i + 1;
    lk.unlock();
    cv_.notify_all();
    return r;
}

}  // namespace pgrb200
