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
    if (!p) return;
    if (pool) pool->give_back(p, cap, pinned);
    else if (pinned) pgr_b200_host_free(p);
    else free(p);
}

PinnedPool::~PinnedPool() {
    for (auto &it : free_) { if (in_slab(it.p)) continue; if (it.pinned) pgr_b200_host_free(it.p); else free(it.p); }
    if (slab_) pgr_b200_host_free(slab_);
}

void PinnedPool::reserve(size_t n, size_t bytes) {
    bytes = (bytes + 4095) & ~(size_t)4095;
    if (!n || !bytes || slab_) return;
    slab_ = (uint8_t *)pgr_b200_host_alloc(n * bytes);
    if (!slab_) return;                               // no device / no memory: acquire() falls back to single allocations
    slab_bytes_ = n * bytes;
    std::lock_guard<std::mutex> lk(mu_);
    for (size_t i = 0; i < n; i++) free_.push_back({slab_ + i * bytes, bytes, true});
}

void PinnedPool::give_back(uint8_t *p, size_t cap, bool pinned) {
    std::lock_guard<std::mutex> lk(mu_);
    free_.push_back({p, cap, pinned});
}

void PinnedPool::acquire(FileBuf &b, size_t bytes) {
    {
        std::lock_guard<std::mutex> lk(mu_);
        int best = -1;
        for (size_t i = 0; i < free_.size(); i++)
            if (free_[i].cap >= bytes && (best < 0 || free_[i].cap < free_[best].cap)) best = (int)i;
        if (best < 0 && !free_.empty()) {            // none is large enough: retire the smallest, allocate a larger one below
            size_t sm = 0;
            for (size_t i = 1; i < free_.size(); i++) if (free_[i].cap < free_[sm].cap) sm = i;
            if (!in_slab(free_[sm].p)) {
                if (free_[sm].pinned) pgr_b200_host_free(free_[sm].p); else free(free_[sm].p);
                free_.erase(free_.begin() + (ptrdiff_t)sm);
            }
        }
        if (best >= 0) {
            b.p = free_[best].p; b.cap = free_[best].cap; b.pinned = free_[best].pinned; b.pool = this; b.size = 0;
            free_.erase(free_.begin() + best);
            return;
        }
    }
    const size_t cap = ((bytes + bytes / 8 + (1u << 20)) + 4095) & ~(size_t)4095;
    void *np = pgr_b200_host_alloc(cap);
    b.pinned = np != nullptr;
    if (!np && posix_memalign(&np, 4096, cap) != 0) np = nullptr;   // no device (CPU tests): plain memory
    b.p = (uint8_t *)np; b.cap = np ? cap : 0; b.pool = this; b.size = 0;
}

// make room for `need` bytes, keeping the first b.size bytes
static bool grow(FileBuf &b, size_t need) {
    if (need <= b.cap) return true;
    const size_t ncap = ((std::max(need, b.cap * 2)) + 4095) & ~(size_t)4095;
    FileBuf nb;
    if (b.pool) b.pool->acquire(nb, ncap);
    else { void *np = nullptr; if (posix_memalign(&np, 4096, ncap) == 0) { nb.p = (uint8_t *)np; nb.cap = ncap; } }
    if (!nb.p) return false;
    if (b.size) memcpy(nb.p, b.p, b.size);
    std::swap(b.p, nb.p); std::swap(b.cap, nb.cap); std::swap(b.pinned, nb.pinned); std::swap(b.pool, nb.pool);
    return true;                                                     // nb's destructor disposes of the old block
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

void parse_fastx_file(const std::string &path, IngestMode mode, PinnedPool *pool, ParsedFile &out) {
    out.path = path;
    out.ok = false;
    out.buf.reset(new FileBuf());
    if (mode == INGEST_PINNED && pool) out.buf->pool = pool;            // grow() then draws page-locked blocks from the pool
    double t0 = now_s();
    if (!read_whole(path, *out.buf, out.err)) return;
    double t1 = now_s();
    out.read_s = t1 - t0;
    if (out.buf->size == 0) { out.err = "empty file: " + path; return; }  // fasta_io.rs:58-63
    if (out.buf->p[0] == '@') parse_fastq(*out.buf, out); else parse_fasta(*out.buf, out);
    double t2 = now_s();
    out.parse_s = t2 - t1;
    if (mode == INGEST_KEEP && pool && out.bases) {
        // the caller keeps the pageable bytes; the GPU call reads a compact page-locked copy that goes back to the pool
        out.gpu_buf.reset(new FileBuf());
        pool->acquire(*out.gpu_buf, out.bases + 64);
        if (out.gpu_buf->p) {
            size_t off = 0;
            for (auto &sp : out.seqs) {
                memcpy(out.gpu_buf->p + off, sp.p, sp.len);
                SeqSpan g; g.p = out.gpu_buf->p + off; g.len = sp.len;
                out.gpu_seqs.push_back(g);
                off += sp.len;
            }
        } else {
            out.gpu_buf.reset();
        }
    }
    out.pin_s = now_s() - t2;
    out.ok = true;
}

FastxPipeline::FastxPipeline(std::vector<std::string> paths, int n_readers, IngestMode mode, size_t window)
    : paths_(std::move(paths)), mode_(mode), window_(window ? window : (size_t)std::max(2, n_readers) + 4) {
    slot_.resize(paths_.size());
    done_.assign(paths_.size(), 0);
    if (mode_ != INGEST_PLAIN) {
        // page-locked slots for the files in flight, sized for the largest file (gzip: its inflated size is a guess, a
        // slot that turns out too small is replaced on the fly)
        size_t mx = 0;
        for (auto &p : paths_) {
            struct stat st;
            if (stat(p.c_str(), &st) == 0) {
                const bool gz = p.size() > 3 && p.compare(p.size() - 3, 3, ".gz") == 0;
                mx = std::max(mx, (size_t)st.st_size * (gz ? 4 : 1));
            }
        }
        pool_.reserve(std::min(paths_.size(), window_ + 4), mx + (1u << 16));   // files in flight (<= window) + the consumer's current batch (<= 4 per GPU pair)
    }
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
        parse_fastx_file(paths_[i], mode_, &pool_, *pf);
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
