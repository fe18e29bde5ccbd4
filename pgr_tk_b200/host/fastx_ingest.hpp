// fastx_ingest.hpp — host ingest for the index builders: file -> parse -> page-locked buffer -> GPU batch
// as a pipeline.  Reader threads take the files of the list in order, read each one whole (gzip inflated on the fly,
// seq_db.rs:420-454 sniffs the magic bytes), parse it IN PLACE with the reference's record rules (fasta_io.rs:46-172:
// id = header up to the first ' ', sequence = the bytes up to the next '>' minus '\n' '\r') — the sequence bytes are
// compacted towards the front of the file buffer, no second copy — and page-lock the buffer (pgr_b200_host_register), so
// that the library's H2D copies run at PCIe rate straight from it.  The consumer takes the files back in list order
// (sequence ids and fragment ids follow file order, seq_db.rs:471-525) while later files are still being parsed.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace pgrb200 {

struct SeqSpan { const uint8_t *p = nullptr; size_t len = 0; };

class PinnedPool;

// a file read into memory; sequences parsed in place point into it.  Page-locked buffers come from a pool and go back to
// it: page-locking per file (cudaHostRegister) costs ~0.5 ms/MB and stalls the CUDA calls of the consumer meanwhile.
struct FileBuf {
    uint8_t *p = nullptr;
    size_t size = 0, cap = 0;
    bool pinned = false;          // p came from pgr_b200_host_alloc
    PinnedPool *pool = nullptr;   // where a pinned buffer returns to
    FileBuf() = default;
    FileBuf(const FileBuf &) = delete;
    FileBuf &operator=(const FileBuf &) = delete;
    ~FileBuf();
};

// page-locked buffers reused across files (a buffer that is too small is replaced by a larger one)
class PinnedPool {
public:
    ~PinnedPool();
    // ONE page-locked allocation carved into n slots of `bytes` each, made before the readers start: cudaHostAlloc is slow
    // (~0.3 ms/MB) and holds up the CUDA calls of other threads while it runs
    void reserve(size_t n, size_t bytes);
    // a buffer of at least `bytes` (page-locked if the device allows it, plain memory otherwise)
    void acquire(FileBuf &b, size_t bytes);
    void give_back(uint8_t *p, size_t cap, bool pinned);
private:
    struct Item { uint8_t *p; size_t cap; bool pinned; };
    std::mutex mu_;
    std::vector<Item> free_;
    uint8_t *slab_ = nullptr;      // the reserved block; its slots are never freed one by one
    size_t slab_bytes_ = 0;
    bool in_slab(const uint8_t *p) const { return slab_ && p >= slab_ && p < slab_ + slab_bytes_; }
};

struct ParsedFile {
    std::string path, err;
    bool ok = false;
    std::unique_ptr<FileBuf> buf;          // the bytes `seqs` point into
    std::unique_ptr<FileBuf> gpu_buf;      // INGEST_KEEP: page-locked copy of the sequences for the H2D copy (`gpu_seqs`)
    std::vector<std::string> ids;
    std::vector<SeqSpan> seqs, gpu_seqs;
    uint64_t bases = 0;
    double read_s = 0, parse_s = 0, pin_s = 0;   // wall seconds of this file's stages (reader thread)
};

// INGEST_PLAIN: pageable memory only (queries, tests).  INGEST_PINNED: the file is read straight into a pooled page-locked
// buffer and parsed there (index-only builds: the buffer is recycled after the GPU call).  INGEST_KEEP: parsed in pageable
// memory that the caller keeps (the fragment store needs the bases) plus a pooled page-locked copy for the GPU call.
enum IngestMode { INGEST_PLAIN = 0, INGEST_PINNED = 1, INGEST_KEEP = 2 };
// read + parse one file (the reference's FastaReader / FastqReader rules)
void parse_fastx_file(const std::string &path, IngestMode mode, PinnedPool *pool, ParsedFile &out);

// parallel readers, in-order delivery, at most `window` parsed files waiting
class FastxPipeline {
public:
    FastxPipeline(std::vector<std::string> paths, int n_readers, IngestMode mode, size_t window = 0);
    ~FastxPipeline();
    size_t n_files() const { return paths_.size(); }
    // blocks until file `i` (0-based, must be requested in increasing order) is parsed; the caller takes ownership
    std::unique_ptr<ParsedFile> take(size_t i);
    // true if file i is already parsed (non-blocking): lets the consumer grow a batch with what is ready
    bool ready(size_t i);
private:
    void worker();
    std::vector<std::string> paths_;
    IngestMode mode_;
    PinnedPool pool_;
    size_t window_;
    std::mutex mu_;
    std::condition_variable cv_;
    size_t next_claim_ = 0, taken_ = 0;
    std::vector<std::unique_ptr<ParsedFile>> slot_;
    std::vector<char> done_;
    std::vector<std::thread> threads_;
    bool stop_ = false;
};

}  // namespace pgrb200
