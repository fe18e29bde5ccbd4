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

// a file read into memory; sequences parsed in place point into it
struct FileBuf {
    uint8_t *p = nullptr;
    size_t size = 0, cap = 0;
    bool pinned = false;
    FileBuf() = default;
    FileBuf(const FileBuf &) = delete;
    FileBuf &operator=(const FileBuf &) = delete;
    ~FileBuf();
};

struct ParsedFile {
    std::string path, err;
    bool ok = false;
    std::unique_ptr<FileBuf> buf;
    std::vector<std::string> ids;
    std::vector<SeqSpan> seqs;
    uint64_t bases = 0;
    double read_s = 0, parse_s = 0, pin_s = 0;   // wall seconds of this file's stages (reader thread)
};

// read + parse one file (the reference's FastaReader / FastqReader rules); pin = page-lock the buffer afterwards
void parse_fastx_file(const std::string &path, bool pin, ParsedFile &out);

// parallel readers, in-order delivery, at most `window` parsed files waiting
class FastxPipeline {
public:
    FastxPipeline(std::vector<std::string> paths, int n_readers, bool pin, size_t window = 0);
    ~FastxPipeline();
    size_t n_files() const { return paths_.size(); }
    // blocks until file `i` (0-based, must be requested in increasing order) is parsed; the caller takes ownership
    std::unique_ptr<ParsedFile> take(size_t i);
    // true if file i is already parsed (non-blocking): lets the consumer grow a batch with what is ready
    bool ready(size_t i);
private:
    void worker();
    std::vector<std::string> paths_;
    bool pin_;
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
