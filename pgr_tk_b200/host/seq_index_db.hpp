// seq_index_db.hpp — C++ host mirror of the pgr-db interfaces on the SHIMMER indexing path, on top of the C ABI
// (include/pgr_b200.h).  The reference is Rust; this image has no rustc, so the host side above the ABI is C++ here:
//   FastaReader        fasta_io.rs:46-118  (record parsing rules; gz sniffing of seq_db.rs:420-454 via zlib)
//   SeqIndexDB         ext.rs:48-64,152-199 load_from_fastx / append_from_fastx (FASTX back end, index part)
//                      seq_db.rs:471-525 record batching (<=129 records per call), sid = running record index
//                      seq_db.rs:790-810 write_shmmr_map_index (.mdb + .midx)
//                      seq_db.rs:814-873 write_to_frag_files (.sdx + .frg)
//                      ext.rs:87-150 load_from_*_index, index part (.mdb + .midx read back; no sequence store)
//                      ext.rs:252-282 query_fragment_to_hps (batched), ext.rs:455-489 get_sub_seq_by_id (FASTX back end)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/pgr_b200.h"
#include "fastx_ingest.hpp"

namespace pgrb200 {

struct SeqRec { std::string id; std::vector<uint8_t> seq; };

// whole-file FASTA/FASTQ parser with the reference's exact field rules (fastx_ingest.cpp does the parsing)
bool read_fastx(const std::string &path, std::vector<SeqRec> &out, std::string &err);

struct CompactSeq { uint32_t id; uint64_t len; std::string name, source; };

// <prefix>.sdx + <prefix>.frg as a sequence store (frag_file_io.rs:14-229 CompactSeqFragFileStorage; seq_db.rs:685-735):
// bincode 2 standard-config tables, one raw-deflate stream per chunk of fragments, fragments decoded chunk-wise on demand
class FragStore {
public:
    struct Seg { uint32_t type, a, b; };
    struct Fragment { uint8_t kind = 1; bool reversed = false; uint32_t ref = 0, len = 0; std::vector<Seg> segs; std::vector<uint8_t> bases; };
    struct SeqEntry { std::string source, name; uint32_t id = 0, frag_first = 0, frag_count = 0; uint64_t len = 0; bool has_source = false; };
    bool load(const std::string &prefix, uint32_t k, std::string &err);
    bool loaded() const { return !frg_.empty() || !addr_.empty(); }
    const std::vector<SeqEntry> &seqs() const { return seqs_; }
    bool get_seq_by_id(uint32_t sid, std::vector<uint8_t> &out, std::string &err);
private:
    const Fragment *fragment(uint32_t id, std::string &err);
    struct Addr { uint64_t off, len, bases; };
    uint64_t chunk_size_ = 256;
    std::vector<Addr> addr_;
    std::vector<SeqEntry> seqs_;
    std::vector<uint8_t> frg_;
    uint32_t k_ = 0;
    std::vector<std::pair<uint64_t, std::vector<Fragment>>> cache_;   // (chunk id, fragments), a few recent chunks
};

// seq_db.rs:814-873 write_to_frag_files on fragment records laid out as pgr_b200_index_compress_fragments returns them
int write_frag_store(const std::string &prefix, size_t chunk_size, uint32_t k, const pgr_fragment *frags, size_t n_frags, const pgr_aln_seg *segs,
                     const std::vector<CompactSeq> &seqs, const std::vector<SeqSpan> &seq_data, std::string &err, int n_threads = 0);

class SeqIndexDB {
public:
    SeqIndexDB() = default;
    ~SeqIndexDB();
    SeqIndexDB(const SeqIndexDB &) = delete;
    SeqIndexDB &operator=(const SeqIndexDB &) = delete;
    // ext.rs:152-178
    int load_from_fastx(const std::string &path, uint32_t w, uint32_t k, uint32_t r, uint32_t min_span);
    // ext.rs:180-199
    int append_from_fastx(const std::string &path);
    // pgr-make-frgdb.rs:48-63 as a pipeline: the files of the list are read, parsed and page-locked by `n_readers` threads while
    // the GPU(s) index the files that are ready (same result as load_from_fastx + append_from_fastx file by file).  n_gpus > 1
    // builds the map sharded over the GPUs (pgr_b200_mindex_*, one NCCL all-to-all) and gathers it onto device 0 at the end.
    int load_from_fastx_list(const std::vector<std::string> &paths, uint32_t w, uint32_t k, uint32_t r, uint32_t min_span, int n_readers, int n_gpus,
                             const std::vector<int> &devices = {});   // devices: one per shard (may repeat: sharded build on one GPU, a test set-up)
    // wall seconds per phase of the calls above and of the writers (for the CLI's --timing report)
    struct Timing { double device_init_s = 0, wait_parse_s = 0, gpu_index_s = 0, merge_s = 0, frag_gpu_s = 0, frag_encode_s = 0, mdb_write_s = 0, reader_read_s = 0, reader_parse_s = 0, reader_pin_s = 0; uint64_t bases = 0; };
    const Timing &timing() const { return timing_; }
    // ext.rs:212-250 load_from_seq_list: sequences given in memory, sids in list order
    int load_from_seq_list(const std::vector<SeqRec> &seq_list, const std::string &source, uint32_t w, uint32_t k, uint32_t r, uint32_t min_span);
    // index part of ext.rs:87-150 (load_from_agc_index / load_from_frg_index): <prefix>.mdb + <prefix>.midx; sequences are
    // not available afterwards (the .agc / .frg stores are out of scope)
    // mdb_resident = true: read_mdb_file_to_frag_locations (seq_db.rs:1409-1471, what ext.rs:87-150 does): only the key table of the
    // .mdb is read, the file stays memory-mapped and queries go through pgr_b200_query_batch_mmap (ext.rs:285-342)
    int load_from_index_files(const std::string &prefix, bool mdb_resident = false);
    // ext.rs:131-150 load_from_frg_index: the index files plus the .sdx/.frg sequence store
    int load_from_frg_index(const std::string &prefix);
    // keep the sequences of load_from_fastx in host memory (needed by get_sub_seq_by_id); set before loading
    void keep_sequences(bool on) { keep_seqs_ = on; }
    // ext.rs:455-489 for the FASTX back end: seq[bgn..end); false when the sequence is not held
    bool get_sub_seq_by_id(uint32_t sid, size_t bgn, size_t end, std::vector<uint8_t> &out);
    // ext.rs:252-282 for a batch of queries (pgr-query.rs:135 runs one rayon task per query); result freed by the caller
    // with pgr_b200_query_result_free
    int query_fragment_to_hps(const std::vector<SeqRec> &queries, const pgr_query_params &params, pgr_query_result **out);
    // seq_db.rs:790-810 (index part of ext.rs:201-207 write_frag_and_index_files)
    int write_shmmr_map_index(const std::string &prefix);
    // seq_db.rs:814-873 write_to_frag_files: <prefix>.sdx + <prefix>.frg (fragments compressed on the GPU, bincode 2
    // standard-config encoding and one raw-deflate stream per 256-fragment chunk on the host); needs keep_sequences(true)
    int write_to_frag_files(const std::string &prefix, size_t chunk_size = 256);
    const std::vector<CompactSeq> &seqs() const { return seqs_; }
    pgr_b200_index *index() { return idx_; }
    const std::string &error() const { return err_; }

private:
    int load_seqs_from_fastx(const std::string &path);
    int add_records(std::vector<SeqRec> &recs, const std::string &source);
    int add_parsed(std::vector<std::unique_ptr<ParsedFile>> &files);
    void reset();
    pgr_b200_index *idx_ = nullptr;
    pgr_b200_mindex *midx_ = nullptr;   // multi-GPU build in progress
    pgr_b200_mdb_map *map_ = nullptr;   // the .mdb-resident back end (index left on disk)
    pgr_shmmr_spec spec_{};
    std::vector<CompactSeq> seqs_;
    std::vector<SeqSpan> seq_data_;                       // views of the sequences (FASTX back end) ...
    std::vector<std::unique_ptr<FileBuf>> file_bufs_;     // ... into the page-locked file buffers they were parsed in
    std::vector<std::vector<uint8_t>> owned_seqs_;        // ... or into these (load_from_seq_list)
    Timing timing_;
    bool keep_seqs_ = false;
    bool fastx_backend_ = false;   // created by load_from_fastx / load_from_seq_list: the only back end append_from_fastx accepts (ext.rs:180-199)
    FragStore frag_store_;
    std::string err_;
};

}  // namespace pgrb200
