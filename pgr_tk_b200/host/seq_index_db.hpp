// seq_index_db.hpp — C++ host mirror of the pgr-db interfaces on the SHIMMER indexing path, on top of the C ABI
// (include/pgr_b200.h).  The reference is Rust; this image has no rustc, so the host side above the ABI is C++ here:
//   FastaReader        fasta_io.rs:46-118  (record parsing rules; gz sniffing of seq_db.rs:420-454 via zlib)
//   SeqIndexDB         ext.rs:48-64,152-199 load_from_fastx / append_from_fastx (FASTX back end, index part)
//                      seq_db.rs:471-525 record batching (<=129 records per call), sid = running record index
//                      seq_db.rs:790-810 write_shmmr_map_index (.mdb + .midx)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/pgr_b200.h"

namespace pgrb200 {

struct SeqRec { std::string id; std::vector<uint8_t> seq; };

// whole-file FASTA/FASTQ parser with the reference's exact field rules
bool read_fastx(const std::string &path, std::vector<SeqRec> &out, std::string &err);

struct CompactSeq { uint32_t id; uint64_t len; std::string name, source; };

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
    // seq_db.rs:790-810 (index part of ext.rs:201-207 write_frag_and_index_files)
    int write_shmmr_map_index(const std::string &prefix);
    const std::vector<CompactSeq> &seqs() const { return seqs_; }
    pgr_b200_index *index() { return idx_; }
    const std::string &error() const { return err_; }

private:
    int load_seqs_from_fastx(const std::string &path);
    pgr_b200_index *idx_ = nullptr;
    pgr_shmmr_spec spec_{};
    std::vector<CompactSeq> seqs_;
    std::string err_;
};

}  // namespace pgrb200
