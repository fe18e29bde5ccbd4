// agc_reader.hpp — AGC archives as an input of the index builder (pgr-mdb; pgr-db/src/agc_io.rs).
// AGC (refresh-bio/agc) is a third-party C++ library the reference vendors under agc/ and binds through its C API
// (agc/src/lib-cxx/agc-api.h; pgr-db/build.rs, agc_io.rs:60-130).  This reader binds the SAME C API at run time from
// libagc_ref.so (built from the reference's vendored sources by host/Makefile when /root/reference/agc is present; set
// PGR_B200_LIBAGC to use another build): no AGC code lives in this repository.
//   order of records  = samples in agc_list_sample order, contigs in agc_list_ctg order   (agc_io.rs:79-104)
//   one handle per reader thread, contigs decoded in parallel, delivered in order           (agc_io.rs:219-333)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pgrb200 {

struct AgcContig { std::string sample, name; size_t len; };

class AgcFile {
public:
    ~AgcFile();
    // false (err set) when libagc_ref.so or the archive cannot be opened
    bool open(const std::string &path, bool prefetching, std::string &err);
    const std::vector<AgcContig> &contigs() const { return ctgs_; }
    // decode contigs [i0, i1) with n_threads reader threads (each with its own archive handle) into out[i - i0]
    bool fetch(size_t i0, size_t i1, int n_threads, std::vector<std::vector<uint8_t>> &out, std::string &err);
private:
    std::string path_;
    bool prefetching_ = false;
    std::vector<AgcContig> ctgs_;
};

}  // namespace pgrb200
