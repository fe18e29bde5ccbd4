// pgr-b200-make-frgdb — same command line as pgr-make-frgdb (pgr-bin/src/bin/pgr-make-frgdb.rs:18-46):
//   pgr-b200-make-frgdb <filelist> <prefix> [-w 80] [-k 56] [-r 4] [--min-span 64]
//                       [--gpus N] [--readers R] [--index-only] [--timing]      (additions; not in the reference)
// --gpus N      build the map sharded over N GPUs of this box (one NCCL all-to-all merges it; same files, byte for byte)
// --readers R   reader/parser threads of the ingest pipeline (default 4, pgr-mdb's --number-of-readers default)
// --index-only  write .mdb + .midx only (what pgr-mdb writes); skips the fragment store
// --devices a,b,..  one device per shard instead of 0..N-1 (may repeat: the sharded build on one GPU, a test set-up)
// --timing      one JSON line on stderr with the wall seconds of every phase
// Builds the SHIMMER index of the FASTA/FASTQ(.gz) files listed in <filelist> on the B200 and writes
// <prefix>.mdb + <prefix>.midx + <prefix>.sdx + <prefix>.frg (ext.rs:201-207 write_frag_and_index_files); the fragments are
// compressed on the GPU (pgr_b200_index_compress_fragments).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "seq_index_db.hpp"

static std::string trim(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}

int main(int argc, char **argv) {
    uint32_t w = 80, k = 56, r = 4, min_span = 64;   // pgr-make-frgdb.rs:21-31
    int n_gpus = 1, n_readers = 4;
    bool index_only = false, timing = false;
    std::vector<int> devices;
    std::string filelist, prefix;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto need = [&](const char *name) -> uint32_t {
            if (i + 1 >= argc) { fprintf(stderr, "error: %s needs a value\n", name); exit(2); }
            return (uint32_t)strtoul(argv[++i], nullptr, 10);
        };
        if (a == "-w") w = need("-w");
        else if (a == "-k") k = need("-k");
        else if (a == "-r") r = need("-r");
        else if (a == "--min-span" || a == "-m") min_span = need("--min-span");
        else if (a == "--gpus") n_gpus = (int)need("--gpus");
        else if (a == "--readers") n_readers = (int)need("--readers");
        else if (a == "--devices") {
            if (i + 1 >= argc) { fprintf(stderr, "error: --devices needs a value\n"); return 2; }
            const std::string v = argv[++i];
            size_t p0 = 0;
            while (p0 <= v.size()) { const size_t c = v.find(',', p0); devices.push_back(atoi(v.substr(p0, c == std::string::npos ? c : c - p0).c_str())); if (c == std::string::npos) break; p0 = c + 1; }
        }
        else if (a == "--index-only") index_only = true;
        else if (a == "--timing") timing = true;
        else if (a == "-h" || a == "--help") {
            printf("usage: pgr-b200-make-frgdb <filelist> <prefix> [-w 80] [-k 56] [-r 4] [--min-span 64] [--gpus N] [--readers R] [--index-only] [--timing]\n");
            return 0;
        } else if (filelist.empty()) filelist = a;
        else if (prefix.empty()) prefix = a;
        else { fprintf(stderr, "error: unexpected argument %s\n", a.c_str()); return 2; }
    }
    if (!devices.empty()) n_gpus = (int)devices.size();
    if (filelist.empty() || prefix.empty()) { fprintf(stderr, "usage: pgr-b200-make-frgdb <filelist> <prefix> [-w 80] [-k 56] [-r 4] [--min-span 64]\n"); return 2; }
    std::ifstream in(filelist);
    if (!in) { fprintf(stderr, "can't open the input file that contains the paths to the fastx files\n"); return 1; }
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count(); };
    std::vector<std::string> paths;
    std::string line;
    while (std::getline(in, line)) { const std::string path = trim(line); if (!path.empty()) paths.push_back(path); }
    if (paths.empty()) { fprintf(stderr, "empty file list\n"); return 1; }
    pgrb200::SeqIndexDB sdb;
    sdb.keep_sequences(!index_only);   // the fragment store holds the bases
    int rc = sdb.load_from_fastx_list(paths, w, k, r, min_span, n_readers, n_gpus, devices);   // pgr-make-frgdb.rs:48-63
    if (rc != PGR_OK) { fprintf(stderr, "fail to build the index: %s\n", sdb.error().c_str()); return 1; }
    const double t_index = since();
    if (!index_only) {
        rc = sdb.write_to_frag_files(prefix);                // seq_db.rs:814-873
        if (rc != PGR_OK) { fprintf(stderr, "%s\n", sdb.error().c_str()); return 1; }
    }
    const double t_frags = since();
    rc = sdb.write_shmmr_map_index(prefix);                  // seq_db.rs:790-810
    if (rc != PGR_OK) { fprintf(stderr, "%s\n", sdb.error().c_str()); return 1; }
    const double t_all = since();
    if (timing) {
        const auto &t = sdb.timing();
        fprintf(stderr, "{\"files\": %zu, \"bases\": %llu, \"gpus\": %d, \"readers\": %d, \"wall_s\": %.4f, \"index_wall_s\": %.4f, \"frag_store_wall_s\": %.4f, "
                        "\"mdb_midx_wall_s\": %.4f, \"consumer\": {\"device_init_s\": %.4f, \"wait_for_parser_s\": %.4f, \"gpu_index_calls_s\": %.4f, \"finalize_merge_s\": %.4f, "
                        "\"frag_compress_gpu_s\": %.4f, \"frag_encode_deflate_write_s\": %.4f, \"mdb_write_s\": %.4f}, "
                        "\"reader_threads_total\": {\"read_s\": %.4f, \"parse_s\": %.4f, \"page_lock_s\": %.4f}}\n",
                paths.size(), (unsigned long long)t.bases, n_gpus, n_readers, t_all, t_index, t_frags - t_index, t_all - t_frags, t.device_init_s, t.wait_parse_s, t.gpu_index_s, t.merge_s,
                t.frag_gpu_s, t.frag_encode_s, t.mdb_write_s, t.reader_read_s, t.reader_parse_s, t.reader_pin_s);
    }
    return 0;
}
