// pgr-b200-mdb — same command line as pgr-mdb (pgr-bin/src/bin/pgr-mdb.rs:27-50):
//   pgr-b200-mdb <filelist of .agc files> <prefix> [-w 80] [-k 56] [-r 4] [--min-span 64] [--sketch] [--prefetching]
//                [--number-of-readers 4] [--gpus N] [--timing]
// Builds the SHIMMER index of the AGC archives on the B200 and writes <prefix>.mdb + <prefix>.midx
// (load_write_index_from_agcfile, pgr-mdb.rs:53-79 -> CompactSeqDB::load_index_from_agcfile, seq_db.rs:676-683 ->
// load_index_from_reader :541-571 -> load_index_from_seq_vec :573-615): fragment ids are per-sequence pair ordinals
// (seq_to_index, frg_id_mode 1), sequence ids RESTART AT 0 FOR EVERY ARCHIVE of the list (`let mut sid = 0` in
// load_index_from_reader, :543) — reproduced as is.  Contigs are decoded by reader threads with one archive handle each
// (agc_io.rs:219-333) while the GPU indexes the previous batch.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <future>
#include <string>
#include <vector>

#include "../../include/pgr_b200.h"
#include "agc_reader.hpp"

static std::string trim(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}

int main(int argc, char **argv) {
    pgr_shmmr_spec spec{80, 56, 4, 64, 0};           // pgr-mdb.rs:33-44
    int n_readers = 4, n_gpus = 1;
    bool prefetching = false, timing = false;
    std::string filelist, prefix;
    const char *usage = "usage: pgr-b200-mdb <filelist of .agc files> <prefix> [-w 80] [-k 56] [-r 4] [--min-span 64] [--sketch] [--prefetching] [--number-of-readers 4] [--gpus N] [--timing]\n";
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto need = [&](const char *name) -> uint32_t {
            if (i + 1 >= argc) { fprintf(stderr, "error: %s needs a value\n", name); exit(2); }
            return (uint32_t)strtoul(argv[++i], nullptr, 10);
        };
        if (a == "-w") spec.w = need("-w");
        else if (a == "-k") spec.k = need("-k");
        else if (a == "-r") spec.r = need("-r");
        else if (a == "--min-span" || a == "-m") spec.min_span = need("--min-span");
        else if (a == "--sketch") spec.sketch = 1;
        else if (a == "--prefetching") prefetching = true;
        else if (a == "--number-of-readers") n_readers = (int)need("--number-of-readers");
        else if (a == "--gpus") n_gpus = (int)need("--gpus");
        else if (a == "--timing") timing = true;
        else if (a == "-h" || a == "--help") { printf("%s", usage); return 0; }
        else if (filelist.empty()) filelist = a;
        else if (prefix.empty()) prefix = a;
        else { fprintf(stderr, "error: unexpected argument %s\n", a.c_str()); return 2; }
    }
    if (filelist.empty() || prefix.empty()) { fprintf(stderr, "%s", usage); return 2; }
    std::ifstream in(filelist);
    if (!in) { fprintf(stderr, "can't open the input file that contains the paths to the agc files\n"); return 1; }
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count(); };
    pgr_b200_mindex *m = nullptr;
    pgr_b200_index *idx = nullptr;
    if (n_gpus > 1) m = pgr_b200_mindex_new(&spec, 1 /* AGC fragment numbering */, n_gpus);
    else idx = pgr_b200_index_new(&spec, 1, -1);
    if (!m && !idx) { fprintf(stderr, "%s\n", pgr_b200_last_error()); return 1; }
    struct MidxRow { uint32_t sid; size_t len; std::string name, source; };
    std::vector<MidxRow> midx;
    double decode_wait_s = 0, gpu_s = 0;
    uint64_t bases = 0;
    std::string line, err;
    while (std::getline(in, line)) {
        const std::string path = trim(line);
        if (path.empty()) continue;
        pgrb200::AgcFile agc;
        if (!agc.open(path, prefetching, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        const auto &ctgs = agc.contigs();
        // batches of contigs: up to 1024 (the reference's decode batch, agc_io.rs:223) or ~1 Gbases; batch b+1 is decoded while
        // the GPU indexes batch b
        std::vector<std::pair<size_t, size_t>> batches;
        for (size_t i = 0; i < ctgs.size();) {
            size_t j = i, b = 0;
            while (j < ctgs.size() && j - i < 1024 && (j == i || b + ctgs[j].len <= (1ull << 30))) { b += ctgs[j].len; j++; }
            batches.emplace_back(i, j);
            i = j;
        }
        auto decode = [&](size_t bi) {
            auto out = std::make_shared<std::vector<std::vector<uint8_t>>>();
            std::string e;
            if (!agc.fetch(batches[bi].first, batches[bi].second, n_readers, *out, e)) { out->clear(); fprintf(stderr, "%s\n", e.c_str()); }
            return out;
        };
        std::future<std::shared_ptr<std::vector<std::vector<uint8_t>>>> nextf;
        if (!batches.empty()) nextf = std::async(std::launch::async, decode, (size_t)0);
        uint32_t sid = 0;                             // restarts for every archive (seq_db.rs:543)
        for (size_t bi = 0; bi < batches.size(); bi++) {
            const double t0 = since();
            auto seqs = nextf.get();
            decode_wait_s += since() - t0;
            if (bi + 1 < batches.size()) nextf = std::async(std::launch::async, decode, bi + 1);
            const size_t i0 = batches[bi].first, n = batches[bi].second - i0;
            if (seqs->size() != n) return 1;
            std::vector<uint32_t> sids(n);
            std::vector<const uint8_t *> ptrs(n);
            std::vector<size_t> lens(n);
            for (size_t q = 0; q < n; q++) {
                sids[q] = sid; ptrs[q] = (*seqs)[q].data(); lens[q] = (*seqs)[q].size();
                midx.push_back({sid, lens[q], ctgs[i0 + q].name, ctgs[i0 + q].sample});
                bases += lens[q];
                sid++;
            }
            const double t1 = since();
            const int rc = m ? pgr_b200_mindex_add_batch(m, n, sids.data(), ptrs.data(), lens.data()) : pgr_b200_index_add_batch(idx, n, sids.data(), ptrs.data(), lens.data());
            gpu_s += since() - t1;
            if (rc != PGR_OK) { fprintf(stderr, "%s\n", pgr_b200_last_error()); return 1; }
        }
    }
    const double t_index = since();
    // write_shmmr_map_index (seq_db.rs:790-810): <prefix>.mdb (keys ascending) + <prefix>.midx
    const int rc = m ? pgr_b200_mindex_write_mdb(m, (prefix + ".mdb").c_str()) : pgr_b200_index_write_mdb(idx, (prefix + ".mdb").c_str());
    if (rc != PGR_OK) { fprintf(stderr, "%s\n", pgr_b200_last_error()); return 1; }
    FILE *f = fopen((prefix + ".midx").c_str(), "wb");
    if (!f) { fprintf(stderr, "file create error\n"); return 1; }
    for (auto &r : midx) fprintf(f, "%u\t%zu\t%s\t%s\n", r.sid, r.len, r.name.c_str(), r.source.c_str());
    fclose(f);
    if (timing)
        fprintf(stderr, "{\"contigs\": %zu, \"bases\": %llu, \"gpus\": %d, \"readers\": %d, \"wall_s\": %.4f, \"index_wall_s\": %.4f, \"wait_for_decoder_s\": %.4f, \"gpu_index_calls_s\": %.4f}\n",
                midx.size(), (unsigned long long)bases, n_gpus, n_readers, since(), t_index, decode_wait_s, gpu_s);
    if (m) pgr_b200_mindex_free(m);
    if (idx) pgr_b200_index_free(idx);
    return 0;
}
