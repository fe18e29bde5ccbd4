// pgr-b200-make-frgdb — same command line as pgr-make-frgdb (pgr-bin/src/bin/pgr-make-frgdb.rs:18-46):
//   pgr-b200-make-frgdb <filelist> <prefix> [-w 80] [-k 56] [-r 4] [--min-span 64]
// Builds the SHIMMER index of the FASTA/FASTQ(.gz) files listed in <filelist> on the B200 and writes
// <prefix>.mdb + <prefix>.midx + <prefix>.sdx + <prefix>.frg (ext.rs:201-207 write_frag_and_index_files); the fragments are
// compressed on the GPU (pgr_b200_index_compress_fragments).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>

#include "seq_index_db.hpp"

static std::string trim(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}

int main(int argc, char **argv) {
    uint32_t w = 80, k = 56, r = 4, min_span = 64;   // pgr-make-frgdb.rs:21-31
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
        else if (a == "-h" || a == "--help") {
            printf("usage: pgr-b200-make-frgdb <filelist> <prefix> [-w 80] [-k 56] [-r 4] [--min-span 64]\n");
            return 0;
        } else if (filelist.empty()) filelist = a;
        else if (prefix.empty()) prefix = a;
        else { fprintf(stderr, "error: unexpected argument %s\n", a.c_str()); return 2; }
    }
    if (filelist.empty() || prefix.empty()) { fprintf(stderr, "usage: pgr-b200-make-frgdb <filelist> <prefix> [-w 80] [-k 56] [-r 4] [--min-span 64]\n"); return 2; }
    std::ifstream in(filelist);
    if (!in) { fprintf(stderr, "can't open the input file that contains the paths to the fastx files\n"); return 1; }
    pgrb200::SeqIndexDB sdb;
    sdb.keep_sequences(true);   // the fragment store holds the bases
    std::string line;
    size_t fid = 0;
    while (std::getline(in, line)) {
        const std::string path = trim(line);
        const int rc = fid == 0 ? sdb.load_from_fastx(path, w, k, r, min_span) : sdb.append_from_fastx(path);
        if (rc != PGR_OK) { fprintf(stderr, "fail to read the fastx file: %s (%s)\n", path.c_str(), sdb.error().c_str()); return 1; }
        fid++;
    }
    if (fid == 0) { fprintf(stderr, "empty file list\n"); return 1; }
    int rc = sdb.write_to_frag_files(prefix);                // seq_db.rs:814-873
    if (rc != PGR_OK) { fprintf(stderr, "%s\n", sdb.error().c_str()); return 1; }
    rc = sdb.write_shmmr_map_index(prefix);                  // seq_db.rs:790-810
    if (rc != PGR_OK) { fprintf(stderr, "%s\n", sdb.error().c_str()); return 1; }
    return 0;
}
