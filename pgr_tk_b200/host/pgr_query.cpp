// pgr-b200-query — pgr-query (pgr-bin/src/bin/pgr-query.rs:17-79) on the B200:
//   pgr-b200-query <pgr_db_prefix | fastx> <query.fa> <output_prefix> [--fastx-file] [-w 80 -k 56 -r 4 --min-span 64]
//       [--gap-penalty-factor 0.025] [--merge-range-tol 100000] [--max-count 128] [--max-query-count 128]
//       [--max-target-count 128] [--max-aln-chain-span 8] [--only-summary] [--bed-summary]
// --fastx-file: the database is built from the FASTA/FASTQ(.gz) file (SeqIndexDB::load_from_fastx, ext.rs:152).
// --frg-file: <pgr_db_prefix>.mdb/.midx/.sdx/.frg are loaded (ext.rs:131 load_from_frg_index); target sub-sequences come from
// the fragment store.  Otherwise only <pgr_db_prefix>.mdb/.midx are read (the AGC store is out of scope): needs --only-summary.  All queries go to the GPU in ONE batched call (the reference runs one
// rayon task per query, pgr-query.rs:135); the range merging (pgr-query.rs:167-285, query_post.hpp) is host bookkeeping as
// in the reference.  Targets are written in ascending sid order (the reference's order is FxHashMap iteration order).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "query_post.hpp"
#include "seq_index_db.hpp"

using namespace pgrb200;

// Path::file_stem (std): the file name without its last extension; ".name" and names without a dot are kept whole
static std::string file_stem(const std::string &path) {
    const size_t sl = path.find_last_of('/');
    const std::string name = sl == std::string::npos ? path : path.substr(sl + 1);
    const size_t dot = name.find_last_of('.');
    if (dot == std::string::npos || dot == 0) return name;
    return name.substr(0, dot);
}
// Path::with_extension: the extension of the last component, if any, is replaced
static std::string with_extension(const std::string &prefix, const std::string &ext) {
    const size_t sl = prefix.find_last_of('/');
    const size_t start = sl == std::string::npos ? 0 : sl + 1;
    const size_t dot = prefix.find_last_of('.');
    std::string base = prefix;
    if (dot != std::string::npos && dot > start) base = prefix.substr(0, dot);
    return base + "." + ext;
}
// fasta_io.rs:26-44
static void reverse_complement(std::vector<uint8_t> &s) {
    std::vector<uint8_t> o(s.rbegin(), s.rend());
    for (auto &b : o) {
        switch (b) {
            case 'A': b = 'T'; break; case 'C': b = 'G'; break; case 'G': b = 'C'; break; case 'T': b = 'A'; break;
            case 'a': b = 't'; break; case 'c': b = 'g'; break; case 'g': b = 'c'; break; case 't': b = 'a'; break;
            default: break;
        }
    }
    s.swap(o);
}

int main(int argc, char **argv) {
    uint32_t w = 80, k = 56, r = 4, min_span = 64;
    double gap_penalty = 0.025;
    long merge_range_tol = 100000, max_count = 128, max_query_count = 128, max_target_count = 128, max_aln_chain_span = 8;
    bool fastx_file = false, frg_file = false, only_summary = false, bed_summary = false, mdb_resident = false;
    std::vector<std::string> pos;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto val = [&]() -> const char * { if (i + 1 >= argc) { fprintf(stderr, "error: %s needs a value\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "-w") w = (uint32_t)atol(val());
        else if (a == "-k") k = (uint32_t)atol(val());
        else if (a == "-r") r = (uint32_t)atol(val());
        else if (a == "--min-span") min_span = (uint32_t)atol(val());
        else if (a == "--gap-penalty-factor") gap_penalty = atof(val());
        else if (a == "--merge-range-tol") merge_range_tol = atol(val());
        else if (a == "--max-count") max_count = atol(val());
        else if (a == "--max-query-count") max_query_count = atol(val());
        else if (a == "--max-target-count") max_target_count = atol(val());
        else if (a == "--max-aln-chain-span") max_aln_chain_span = atol(val());
        else if (a == "--number-of-thread") (void)val();   // the GPU path has no host thread pool
        else if (a == "--fastx-file") fastx_file = true;
        else if (a == "--frg-file") frg_file = true;
        else if (a == "--only-summary") only_summary = true;
        else if (a == "--mdb-resident") mdb_resident = true;   // addition: leave the .mdb on disk (memory-mapped), as the reference's agc / frg back ends do
        else if (a == "--bed-summary") bed_summary = true;
        else if (a == "-h" || a == "--help") { printf("usage: pgr-b200-query <pgr_db_prefix|fastx> <query_fastx> <output_prefix> [--fastx-file] [options of pgr-query]\n"); return 0; }
        else pos.push_back(a);
    }
    if (pos.size() != 3) { fprintf(stderr, "usage: pgr-b200-query <pgr_db_prefix|fastx> <query_fastx> <output_prefix> [--fastx-file] [options of pgr-query]\n"); return 2; }
    std::vector<SeqRec> queries;
    std::string err;
    if (!read_fastx(pos[1], queries, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }

    SeqIndexDB db;
    int rc;
    if (fastx_file) {
        fprintf(stderr, "the option `--fastx_file` is specified, read the input file as a fastx file.\n");
        db.keep_sequences(!only_summary);
        rc = db.load_from_fastx(pos[0], w, k, r, min_span);
    } else if (frg_file) {
        fprintf(stderr, "the option `--frg_file` is specified, read the input file as a FRG backed index database files.\n");
        rc = only_summary ? db.load_from_index_files(pos[0], mdb_resident) : db.load_from_frg_index(pos[0]);
    } else {
        if (!only_summary) { fprintf(stderr, "error: the AGC back end is out of scope: use --frg-file or --fastx-file, or add --only-summary\n"); return 2; }
        rc = db.load_from_index_files(pos[0], mdb_resident);
    }
    if (rc != PGR_OK) { fprintf(stderr, "%s\n", db.error().c_str()); return 1; }

    pgr_query_params prm;
    prm.penalty = (float)gap_penalty;
    prm.max_count = max_count; prm.max_count_query = max_query_count; prm.max_count_target = max_target_count;
    prm.max_aln_span = max_aln_chain_span; prm.max_gap = -1; prm.oriented = 0;       // pgr-query.rs:144-164
    pgr_query_result *res = nullptr;
    rc = db.query_fragment_to_hps(queries, prm, &res);
    if (rc != PGR_OK) { fprintf(stderr, "%s\n", db.error().c_str()); return 1; }

    for (size_t idx = 0; idx < queries.size(); idx++) {
        const std::string &q_name = queries[idx].id;
        const size_t q_len = queries[idx].seq.size();
        const auto targets = merge_query_hits(*res, idx, merge_range_tol);
        char ext[64];
        snprintf(ext, sizeof ext, bed_summary ? "%03zu.hit.bed" : "%03zu.hit", idx);
        FILE *hit = fopen(with_extension(pos[2], ext).c_str(), "wb");
        if (!hit) { fprintf(stderr, "cannot create the hit file\n"); return 1; }
        FILE *fa = nullptr;
        if (!only_summary) {
            snprintf(ext, sizeof ext, "%03zu.fa", idx);
            fa = fopen(with_extension(pos[2], ext).c_str(), "wb");
            if (!fa) { fprintf(stderr, "cannot create the fasta file\n"); return 1; }
        }
        if (bed_summary) fprintf(hit, "#target\tbgn\tend\tquery\tcolor\torientation\tq_len\taln_anchor_count\tq_idx\tsrc\tctg_bgn\tctg_end\n");
        else fprintf(hit, "#idx\tq_ctg_name\tq_ctg_bgn\tq_ctg_end\tq_ctg_len\taln_anchor_count\tsrc\tctg\tctg_bgn\tctg_end\torientation\tctg_name\n");
        for (const auto &t : targets) {
            if (t.sid >= db.seqs().size()) { fprintf(stderr, "target sid %u has no .midx record\n", t.sid); return 1; }
            const CompactSeq &cs = db.seqs()[t.sid];
            const std::string src = cs.source.empty() ? "N/A" : cs.source;
            for (auto rg : t.regions) {
                std::sort(rg.aln.begin(), rg.aln.end());
                const uint32_t q_bgn = rg.aln.front().qb, q_end = rg.aln.back().qe;
                char nm[4096];
                snprintf(nm, sizeof nm, "%s::%s_%u_%u_%u", file_stem(src).c_str(), cs.name.c_str(), rg.bgn, rg.end, rg.orientation);
                if (bed_summary)
                    fprintf(hit, "%s\t%u\t%u\t%s\t#AAAAAA\t%u\t%zu\t%zu\t%zu\t%s\t%u\t%u\t%s\n", cs.name.c_str(), rg.bgn, rg.end, q_name.c_str(), rg.orientation,
                            q_len, rg.aln.size(), idx, src.c_str(), q_bgn, q_end, nm);
                else
                    fprintf(hit, "%03zu\t%s\t%u\t%u\t%zu\t%zu\t%s\t%s\t%u\t%u\t%u\t%s\n", idx, q_name.c_str(), q_bgn, q_end, q_len, rg.aln.size(), src.c_str(),
                            cs.name.c_str(), rg.bgn, rg.end, rg.orientation, nm);
                if (fa) {
                    std::vector<uint8_t> sub;
                    if (!db.get_sub_seq_by_id(t.sid, rg.bgn, rg.end, sub)) { fprintf(stderr, "cannot fetch the target sub-sequence\n"); return 1; }
                    if (rg.orientation == 1) reverse_complement(sub);
                    fprintf(fa, ">%s\n", nm);
                    fwrite(sub.data(), 1, sub.size(), fa);
                    fputc('\n', fa);
                }
            }
        }
        fclose(hit);
        if (fa) fclose(fa);
    }
    pgr_b200_query_result_free(res);
    return 0;
}
