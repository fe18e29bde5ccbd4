// pgr-b200-pbundle-decomp — pgr-pbundle-decomp (pgr-bin/src/bin/pgr-pbundle-decomp.rs:17-58,139-529) on the B200:
//   pgr-b200-pbundle-decomp <fastx_path> <output_prefix> [-d <decomp_fastx_path>] [-w 48] [-k 56] [-r 4] [--min-span 12]
//       [--min-cov 0] [--min-branch-size 8] [--bundle-length-cutoff 2500] [--bundle-merge-distance 10000]
// Shimmers, ShmmrFragMap, MAP-graph adjacency list and vertex weights come from the GPU; the bundle walk, the consensus
// ordering, the per-contig decomposition and the writers are host bookkeeping as in the reference.  Written:
// <prefix>.mapg.gfa, <prefix>.mapg.idx, <prefix>.pmapg.gfa, <prefix>.pdb, <prefix>.bed and <prefix>.ctg.summary.tsv; with
// -p <file.pdb> (--precomputed-bundles) the bundles are read instead of computed and only the .bed / summary are written
// (pgr-pbundle-decomp.rs:158-226); with -i <file> (--include) only the listed contigs are decomposed (:278-301).
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "mapg_gfa.hpp"
#include "pdb_io.hpp"
#include "pbundle.hpp"
#include "seq_index_db.hpp"

using namespace pgrb200;

static std::string with_extension(const std::string &prefix, const std::string &ext) {   // Path::with_extension
    const size_t sl = prefix.find_last_of('/');
    const size_t start = sl == std::string::npos ? 0 : sl + 1;
    const size_t dot = prefix.find_last_of('.');
    std::string base = prefix;
    if (dot != std::string::npos && dot > start) base = prefix.substr(0, dot);
    return base + "." + ext;
}
// Rust `{}` of an f32: shortest digits that round-trip, never in exponent form
static std::string f32_display(float v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[512];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

// SeqIndexDB::get_smps (ext.rs:533-550) for every sequence of a database, on the GPU
static int all_smps(SeqIndexDB &db, std::vector<std::vector<Smp>> &out) {
    out.assign(db.seqs().size(), {});
    std::vector<uint8_t> seq;
    for (const auto &cs : db.seqs()) {
        if (!db.get_sub_seq_by_id(cs.id, 0, cs.len, seq)) return PGR_E_ARG;
        pgr_query_pair *pairs = nullptr;
        size_t n = 0;
        uint64_t *hit_off = nullptr;
        pgr_frag_sig *hits = nullptr;
        const int rc = pgr_b200_raw_query(db.index(), seq.data(), seq.size(), &pairs, &n, &hit_off, &hits);
        if (rc != PGR_OK) return rc;
        auto &v = out[cs.id];
        v.reserve(n);
        for (size_t i = 0; i < n; i++) v.push_back({pairs[i].h0, pairs[i].h1, pairs[i].bgn, pairs[i].end, pairs[i].ori});
        pgr_b200_free(pairs); pgr_b200_free(hit_off); pgr_b200_free(hits);
    }
    return PGR_OK;
}

int main(int argc, char **argv) {
    uint32_t w = 48, k = 56, r = 4, min_span = 12;
    size_t min_cov = 0, min_branch_size = 8, bundle_length_cutoff = 2500, bundle_merge_distance = 10000;
    std::string decomp_path, precomputed, include_path;
    std::vector<std::string> pos;
    std::string cmd_string;
    for (int i = 0; i < argc; i++) { if (i) cmd_string += " "; cmd_string += argv[i]; }
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto val = [&]() -> const char * { if (i + 1 >= argc) { fprintf(stderr, "error: %s needs a value\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "-w") w = (uint32_t)atol(val());
        else if (a == "-k") k = (uint32_t)atol(val());
        else if (a == "-r") r = (uint32_t)atol(val());
        else if (a == "--min-span") min_span = (uint32_t)atol(val());
        else if (a == "--min-cov") min_cov = (size_t)atol(val());
        else if (a == "--min-branch-size") min_branch_size = (size_t)atol(val());
        else if (a == "--bundle-length-cutoff") bundle_length_cutoff = (size_t)atol(val());
        else if (a == "--bundle-merge-distance") bundle_merge_distance = (size_t)atol(val());
        else if (a == "-d" || a == "--decomp-fastx-path") decomp_path = val();
        else if (a == "-p" || a == "--precomputed-bundles") precomputed = val();
        else if (a == "-i" || a == "--include") include_path = val();
        else if (a == "-h" || a == "--help") { printf("usage: pgr-b200-pbundle-decomp <fastx_path> <output_prefix> [options of pgr-pbundle-decomp]\n"); return 0; }
        else pos.push_back(a);
    }
    if (pos.size() != 2) { fprintf(stderr, "usage: pgr-b200-pbundle-decomp <fastx_path> <output_prefix> [options of pgr-pbundle-decomp]\n"); return 2; }

    // ---- principal bundles: computed from <fastx_path> (pgr-pbundle-decomp.rs:228-246, ext.rs:491-510,552-650) or read from a
    // .pdb written earlier (:158-226; its parameters override the command line) ----
    SeqIndexDB db;
    db.keep_sequences(true);
    std::vector<std::vector<Smp>> smps;
    std::vector<BundleWithId> pbid;
    VertexMap vmap;
    if (precomputed.empty()) {
        if (db.load_from_fastx(pos[0], w, k, r, min_span) != PGR_OK) { fprintf(stderr, "can't read file %s (%s)\n", pos[0].c_str(), db.error().c_str()); return 1; }
        std::vector<std::vector<BundleVertex>> pb;
        std::vector<pgr_adj_pair> filtered_adj;
        {
            pgr_adj_pair *adj = nullptr;
            size_t n_adj = 0;
            if (pgr_b200_adj_list(db.index(), min_cov, nullptr, 0, 0, &adj, &n_adj) != PGR_OK) { fprintf(stderr, "%s\n", pgr_b200_last_error()); return 1; }
            if (n_adj) {
                pgr_graph_node *verts = nullptr;
                uint64_t *off = nullptr;
                pgr_adj_pair *flt = nullptr;
                size_t nb = 0, nf = 0;
                if (pgr_b200_principal_bundles(db.index(), adj, n_adj, min_branch_size, &verts, &off, &nb, &flt, &nf) != PGR_OK) { fprintf(stderr, "%s\n", pgr_b200_last_error()); return 1; }
                pb.resize(nb);
                for (size_t b = 0; b < nb; b++) for (uint64_t i = off[b]; i < off[b + 1]; i++) pb[b].push_back({verts[i].h0, verts[i].h1, verts[i].ori});
                filtered_adj.assign(flt, flt + nf);
                pgr_b200_free(verts); pgr_b200_free(off); pgr_b200_free(flt);
            }
            pgr_b200_free(adj);
        }
        // ---- MAP-graph files (pgr-pbundle-decomp.rs:304-331): the whole graph (min_count 0), its index, the principal graph ----
        {
            IndexCsr csr;
            if (export_csr(db.index(), csr) != PGR_OK) { fprintf(stderr, "%s\n", pgr_b200_last_error()); return 1; }
            pgr_adj_pair *adj0 = nullptr;
            size_t n0 = 0;
            if (pgr_b200_adj_list(db.index(), 0, nullptr, 0, 0, &adj0, &n0) != PGR_OK) { fprintf(stderr, "%s\n", pgr_b200_last_error()); return 1; }
            pgr_shmmr_spec sp{w, k, r, min_span, 0};
            VertexMap plain;                                                      // get_vertex_map_from_principal_bundles (ext.rs:512-531)
            for (size_t b = 0; b < pb.size(); b++) for (size_t p = 0; p < pb[b].size(); p++) plain[{pb[b][p].h0, pb[b][p].h1}] = {b, pb[b][p].ori, p};
            const bool ok = write_gfa(with_extension(pos[1], "mapg.gfa"), adj0, n0, csr, k, nullptr) &&
                            write_mapg_idx(with_extension(pos[1], "mapg.idx"), sp, db.seqs(), csr) &&
                            write_gfa(with_extension(pos[1], "pmapg.gfa"), filtered_adj.data(), filtered_adj.size(), csr, k, &plain);
            pgr_b200_free(adj0);
            if (!ok) { fprintf(stderr, "cannot write the MAP-graph files\n"); return 1; }
        }
        if (all_smps(db, smps) != PGR_OK) { fprintf(stderr, "%s\n", pgr_b200_last_error()); return 1; }
        principal_bundles_with_id(pb, smps, pbid, vmap);
        PdbData pd;                                                           // pgr-pbundle-decomp.rs:367-396
        pd.w = w; pd.k = k; pd.r = r; pd.min_span = min_span; pd.min_branch_size = min_branch_size; pd.min_cov = min_cov;
        pd.bundles = pbid; pd.vmap = vmap;
        if (!write_pdb(with_extension(pos[1], "pdb"), pd)) { fprintf(stderr, "pdb file creating error\n"); return 1; }
    } else {
        PdbData pd;
        if (!read_pdb(precomputed, pd)) { fprintf(stderr, "pdb input file open / reading error\n"); return 1; }
        w = pd.w; k = pd.k; r = pd.r; min_span = pd.min_span; min_branch_size = (size_t)pd.min_branch_size; min_cov = (size_t)pd.min_cov;
        pbid = std::move(pd.bundles); vmap = std::move(pd.vmap);
    }

    // ---- the database that is decomposed (pgr-pbundle-decomp.rs:257-276) ----
    SeqIndexDB ddb_other;
    SeqIndexDB *ddb = &db;
    if (!decomp_path.empty() || !precomputed.empty()) {
        const std::string dpath = decomp_path.empty() ? pos[0] : decomp_path;
        ddb_other.keep_sequences(true);
        if (ddb_other.load_from_fastx(dpath, w, k, r, min_span) != PGR_OK) { fprintf(stderr, "can't read file %s\n", dpath.c_str()); return 1; }
        ddb = &ddb_other;
        if (all_smps(*ddb, smps) != PGR_OK) { fprintf(stderr, "%s\n", pgr_b200_last_error()); return 1; }
    }
    SeqIndexDB ddb_included;
    if (!include_path.empty()) {                                              // pgr-pbundle-decomp.rs:278-301
        FILE *inc = fopen(include_path.c_str(), "rb");
        if (!inc) { fprintf(stderr, "can't open the include file\n"); return 1; }
        std::vector<std::string> names;
        char *line = nullptr;
        size_t cap = 0;
        ssize_t n;
        while ((n = getline(&line, &cap, inc)) > 0) {
            std::string l(line, (size_t)n);
            while (!l.empty() && (l.back() == '\n' || l.back() == '\r')) l.pop_back();
            if (std::find(names.begin(), names.end(), l) == names.end()) names.push_back(l);   // a set in the reference; file order here
        }
        free(line);
        fclose(inc);
        std::vector<SeqRec> list;
        for (const auto &nm : names) {
            const CompactSeq *hit = nullptr;
            for (const auto &cs : ddb->seqs()) if (cs.name == nm) { hit = &cs; break; }
            SeqRec rec;
            rec.id = nm;
            if (!hit || !ddb->get_sub_seq_by_id(hit->id, 0, hit->len, rec.seq)) { fprintf(stderr, "fail to fetch sequence %s\n", nm.c_str()); return 1; }
            list.push_back(std::move(rec));
        }
        const std::string src = decomp_path.empty() ? pos[0] : decomp_path;
        ddb_included.keep_sequences(true);
        if (ddb_included.load_from_seq_list(list, src, w, k, r, min_span) != PGR_OK) { fprintf(stderr, "%s\n", ddb_included.error().c_str()); return 1; }
        ddb = &ddb_included;
        if (all_smps(*ddb, smps) != PGR_OK) { fprintf(stderr, "%s\n", pgr_b200_last_error()); return 1; }
    }

    FILE *bed = fopen(with_extension(pos[1], "bed").c_str(), "wb");
    FILE *summary = fopen(with_extension(pos[1], "ctg.summary.tsv").c_str(), "wb");
    if (!bed || !summary) { fprintf(stderr, "cannot create the output files\n"); return 1; }
    fprintf(bed, "# cmd: %s\n", cmd_string.c_str());

    std::unordered_map<size_t, size_t> bid_to_size;
    for (const auto &b : pbid) bid_to_size[b.bundle_id] = b.vertices.size();
    std::vector<const CompactSeq *> order;                                    // seq_info.sort_by_key(ctg name), stable
    for (const auto &cs : ddb->seqs()) order.push_back(&cs);
    std::stable_sort(order.begin(), order.end(), [](const CompactSeq *a, const CompactSeq *b) { return a->name < b->name; });
    std::unordered_map<uint32_t, std::vector<uint32_t>> repeat_count, non_repeat_count;
    for (const CompactSeq *cs : order) {
        const auto parts = group_smps_by_principle_bundle_id(smps[cs->id], vmap, bundle_length_cutoff, bundle_merge_distance);
        std::unordered_map<size_t, size_t> ctg_bundle_count;
        for (const auto &p : parts) ctg_bundle_count[p[0].bid]++;
        for (const auto &p : parts) {
            const uint32_t b = p.front().smp.bgn, e = p.back().smp.end + k;
            const size_t bid = p[0].bid;
            const bool is_repeat = ctg_bundle_count[bid] > 1;
            (is_repeat ? repeat_count : non_repeat_count)[cs->id].push_back(e - b - k);
            fprintf(bed, "%s\t%u\t%u\t%zu:%zu:%u:%zu:%zu:%s\n", cs->name.c_str(), b, e, bid, bid_to_size[bid], p[0].d, p.front().bpos, p.back().bpos,
                    is_repeat ? "R" : "U");
        }
    }
    fprintf(summary, "#ctg\tlength\trepeat_bundle_count\trepeat_bundle_sum\trepeat_bundle_percentage\trepeat_bundle_mean\trepeat_bundle_min\trepeat_bundle_max\t"
                     "non_repeat_bundle_count\tnon_repeat_bundle_sum\tnon_repeat_bundle_percentage\tnon_repeat_bundle_mean\tnon_repeat_bundle_min\tnon_repeat_bundle_max\t"
                     "total_bundle_count\ttotal_bundle_coverage_percentage\n");
    for (const CompactSeq *cs : order) {
        const uint32_t len = (uint32_t)cs->len;
        auto stats = [&](const std::vector<uint32_t> &v, uint32_t &sum, std::string &mean, std::string &mn, std::string &mx) {
            sum = 0;
            uint32_t lo = len, hi = 0;
            for (uint32_t x : v) { sum += x; lo = x < lo ? x : lo; hi = x > hi ? x : hi; }
            if (v.empty()) { mean = mn = mx = "NA"; return; }
            mean = f32_display((float)sum / (float)v.size()); mn = std::to_string(lo); mx = std::to_string(hi);
        };
        uint32_t rs, ns;
        std::string rmean, rmin, rmax, nmean, nmin, nmax;
        const auto &rv = repeat_count[cs->id], &nv = non_repeat_count[cs->id];
        stats(rv, rs, rmean, rmin, rmax);
        stats(nv, ns, nmean, nmin, nmax);
        fprintf(summary, "%s\t%u\t%zu\t%u\t%s\t%s\t%s\t%s\t%zu\t%u\t%s\t%s\t%s\t%s\t%zu\t%s\n", cs->name.c_str(), len, rv.size(), rs,
                f32_display(100.0f * (float)rs / (float)len).c_str(), rmean.c_str(), rmin.c_str(), rmax.c_str(), nv.size(), ns,
                f32_display(100.0f * (float)ns / (float)len).c_str(), nmean.c_str(), nmin.c_str(), nmax.c_str(), rv.size() + nv.size(),
                f32_display(100.0f * (float)(rs + ns) / (float)len).c_str());
    }
    fclose(bed);
    fclose(summary);
    return 0;
}
