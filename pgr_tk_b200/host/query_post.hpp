// query_post.hpp — pgr-query's post-processing of query_fragment_to_hps results (pgr-bin/src/bin/pgr-query.rs:167-285):
// alignment chains with more than two anchors -> per-target ranges -> forward / reverse ranges merged when they lie
// closer than --merge-range-tol.  Host-side bookkeeping in the reference too; the chains come from pgr_b200_query_batch.
//
// Canonical form where the reference leaks FxHashMap iteration order (pgr-query.rs:167,187,200: which target comes
// first): targets ascending by sid.  Everything else is deterministic and reproduced as written, including the
// orientation counters that are NOT reset between the chains of one target (pgr-query.rs:170-171).
#pragma once
#include <algorithm>
#include <cstdint>
#include <tuple>
#include <vector>

#include "../../include/pgr_b200.h"

namespace pgrb200 {

// aln.rs:10 HitPair ((q_bgn,q_end,q_ori),(t_bgn,t_end,t_ori)) with the tuple ordering Rust derives
struct HitPair {
    uint32_t qb, qe; uint8_t qo; uint32_t tb, te; uint8_t to;
    std::tuple<uint32_t, uint32_t, uint8_t, uint32_t, uint32_t, uint8_t> key() const { return {qb, qe, qo, tb, te, to}; }
    bool operator<(const HitPair &o) const { return key() < o.key(); }
    bool operator==(const HitPair &o) const { return key() == o.key(); }
};

// (bgn, end, len, orientation, aln) — pgr-query.rs:198
struct Region {
    uint32_t bgn = 0, end = 0, len = 0, orientation = 0;
    std::vector<HitPair> aln;
    bool operator<(const Region &o) const {
        return std::tie(bgn, end, len, orientation, aln) < std::tie(o.bgn, o.end, o.len, o.orientation, o.aln);
    }
};

struct TargetRegions { uint32_t sid; std::vector<Region> regions; };

// pgr-query.rs:219-245 (the same block is written twice, for the forward and for the reverse regions)
inline void merge_sorted_regions(std::vector<Region> rgns, int64_t merge_range_tol, std::vector<Region> &out) {
    Region last;   // (0, 0, 0, 0, vec![])
    for (auto &r : rgns) {
        if (last.aln.empty()) { last = std::move(r); continue; }
        if ((int64_t)r.bgn - (int64_t)last.end < merge_range_tol) {
            last.end = r.end > last.end ? r.end : last.end;
            last.len = last.end - last.bgn;
            last.aln.insert(last.aln.end(), r.aln.begin(), r.aln.end());
        } else {
            out.push_back(last);
            last = std::move(r);
        }
    }
    if (last.len > 0) out.push_back(std::move(last));
}

// One query of a pgr_query_result -> merged regions per target (pgr-query.rs:166-285)
inline std::vector<TargetRegions> merge_query_hits(const pgr_query_result &r, size_t q, int64_t merge_range_tol) {
    std::vector<TargetRegions> out;
    for (uint64_t t = r.q_target_off[q]; t < r.q_target_off[q + 1]; t++) {     // ascending sid (canonical)
        size_t f_count = 0, r_count = 0;                                       // per target, not per chain
        std::vector<Region> rgns;
        for (uint64_t c = r.target_chain_off[t]; c < r.target_chain_off[t + 1]; c++) {
            const uint64_t h0 = r.chain_hit_off[c], h1 = r.chain_hit_off[c + 1];
            if (h1 - h0 <= 2) continue;                                        // aln.len() > 2
            Region g;
            for (uint64_t h = h0; h < h1; h++) {
                const pgr_hit_pair &p = r.hits[h];
                g.aln.push_back({p.qb, p.qe, p.qo, p.tb, p.te, p.to});
                if (p.qo == p.to) f_count++; else r_count++;
            }
            g.orientation = f_count > r_count ? 0u : 1u;
            // range = smallest (t_bgn, t_end) .. end of the largest (pgr-query.rs:189-196)
            std::pair<uint32_t, uint32_t> lo{g.aln[0].tb, g.aln[0].te}, hi = lo;
            for (const auto &hp : g.aln) {
                const std::pair<uint32_t, uint32_t> v{hp.tb, hp.te};
                if (v < lo) lo = v;
                if (hi < v) hi = v;
            }
            g.bgn = lo.first; g.end = hi.second; g.len = g.end - g.bgn;
            rgns.push_back(std::move(g));
        }
        if (rgns.empty()) continue;
        std::vector<Region> f, rv;
        for (auto &g : rgns) (g.orientation == 0 ? f : rv).push_back(g);
        std::sort(f.begin(), f.end());
        std::sort(rv.begin(), rv.end());
        TargetRegions tr;
        tr.sid = r.target_sid[t];
        merge_sorted_regions(std::move(f), merge_range_tol, tr.regions);
        merge_sorted_regions(std::move(rv), merge_range_tol, tr.regions);
        out.push_back(std::move(tr));
    }
    return out;
}

}  // namespace pgrb200
