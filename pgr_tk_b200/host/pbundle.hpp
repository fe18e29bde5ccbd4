// pbundle.hpp — principal-bundle bookkeeping above the C ABI (host-side in the reference too):
//   SeqIndexDB::get_principal_bundles_with_id      ext.rs:552-650  (bundle order / direction by consensus voting)
//   get_principal_bundle_decomposition             ext.rs:976-1015
//   group_smps_by_principle_bundle_id              pgr-bin/src/bin/pgr-pbundle-decomp.rs:62-137
// The shimmer pairs of every sequence (SeqIndexDB::get_smps, ext.rs:533-550) come from the GPU (pgr_b200_raw_query's
// pair list: same strict-'<' canonical form), the bundles from pgr_b200_adj_list + pgr_b200_principal_bundles.
// Where the reference iterates seq_info (an FxHashMap) the f32 sums it accumulates are exact for any order as long as they
// stay below 2^24 (orders are pair ordinals), so ascending sid is used.
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <string>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/pgr_b200.h"

namespace pgrb200 {

struct Smp { uint64_t h0, h1; uint32_t bgn, end; uint8_t ori; };                 // (u64, u64, u32, u32, u8), ext.rs:533
struct BundleVertex { uint64_t h0, h1; uint8_t ori; };
struct BundleRef { size_t bundle_id; uint8_t direction; size_t pos; };            // VertexToBundleIdMap value, ext.rs:30
struct BundleWithId { size_t bundle_id, mean_order; std::vector<BundleVertex> vertices; };   // ext.rs:27

struct KeyHash { size_t operator()(const std::pair<uint64_t, uint64_t> &k) const { return (size_t)(k.first * 0x9E3779B97F4A7C15ull ^ k.second); } };
using VertexMap = std::unordered_map<std::pair<uint64_t, uint64_t>, BundleRef, KeyHash>;

// ext.rs:552-650; pb = get_principal_bundles(), smps[i] = shimmer pairs of the i-th sequence
inline void principal_bundles_with_id(const std::vector<std::vector<BundleVertex>> &pb, const std::vector<std::vector<Smp>> &smps,
                                      std::vector<BundleWithId> &out, VertexMap &vmap) {
    vmap.clear();
    for (size_t b = 0; b < pb.size(); b++)
        for (size_t p = 0; p < pb[b].size(); p++) vmap[{pb[b][p].h0, pb[b][p].h1}] = {b, pb[b][p].ori, p};   // later entries win, as collect() does
    std::vector<std::vector<uint32_t>> directions(pb.size());
    std::vector<std::vector<float>> orders(pb.size());
    for (const auto &sm : smps) {
        std::unordered_set<size_t> visited;
        for (size_t order = 0; order < sm.size(); order++) {
            auto it = vmap.find({sm[order].h0, sm[order].h1});
            if (it == vmap.end()) continue;
            const BundleRef &bid = it->second;
            if (visited.insert(bid.bundle_id).second) orders[bid.bundle_id].push_back((float)order);
            directions[bid.bundle_id].push_back(bid.direction == sm[order].ori ? 0u : 1u);
        }
    }
    std::vector<std::tuple<size_t, size_t, uint8_t>> mod;   // (mean_ord, bid, direction)
    for (size_t bid = 0; bid < pb.size(); bid++) {
        if (!orders[bid].empty()) {
            float sum = 0.0f;
            for (float o : orders[bid]) sum += o;
            const float mean = sum / (float)orders[bid].size();
            size_t dir_sum = 0;
            for (uint32_t d : directions[bid]) dir_sum += d;
            mod.emplace_back((size_t)mean, bid, dir_sum < (directions[bid].size() >> 1) ? (uint8_t)0 : (uint8_t)1);
        } else {
            mod.emplace_back(SIZE_MAX, bid, (uint8_t)0);
        }
    }
    std::sort(mod.begin(), mod.end());
    out.clear();
    for (const auto &[ord, bid, direction] : mod) {
        BundleWithId bw{bid, ord, {}};
        if (direction == 1) {
            for (auto it = pb[bid].rbegin(); it != pb[bid].rend(); ++it) bw.vertices.push_back({it->h0, it->h1, (uint8_t)(1 - it->ori)});
            for (size_t p = 0; p < bw.vertices.size(); p++) vmap[{bw.vertices[p].h0, bw.vertices[p].h1}] = {bid, bw.vertices[p].ori, p};
        } else {
            bw.vertices = pb[bid];
        }
        out.push_back(std::move(bw));
    }
}

// one element of a partition: (smp, bundle id, direction, position in the bundle) — pgr-pbundle-decomp.rs:66
struct SmpInBundle { Smp smp; size_t bid; uint32_t d; size_t bpos; };

// pgr-pbundle-decomp.rs:62-137 on one sequence's decomposition (ext.rs:976-1015: each smp with its optional bundle vertex)
inline std::vector<std::vector<SmpInBundle>> group_smps_by_principle_bundle_id(const std::vector<Smp> &smps, const VertexMap &vmap,
                                                                               size_t bundle_length_cutoff, size_t bundle_merge_distance) {
    std::vector<std::vector<SmpInBundle>> all, rtn;
    std::vector<SmpInBundle> cur;
    bool have_pre = false;
    size_t pre_bid = 0;
    uint32_t pre_d = 0;
    auto long_enough = [&](const std::vector<SmpInBundle> &p) { return (size_t)p.back().smp.end - (size_t)p.front().smp.bgn > bundle_length_cutoff; };
    for (const Smp &s : smps) {
        auto it = vmap.find({s.h0, s.h1});
        if (it == vmap.end()) continue;
        const uint32_t d = s.ori == it->second.direction ? 0u : 1u;
        const size_t bid = it->second.bundle_id, bpos = it->second.pos;
        if (!have_pre) {
            cur.clear();
            cur.push_back({s, bid, d, bpos});
            have_pre = true; pre_bid = bid; pre_d = d;
            continue;
        }
        if (bid != pre_bid || d != pre_d) {
            if (long_enough(cur)) all.push_back(cur);
            cur.clear();
            pre_bid = bid; pre_d = d;
        }
        cur.push_back({s, bid, d, bpos});
    }
    if (!cur.empty() && long_enough(cur)) all.push_back(cur);
    if (all.empty()) return rtn;
    std::vector<SmpInBundle> part = all[0];
    for (size_t i = 1; i < all.size(); i++) {
        const auto &p = all[i];
        const SmpInBundle &l = part.back();
        const int64_t gap = (int64_t)p[0].smp.bgn - (int64_t)l.smp.end;
        if (l.bid == p[0].bid && l.d == p[0].d && (gap < 0 ? -gap : gap) < (int64_t)bundle_merge_distance) part.insert(part.end(), p.begin(), p.end());
        else { rtn.push_back(part); part = p; }
    }
    if (!part.empty()) rtn.push_back(part);
    return rtn;
}

}  // namespace pgrb200
