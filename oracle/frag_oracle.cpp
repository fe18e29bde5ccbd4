// frag_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY) for the reference's fragment compression, in C++ so that it can
// run at benchmark sizes and serve as the CPU baseline beside the GPU path:
//   shmmrutils.rs:35-54   track_delta_point
//   shmmrutils.rs:57-223  match_reads (banded O(nD) variant)
//   seq_db.rs:113-156     deltas_to_aln_segs
//   seq_db.rs:189-358     CompactSeqDB::seq_to_compressed (sequences in order; the pairs of one sequence in parallel like the
//                         reference's par_iter, inserts single-threaded)
// It is the same restatement as oracle/frag_oracle.py (which tests/test_frag_format.py pins to the reference's own
// test_seqs_frag.frg fixture) and tests/test_frag_format.py checks the two against each other.  Output records have the
// layout of pgr_fragment / pgr_aln_seg (include/pgr_b200.h) so that the GPU result can be compared array to array.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <vector>

#include "pgr_oracle.h"

namespace {

struct Delta { uint32_t x, y; int dk; };
struct Match { uint32_t end0, end1; std::vector<Delta> deltas; };

// shmmrutils.rs:57-223 with get_delta = true, tol = 0.1, min_match_len = 0, min_match_start = 0
bool match_reads(const uint8_t *s0, uint32_t len0, const uint8_t *s1, uint32_t len1, uint32_t bandwidth, Match &m) {
    const uint32_t d_max = 32 + (uint32_t)(0.1 * (double)(len0 < len1 ? len0 : len1));
    std::vector<uint32_t> U(2 * d_max + 3, 0), V(2 * d_max + 3, 0);
    const int kb = (int)d_max + 1;
    struct Row { int k_min; std::vector<Delta> pts; };
    std::vector<Row> rows;
    int k_min = 0, k_max = 0, best_m = -1;
    bool matched = false;
    uint32_t d_final = 0;
    int k_final = 0;
    for (uint32_t d = 0; d < d_max; d++) {
        if (k_max - k_min > (int)bandwidth) break;
        rows.push_back({k_min, {}});
        for (int k = k_min; k <= k_max; k += 2) {
            const uint32_t vn = V[k - 1 + kb], vp = V[k + 1 + kb];
            uint32_t x;
            int pre_k;
            if (k == k_min || (k != k_max && vn < vp)) { x = vp; pre_k = k + 1; } else { x = vn + 1; pre_k = k - 1; }
            uint32_t y = (uint32_t)((int)x - k);
            rows.back().pts.push_back({x, y, k - pre_k});
            while (x < len0 && y < len1 && s0[x] == s1[y]) { x++; y++; }
            U[k + kb] = x + y; V[k + kb] = x;
            if ((int)(x + y) > best_m) best_m = (int)(x + y);
            if (x >= len0 || y >= len1) { matched = true; d_final = d; k_final = k; m.end0 = x; m.end1 = y; break; }
        }
        int k_max_new = k_min, k_min_new = k_max;
        for (int k2 = k_min; k2 <= k_max; k2 += 2)
            if ((int)U[k2 + kb] >= best_m - (int)bandwidth) { k_min_new = std::min(k_min_new, k2); k_max_new = std::max(k_max_new, k2); }
        k_max = k_max_new + 1; k_min = k_min_new - 1;
        if (matched) break;
    }
    if (!matched) return false;
    m.deltas.clear();
    uint32_t d = d_final;
    int k = k_final;
    while (d > 0) {                                        // track_delta_point, s = bgn0 = 0, e = end0
        const Row &r = rows[d];
        const Delta &p = r.pts[(size_t)((k - r.k_min) >> 1)];
        if (p.x <= m.end0) m.deltas.push_back(p);
        d -= 1;
        k -= p.dk;
    }
    return true;
}

struct Seg { uint32_t type, a, b; };

// seq_db.rs:113-156
void deltas_to_aln_segs(const Match &m, uint32_t base_len, const uint8_t *frg, uint32_t frg_len, std::vector<Seg> &segs) {
    segs.clear();
    if (m.deltas.empty() && base_len == frg_len) { segs.push_back({0, 0, 0}); return; }
    uint32_t x = m.end0, y = m.end1;
    for (uint32_t yy = frg_len; yy > y; yy--) segs.push_back({2, frg[yy - 1], 0});
    for (const Delta &d : m.deltas) {
        if (d.x < x) segs.push_back({1, d.x, x});
        x = d.x; y = d.y;
        if (d.dk > 0) x -= (uint32_t)d.dk;
        else for (int yy = 0; yy < -d.dk; yy++) segs.push_back({2, frg[y - (uint32_t)yy - 1], 0});
    }
    if (x != 0) segs.push_back({1, 0, x});
    std::reverse(segs.begin(), segs.end());
}

uint8_t rc_base(uint8_t b) {
    switch (b) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
        default: return b;
    }
}

struct Key { uint64_t h0, h1; bool operator==(const Key &o) const { return h0 == o.h0 && h1 == o.h1; } };
struct KeyHash { size_t operator()(const Key &k) const { return (size_t)(k.h0 * 0x9E3779B97F4A7C15ull ^ k.h1); } };
struct Entry { uint32_t frg_id, sid, bgn, end; uint8_t ori; };

}  // namespace

extern "C" {

typedef struct { uint8_t kind, reversed, pad_[2]; uint32_t sid, bgn, end, len, ref_frag, n_segs, pad2_; uint64_t seg_off; } orc_fragment;
typedef struct { uint32_t type, a, b; } orc_alnseg;

// CompactSeqDB::load_seqs_from_seq_vec + seq_to_compressed (try_compress = true) over the sequences in the order given
int orc_compress_fragments(const orc_spec *spec, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens, int nthreads,
                           orc_fragment **frags_out, size_t *n_frags, orc_alnseg **segs_out, size_t *n_segs) {
    const uint32_t k = spec->k;
    std::vector<orc_fragment> frags;
    std::vector<orc_alnseg> segs;
    std::unordered_map<Key, std::vector<Entry>, KeyHash> frag_map;
    std::unordered_map<uint32_t, size_t> seq_of_sid;
    for (size_t q = 0; q < n; q++) seq_of_sid[sids[q]] = q;
    auto push = [&](uint8_t kind, uint32_t sid, uint32_t bgn, uint32_t end) -> orc_fragment & {
        orc_fragment f;
        memset(&f, 0, sizeof f);
        f.kind = kind; f.sid = sid; f.bgn = bgn; f.end = end; f.len = end - bgn;
        frags.push_back(f);
        return frags.back();
    };
    for (size_t si = 0; si < n; si++) {
        const uint32_t sid = sids[si];
        const uint8_t *seq = seqs[si];
        const uint32_t L = (uint32_t)lens[si];
        orc_mm128 *mm = nullptr;
        size_t nm = 0;
        if (orc_sequence_to_shmmrs(sid, seq, L, spec, 0, &mm, &nm) != 0) return -1;
        if (nm == 0) { push(1, sid, 0, L); push(3, sid, L, L); orc_free(mm); continue; }
        auto pos = [&](size_t i) { return (uint32_t)(mm[i].y & 0xFFFFFFFFu) >> 1; };
        push(1, sid, 0, pos(0) + 1);
        const size_t np = nm - 1;
        struct Out { Key key; uint8_t ori; uint32_t bgn, end; bool aligned, rc; uint32_t ref; std::vector<Seg> segs; };
        std::vector<Out> outs(np);
        auto work = [&](size_t a, size_t b) {
            std::vector<uint8_t> frg;
            Match m;
            for (size_t p = a; p < b; p++) {
                const uint64_t s0 = mm[p].x >> 8, s1 = mm[p + 1].x >> 8;
                Out &o = outs[p];
                o.key = s0 <= s1 ? Key{s0, s1} : Key{s1, s0};
                o.ori = s0 <= s1 ? 0 : 1;
                o.bgn = pos(p) + 1; o.end = pos(p + 1) + 1;
                o.aligned = false;
                if (o.end - o.bgn <= 128) continue;
                auto it = frag_map.find(o.key);
                if (it == frag_map.end()) continue;
                for (const Entry &t : it->second) {
                    const orc_fragment &base = frags[t.frg_id];
                    if (base.kind != 2) continue;
                    const uint32_t flen = o.end - o.bgn + k;
                    const bool rc = o.ori != t.ori;
                    frg.assign(seq + (o.bgn - k), seq + o.end);
                    if (rc) { std::reverse(frg.begin(), frg.end()); for (auto &c : frg) c = rc_base(c); }
                    const uint8_t *bseq = seqs[seq_of_sid.at(base.sid)];
                    if (!match_reads(bseq + base.bgn, base.len, frg.data(), flen, 32, m)) continue;
                    deltas_to_aln_segs(m, base.len, frg.data(), flen, o.segs);
                    o.aligned = true; o.rc = rc; o.ref = t.frg_id;
                    break;
                }
            }
        };
        const int nt = std::max(1, std::min<int>(nthreads, (int)np));
        if (nt == 1) work(0, np);
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++) th.emplace_back(work, np * t / nt, np * (t + 1) / nt);
            for (auto &t : th) t.join();
        }
        for (size_t p = 0; p < np; p++) {
            Out &o = outs[p];
            const uint32_t frg_id = (uint32_t)frags.size();
            frag_map[o.key].push_back({frg_id, sid, o.bgn, o.end, o.ori});
            orc_fragment &f = push(o.aligned ? 0 : 2, sid, o.bgn - k, o.end);
            if (o.aligned) {
                f.reversed = o.rc ? 1 : 0; f.ref_frag = o.ref; f.seg_off = segs.size(); f.n_segs = (uint32_t)o.segs.size();
                for (const Seg &g : o.segs) segs.push_back({g.type, g.a, g.b});
            }
        }
        push(3, sid, pos(nm - 1) + 1, L);
        orc_free(mm);
    }
    *frags_out = (orc_fragment *)malloc(std::max<size_t>(frags.size(), 1) * sizeof(orc_fragment));
    *segs_out = (orc_alnseg *)malloc(std::max<size_t>(segs.size(), 1) * sizeof(orc_alnseg));
    if (!frags.empty()) memcpy(*frags_out, frags.data(), frags.size() * sizeof(orc_fragment));
    if (!segs.empty()) memcpy(*segs_out, segs.data(), segs.size() * sizeof(orc_alnseg));
    *n_frags = frags.size();
    *n_segs = segs.size();
    return 0;
}

}  // extern "C"
