// pgr_oracle.cpp — CPU ORACLE (test infrastructure only; see pgr_oracle.h for the contract).
//
// Every function states the reference file:line it restates (paths relative to the reference
// root, GeneDx/pgr-tk).  Integer arithmetic is wrapping (Rust release semantics) wherever the
// reference could under/overflow; f32 arithmetic is never contracted (build with
// -ffp-contract=off) so that it rounds exactly like rustc's f32 code.
#include "pgr_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <vector>

namespace {

typedef orc_mm128 MM;
const MM MM_MAX = {UINT64_MAX, UINT64_MAX};

inline uint64_t mm_hash(const MM &m) { return m.x >> 8; }                       // shmmrutils.rs:247-249
inline uint32_t mm_pos(const MM &m) { return (uint32_t)((m.y & 0xFFFFFFFFull) >> 1); } // :261-263

// shmmrutils.rs:271-280 (Thomas Wang 64-bit mix, wrapping)
inline uint64_t u64hash(uint64_t key) {
    key = (~key) + (key << 21);
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}

// shmmrutils.rs:293-357
struct Ring {
    std::vector<MM> v;
    size_t size, start_pos, end_pos, len;
    explicit Ring(size_t n) : v(n, MM_MAX), size(n), start_pos(0), end_pos(0), len(0) {}
    void push(const MM &m) {
        v[end_pos] = m;
        end_pos = (end_pos + 1) % size;
        if (len < size) {
            len += 1;
        } else {
            start_pos = (start_pos + 1) % size;
        }
    }
    MM get_min() const {  // raw slots 0..len, strict '<' (first minimum in slot order)
        MM mn = MM_MAX;
        for (size_t i = 0; i < len; i++)
            if (v[i].x < mn.x) mn = v[i];
        return mn;
    }
    MM get(size_t i) const { return v[(start_pos + i) % size]; }
};

// shmmrutils.rs:359-415
std::vector<MM> reduce_shmmr(const std::vector<MM> &mers_in, uint32_t r, bool padding) {
    std::vector<MM> shmmrs;
    Ring rbuf(r);
    MM min_mer = MM_MAX;
    std::vector<MM> padded;
    const std::vector<MM> *mers = &mers_in;
    if (padding) {
        for (uint32_t i = 0; i + 1 < r; i++) padded.push_back(min_mer);
        padded.insert(padded.end(), mers_in.begin(), mers_in.end());
        for (uint32_t i = 0; i + 1 < r; i++) padded.push_back(min_mer);
        mers = &padded;
    }
    size_t pos = 0, mdist = 0;
    while (pos < mers->size()) {
        MM m = (*mers)[pos];
        rbuf.push(m);
        if (mdist == (size_t)(r - 1)) {
            min_mer = rbuf.get_min();
            size_t last_i = 0;
            for (size_t i = 0; i < rbuf.size; i++) {
                MM mm = rbuf.get(i);
                if (mm.x == min_mer.x) {
                    shmmrs.push_back(mm);
                    min_mer = mm;
                    last_i = i;
                }
            }
            mdist = (size_t)r - 1 - last_i;
            pos += 1;
            continue;
        } else if (m.x <= min_mer.x && pos >= (size_t)r) {
            shmmrs.push_back(m);
            min_mer = m;
            mdist = 0;
            pos += 1;
            continue;
        }
        mdist += 1;
        pos += 1;
    }
    return shmmrs;
}

// LUT of shmmrutils.rs:426-436: bytes 0..3 map to themselves, ACGT/acgt to 0..3, all else 4
struct Base2Bits {
    uint64_t t[256];
    Base2Bits() {
        for (int i = 0; i < 256; i++) t[i] = 4;
        t[0] = 0; t[1] = 1; t[2] = 2; t[3] = 3;
        t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3;
    }
};
const Base2Bits B2B;

// shmmrutils.rs:536-555 and :632-652 (min_span filter; u32 wrapping subtraction)
std::vector<MM> span_filter(const std::vector<MM> &s, uint32_t min_span) {
    std::vector<MM> out;
    size_t n = s.size();
    for (size_t i = 0; i < n; i++) {
        if (i != 0 && i != n - 1) {
            uint32_t p_pos = mm_pos(s[i - 1]), pos = mm_pos(s[i]), n_pos = mm_pos(s[i + 1]);
            uint64_t px = s[i - 1].x, x = s[i].x, nx = s[i + 1].x;
            if ((uint32_t)(pos - p_pos) > min_span && (uint32_t)(n_pos - pos) > min_span && px != x && x != nx)
                out.push_back(s[i]);
        } else {
            out.push_back(s[i]);
        }
    }
    return out;
}

// rolling registers, shmmrutils.rs:446-476 / :577-602
struct Roller {
    uint64_t f0, f1, r0, r1, mask;
    uint32_t shift;
    explicit Roller(uint32_t k) : f0(0), f1(0), r0(0), r1(0), mask(UINT64_MAX >> (64 - k)), shift(k - 1) {}
    inline void feed(uint8_t ch) {
        uint64_t c = B2B.t[ch];
        if (c < 4) {
            f0 = ((f0 << 1) | (c & 1)) & mask;
            f1 = ((f1 << 1) | ((c & 2) >> 1)) & mask;
            uint64_t rc = 3 ^ c;
            r0 = ((r0 >> 1) | ((rc & 1) << shift)) & mask;
            r1 = ((r1 >> 1) | (((rc & 2) >> 1) << shift)) & mask;
        }
    }
    inline bool palindrome() const { return f0 == r0 && f1 == r1; }
    inline bool forward() const { return !(r0 < f0); }            // :485-488 compares plane 0 only
    inline uint64_t hash() const {                                 // :490-492
        return forward() ? (u64hash(f0) ^ u64hash(f1 ^ 0xAD12CF59ull)) : (u64hash(r0) ^ u64hash(r1 ^ 0xAD12CF59ull));
    }
};

// shmmrutils.rs:417-556
std::vector<MM> sequence_to_shmmrs1(uint32_t rid, const uint8_t *seq, size_t L, uint32_t w, uint32_t k, uint32_t r,
                                    uint32_t min_span, bool padding) {
    std::vector<MM> shmmrs;
    size_t pos = 0, mdist = 0;
    Roller reg(k);
    Ring rbuf(w);
    MM min_mer = MM_MAX;
    const size_t rule2_end = L - (size_t)w + (size_t)k;  // wrapping usize, :518
    while (pos < L) {
        reg.feed(seq[pos]);
        if (reg.palindrome()) { pos += 1; continue; }
        if (pos < (size_t)k) { pos += 1; continue; }
        bool forward = reg.forward();
        uint64_t h = reg.hash();
        MM m;
        m.x = (h << 8) | (uint64_t)k;
        m.y = ((uint64_t)rid << 32) | ((uint64_t)pos << 1) | (forward ? 0ull : 1ull);
        rbuf.push(m);
        if (mdist == (size_t)(w - 1)) {
            min_mer = rbuf.get_min();
            for (size_t i = 0; i < rbuf.size; i++) {
                MM mm = rbuf.get(i);
                if (mm.x == min_mer.x) {
                    shmmrs.push_back(mm);
                    min_mer = mm;
                }
            }
            mdist = pos - (size_t)((min_mer.y & 0xFFFFFFFFull) >> 1);
            pos += 1;
            continue;
        } else if (m.x <= min_mer.x && pos >= (size_t)(w + k) && pos < rule2_end && pos < L) {
            shmmrs.push_back(m);
            min_mer = m;
            mdist = 0;
            pos += 1;
            continue;
        }
        mdist += 1;
        pos += 1;
    }
    if (r > 1) shmmrs = reduce_shmmr(reduce_shmmr(shmmrs, r, padding), r, padding);
    return span_filter(shmmrs, min_span);
}

// shmmrutils.rs:558-655
std::vector<MM> sequence_to_shmmrs2(uint32_t rid, const uint8_t *seq, size_t L, uint32_t k, uint32_t r, uint32_t min_span) {
    std::vector<MM> shmmrs;
    Roller reg(k);
    const uint64_t thr = (UINT64_MAX >> 4) >> r;
    for (size_t pos = 0; pos < L; pos++) {
        reg.feed(seq[pos]);
        if (reg.palindrome()) continue;
        if (pos < (size_t)k) continue;
        bool forward = reg.forward();
        uint64_t h = reg.hash();
        if (h < thr) {
            MM m;
            m.x = (h << 8) | (uint64_t)k;
            m.y = ((uint64_t)rid << 32) | ((uint64_t)pos << 1) | (forward ? 0ull : 1ull);
            shmmrs.push_back(m);
        }
    }
    return span_filter(shmmrs, min_span);
}

int check_spec(const orc_spec *s) {
    // shmmrutils.rs:443-445 / :575-576 asserts, plus the values for which the Rust code would
    // shift/underflow-panic (k == 0, w == 0)
    if (s->k == 0 || s->k > 56) return -2;
    if (!(s->r > 0 && s->r < 13)) return -2;
    if (!s->sketch && (s->w == 0 || s->w > 128)) return -2;
    return 0;
}

// shmmrutils.rs:657-669
std::vector<MM> sequence_to_shmmrs(uint32_t rid, const uint8_t *seq, size_t L, const orc_spec &s, bool padding) {
    if (!s.sketch) return sequence_to_shmmrs1(rid, seq, L, s.w, s.k, s.r, s.min_span, padding);
    return sequence_to_shmmrs2(rid, seq, L, s.k, s.r, s.min_span);
}

template <class F>
void parallel_for(size_t n, int nthreads, F f) {
    if (nthreads <= 1 || n <= 1) {
        for (size_t i = 0; i < n; i++) f(i);
        return;
    }
    std::atomic<size_t> next(0);
    std::vector<std::thread> th;
    int nt = (int)std::min<size_t>((size_t)nthreads, n);
    for (int t = 0; t < nt; t++)
        th.emplace_back([&]() {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= n) break;
                f(i);
            }
        });
    for (auto &t : th) t.join();
}

typedef std::pair<uint64_t, uint64_t> Key;
struct KeyHash {
    size_t operator()(const Key &k) const { return (size_t)(u64hash(k.first) ^ (k.second * 0x9E3779B97F4A7C15ull)); }
};
struct Sig { uint32_t frg_id, sid, bgn, end; uint8_t ori; };
struct SeqInfo { uint32_t sid; uint64_t len; std::string name, source; bool has_source; };

}  // namespace

struct orc_index {
    orc_spec spec;
    int mode;
    std::unordered_map<Key, std::vector<Sig>, KeyHash> frag_map;  // ShmmrToFrags, seq_db.rs:76
    std::vector<SeqInfo> seqs;
    uint32_t n_frags;  // == frags.len() of the FASTX path (seq_db.rs:203)
    std::vector<Key> sorted_keys() const {
        std::vector<Key> ks;
        ks.reserve(frag_map.size());
        for (auto &kv : frag_map) ks.push_back(kv.first);
        std::sort(ks.begin(), ks.end());
        return ks;
    }
};

namespace {

// One sequence's shimmers -> frag_map entries.
//  mode 0: seq_db.rs:189-357 (numbering only: :203-231 prefix/empty, :326-340 pairs, :342-347 suffix)
//  mode 1: seq_db.rs:360-418 + :594-612
void add_seq_to_map(orc_index *idx, uint32_t sid, const std::vector<MM> &shmmrs) {
    uint32_t frg_id;
    if (idx->mode == 0) {
        frg_id = idx->n_frags;
        if (shmmrs.empty()) { idx->n_frags += 2; return; }
        frg_id += 1;  // prefix fragment
    } else {
        frg_id = 0;
        if (shmmrs.empty()) return;
    }
    for (size_t i = 0; i + 1 < shmmrs.size(); i++) {
        uint64_t s0 = mm_hash(shmmrs[i]), s1 = mm_hash(shmmrs[i + 1]);
        Key key;
        uint8_t ori;
        if (s0 <= s1) { key = Key(s0, s1); ori = 0; } else { key = Key(s1, s0); ori = 1; }  // :238-242 / :391-395
        Sig sg = {frg_id, sid, mm_pos(shmmrs[i]) + 1, mm_pos(shmmrs[i + 1]) + 1, ori};
        idx->frag_map[key].push_back(sg);
        frg_id += 1;
    }
    if (idx->mode == 0) idx->n_frags = frg_id + 1;  // suffix fragment
}

void add_batch(orc_index *idx, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens,
               const std::vector<std::string> *names, const std::string *source, int nthreads) {
    // seq_db.rs:483-498 / :549-564: batches of at most 129 records; rayon over the batch (:461);
    // sequential insertion (:512-524 / :594-612)
    for (size_t b0 = 0; b0 < n; b0 += 129) {
        size_t b1 = std::min(n, b0 + 129);
        std::vector<std::vector<MM>> all(b1 - b0);
        parallel_for(b1 - b0, nthreads,
                     [&](size_t i) { all[i] = sequence_to_shmmrs(sids[b0 + i], seqs[b0 + i], lens[b0 + i], idx->spec, false); });
        for (size_t i = b0; i < b1; i++) {
            add_seq_to_map(idx, sids[i], all[i - b0]);
            SeqInfo si;
            si.sid = sids[i];
            si.len = lens[i];
            si.name = names ? (*names)[i] : std::string();
            si.has_source = source != nullptr;
            si.source = source ? *source : std::string();
            idx->seqs.push_back(si);
        }
    }
}

// fasta_io.rs:46-118: FASTA records; id = header up to first ' ' minus \n,' ',\r; seq = bytes up to the
// next '>' minus \n,'>',\r.  The constructor consumes the first byte of the file (:54-57).
bool parse_fasta_file(const char *path, std::vector<std::string> &names, std::vector<std::vector<uint8_t>> &seqs) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    std::vector<uint8_t> buf;
    uint8_t tmp[1 << 16];
    size_t got;
    while ((got = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    fclose(f);
    if (buf.empty()) return false;
    size_t p = 1;  // first byte consumed by FastaReader::new
    const size_t N = buf.size();
    for (;;) {
        if (p >= N) break;  // read_until returned 0 bytes -> None
        size_t e = p;
        while (e < N && buf[e] != '\n') e++;
        size_t line_end = (e < N) ? e + 1 : N;
        std::string id;
        for (size_t i = p; i < line_end; i++) {
            uint8_t c = buf[i];
            if (c == ' ') break;  // read_until(b' ')
            if (c != '\n' && c != '\r') id.push_back((char)c);
        }
        p = line_end;
        std::vector<uint8_t> s;
        while (p < N && buf[p] != '>') {
            uint8_t c = buf[p];
            if (c != '\n' && c != '\r') s.push_back(c);
            p++;
        }
        if (p < N) p++;  // consume '>'
        names.push_back(id);
        seqs.push_back(std::move(s));
    }
    return true;
}

struct HP {
    uint32_t qb, qe; uint8_t qo; uint32_t tb, te; uint8_t to;
    bool operator<(const HP &o) const {
        return std::tie(qb, qe, qo, tb, te, to) < std::tie(o.qb, o.qe, o.qo, o.tb, o.te, o.to);
    }
    bool operator==(const HP &o) const {
        return qb == o.qb && qe == o.qe && qo == o.qo && tb == o.tb && te == o.te && to == o.to;
    }
};
inline HP to_hp(const orc_hitpair &h) { HP r = {h.qb, h.qe, h.qo, h.tb, h.te, h.to}; return r; }
inline orc_hitpair from_hp(const HP &h) { orc_hitpair r; memset(&r, 0, sizeof r); r.qb = h.qb; r.qe = h.qe; r.qo = h.qo; r.tb = h.tb; r.te = h.te; r.to = h.to; return r; }

// aln.rs:12-142.  Maps keyed by HitPair value exactly like the reference (duplicates collapse);
// canonical chain-head rule replaces FxHashSet iteration order (DESIGN.md "canonical forms").
int sparse_aln(std::vector<HP> &hits, uint32_t max_span, float penalty, bool has_gap, uint32_t max_gap_u, bool oriented,
               std::vector<std::pair<float, std::vector<HP>>> &out) {
    std::stable_sort(hits.begin(), hits.end(), [](const HP &a, const HP &b) { return a.qb < b.qb; });  // :21
    if (hits.size() < 2) return -3;  // assert!(sp_hits.len() > 1) :24
    std::map<HP, float> v_s;
    std::map<HP, std::pair<bool, HP>> best_pre;
    const HP first = hits[0];
    v_s[first] = (float)first.qe - (float)first.qb;
    best_pre[first] = std::make_pair(false, first);
    for (size_t i = 1; i < hits.size(); i++) {
        const HP hp = hits[i];
        bool has_best = false;
        HP best_v = hp;
        float best_s = 0.0f;
        size_t j = i;
        std::set<std::tuple<uint32_t, uint32_t, uint8_t>> span_set;
        for (;;) {
            if (j == 0) break;
            j -= 1;
            const HP pre = hits[j];
            if (oriented) {
                if ((pre.qo ^ pre.to) != (hp.qo ^ hp.to)) continue;
            }
            if (has_gap) {
                float mg = (float)max_gap_u;
                if (hp.qo == hp.to) {
                    if (std::fabs((float)hp.qb - (float)pre.qe) > mg || std::fabs((float)hp.tb - (float)pre.te) > mg) continue;
                } else if (std::fabs((float)hp.qb - (float)pre.qe) > mg || std::fabs((float)hp.te - (float)pre.tb) > mg) {
                    continue;
                }
            }
            if (pre.qb == hp.qb && pre.qe == hp.qe && pre.qo == hp.qo) continue;  // :67
            span_set.insert(std::make_tuple(pre.qb, pre.qe, pre.qo));
            auto it = v_s.find(pre);
            float p_s = (it == v_s.end()) ? 0.0f : it->second;
            float s = p_s + ((float)hp.qe - (float)hp.qb);
            if (hp.qo == hp.to) {
                float g = std::fabs((float)hp.qb - (float)pre.qe) + std::fabs((float)hp.tb - (float)pre.te);
                float pg = penalty * g;
                s = s - pg;
            } else {
                float g = std::fabs((float)hp.qb - (float)pre.qe) + std::fabs((float)hp.te - (float)pre.tb);
                float pg = penalty * g;
                s = s - pg;
            }
            if (s > best_s) { best_s = s; best_v = pre; has_best = true; }
            if (span_set.size() >= (size_t)max_span) break;
        }
        if (best_s > 0.0f) {
            v_s[hp] = best_s;
            best_pre[hp] = std::make_pair(has_best, best_v);
        } else {
            v_s[hp] = (float)hp.qe - (float)hp.qb;
            best_pre[hp] = std::make_pair(false, hp);
        }
    }
    std::set<HP> unvisited(hits.begin(), hits.end());
    while (!unvisited.empty()) {
        float best_s = 0.0f;
        bool has = false;
        HP best_v = hits[0];
        for (size_t i = 0; i < hits.size(); i++) {  // canonical: index order in the sorted list
            if (!unvisited.count(hits[i])) continue;
            float s = v_s[hits[i]];
            if (s > best_s) { best_s = s; best_v = hits[i]; has = true; }
        }
        std::vector<HP> track;
        bool vsome = has;
        HP v = best_v;
        while (vsome) {
            if (!unvisited.count(v)) break;
            track.push_back(v);
            auto bp = best_pre[v];
            vsome = bp.first;
            v = bp.second;
        }
        if (track.empty()) return -4;  // the reference would spin forever here (all scores <= 0)
        std::reverse(track.begin(), track.end());
        for (auto &h : track) unvisited.erase(h);
        float bgn_s = v_s[track[0]];
        out.push_back(std::make_pair(best_s - bgn_s, track));
    }
    return 0;
}

struct RawHit { orc_qpair q; const std::vector<Sig> *sigs; };

// seq_db.rs:1200-1228
std::vector<RawHit> raw_query(const orc_index *idx, const uint8_t *seq, size_t len) {
    std::vector<MM> sh = sequence_to_shmmrs(0, seq, len, idx->spec, false);
    std::vector<RawHit> res;
    for (size_t i = 0; i + 1 < sh.size(); i++) {
        uint32_t p0 = mm_pos(sh[i]) + 1, p1 = mm_pos(sh[i + 1]) + 1;
        uint64_t s0 = mm_hash(sh[i]), s1 = mm_hash(sh[i + 1]);
        RawHit h;
        memset(&h.q, 0, sizeof h.q);
        if (s0 < s1) { h.q.h0 = s0; h.q.h1 = s1; h.q.ori = 0; } else { h.q.h0 = s1; h.q.h1 = s0; h.q.ori = 1; }
        h.q.bgn = p0;
        h.q.end = p1;
        auto it = idx->frag_map.find(Key(h.q.h0, h.q.h1));
        h.sigs = (it == idx->frag_map.end()) ? nullptr : &it->second;
        res.push_back(h);
    }
    return res;
}

template <class T>
T *dup_vec(const std::vector<T> &v) {
    T *p = (T *)malloc(std::max<size_t>(1, v.size()) * sizeof(T));
    if (!v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

}  // namespace

extern "C" {

uint64_t orc_u64hash(uint64_t key) { return u64hash(key); }
void orc_free(void *p) { free(p); }

int orc_sequence_to_shmmrs(uint32_t rid, const uint8_t *seq, size_t len, const orc_spec *spec, int padding,
                           orc_mm128 **out, size_t *n_out) {
    int rc = check_spec(spec);
    if (rc) return rc;
    std::vector<MM> v = sequence_to_shmmrs(rid, seq, len, *spec, padding != 0);
    *out = dup_vec(v);
    *n_out = v.size();
    return 0;
}

int orc_shmmrs_batch(size_t n, const uint32_t *rids, const uint8_t *const *seqs, const size_t *lens, const orc_spec *spec,
                     int padding, int nthreads, orc_mm128 **out, size_t *offsets) {
    int rc = check_spec(spec);
    if (rc) return rc;
    std::vector<std::vector<MM>> all(n);
    parallel_for(n, nthreads, [&](size_t i) { all[i] = sequence_to_shmmrs(rids[i], seqs[i], lens[i], *spec, padding != 0); });
    size_t tot = 0;
    for (size_t i = 0; i < n; i++) { offsets[i] = tot; tot += all[i].size(); }
    offsets[n] = tot;
    MM *o = (MM *)malloc(std::max<size_t>(1, tot) * sizeof(MM));
    for (size_t i = 0; i < n; i++)
        if (!all[i].empty()) memcpy(o + offsets[i], all[i].data(), all[i].size() * sizeof(MM));
    *out = o;
    return 0;
}

orc_index *orc_index_new(const orc_spec *spec, int frg_id_mode) {
    if (check_spec(spec)) return nullptr;
    orc_index *idx = new orc_index();
    idx->spec = *spec;
    idx->mode = frg_id_mode;
    idx->n_frags = 0;
    return idx;
}
void orc_index_free(orc_index *idx) { delete idx; }

int orc_index_add_batch(orc_index *idx, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens,
                        int nthreads) {
    add_batch(idx, n, sids, seqs, lens, nullptr, nullptr, nthreads);
    return 0;
}

int orc_index_load_fasta(orc_index *idx, const char *path, int nthreads) {
    std::vector<std::string> names;
    std::vector<std::vector<uint8_t>> seqs;
    if (!parse_fasta_file(path, names, seqs)) return -1;
    // seq_db.rs:473: FASTX path continues sid from seqs.len(); :543: index path restarts at 0
    uint32_t sid0 = (idx->mode == 0) ? (uint32_t)idx->seqs.size() : 0;
    std::vector<uint32_t> sids(seqs.size());
    std::vector<const uint8_t *> ptrs(seqs.size());
    std::vector<size_t> lens(seqs.size());
    for (size_t i = 0; i < seqs.size(); i++) { sids[i] = sid0 + (uint32_t)i; ptrs[i] = seqs[i].data(); lens[i] = seqs[i].size(); }
    std::string src(path);
    add_batch(idx, seqs.size(), sids.data(), ptrs.data(), lens.data(), &names, &src, nthreads);
    return 0;
}

size_t orc_index_n_keys(const orc_index *idx) { return idx->frag_map.size(); }
size_t orc_index_n_sigs(const orc_index *idx) {
    size_t n = 0;
    for (auto &kv : idx->frag_map) n += kv.second.size();
    return n;
}
size_t orc_index_n_seqs(const orc_index *idx) { return idx->seqs.size(); }
void orc_index_get_spec(const orc_index *idx, orc_spec *spec) { *spec = idx->spec; }

void orc_index_export(const orc_index *idx, uint64_t *keys, uint64_t *offsets, orc_sig *sigs) {
    std::vector<Key> ks = idx->sorted_keys();
    uint64_t off = 0;
    for (size_t i = 0; i < ks.size(); i++) {
        keys[2 * i] = ks[i].first;
        keys[2 * i + 1] = ks[i].second;
        offsets[i] = off;
        for (const Sig &s : idx->frag_map.at(ks[i])) {
            orc_sig o;
            memset(&o, 0, sizeof o);
            o.frg_id = s.frg_id; o.sid = s.sid; o.bgn = s.bgn; o.end = s.end; o.ori = s.ori;
            sigs[off++] = o;
        }
    }
    offsets[ks.size()] = off;
}

static void put_u32(std::vector<uint8_t> &b, uint32_t v) { for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i))); }
static void put_u64(std::vector<uint8_t> &b, uint64_t v) { for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i))); }

// seq_db.rs:1291-1326, keys written ascending (canonical form; the reference writes hash-map order)
int orc_index_write_mdb(const orc_index *idx, const char *path) {
    std::vector<uint8_t> buf;
    buf.push_back('m'); buf.push_back('d'); buf.push_back('b');
    put_u32(buf, idx->spec.w); put_u32(buf, idx->spec.k); put_u32(buf, idx->spec.r);
    put_u32(buf, idx->spec.min_span); put_u32(buf, idx->spec.sketch ? 1u : 0u);
    put_u64(buf, idx->frag_map.size());
    for (const Key &k : idx->sorted_keys()) {
        const std::vector<Sig> &v = idx->frag_map.at(k);
        put_u64(buf, k.first); put_u64(buf, k.second); put_u64(buf, v.size());
        for (const Sig &s : v) { put_u32(buf, s.frg_id); put_u32(buf, s.sid); put_u32(buf, s.bgn); put_u32(buf, s.end); buf.push_back(s.ori); }
    }
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    size_t wr = fwrite(buf.data(), 1, buf.size(), f);
    fclose(f);
    return wr == buf.size() ? 0 : -1;
}

// seq_db.rs:795-807
int orc_index_write_midx(const orc_index *idx, const char *path) {
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    for (const SeqInfo &s : idx->seqs)
        fprintf(f, "%u\t%llu\t%s\t%s\n", s.sid, (unsigned long long)s.len, s.name.c_str(), s.has_source ? s.source.c_str() : "-");
    fclose(f);
    return 0;
}

// seq_db.rs:1328-1407
orc_index *orc_index_read_mdb(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) return nullptr;
    std::vector<uint8_t> buf;
    uint8_t tmp[1 << 16];
    size_t got;
    while ((got = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    fclose(f);
    if (buf.size() < 31 || memcmp(buf.data(), "mdb", 3) != 0) return nullptr;
    size_t c = 3;
    auto rd32 = [&]() { uint32_t v; memcpy(&v, &buf[c], 4); c += 4; return v; };
    auto rd64 = [&]() { uint64_t v; memcpy(&v, &buf[c], 8); c += 8; return v; };
    orc_index *idx = new orc_index();
    idx->spec.w = rd32(); idx->spec.k = rd32(); idx->spec.r = rd32(); idx->spec.min_span = rd32();
    idx->spec.sketch = (rd32() & 1u) ? 1 : 0;
    idx->mode = 0;
    idx->n_frags = 0;
    uint64_t nk = rd64();
    for (uint64_t i = 0; i < nk; i++) {
        if (c + 24 > buf.size()) { delete idx; return nullptr; }
        uint64_t k1 = rd64(), k2 = rd64(), vl = rd64();
        if (c + 17 * vl > buf.size()) { delete idx; return nullptr; }
        std::vector<Sig> v((size_t)vl);
        for (uint64_t j = 0; j < vl; j++) {
            v[j].frg_id = rd32(); v[j].sid = rd32(); v[j].bgn = rd32(); v[j].end = rd32(); v[j].ori = buf[c]; c += 1;
        }
        idx->frag_map[Key(k1, k2)] = std::move(v);
    }
    return idx;
}

int orc_index_seq_info(const orc_index *idx, size_t i, uint32_t *sid, uint64_t *len, const char **name, const char **source) {
    if (i >= idx->seqs.size()) return -1;
    const SeqInfo &s = idx->seqs[i];
    *sid = s.sid; *len = s.len; *name = s.name.c_str(); *source = s.has_source ? s.source.c_str() : "-";
    return 0;
}

int orc_parse_fasta(const char *path, size_t *n, char ***names, uint8_t ***seqs, size_t **lens) {
    std::vector<std::string> nm;
    std::vector<std::vector<uint8_t>> sq;
    if (!parse_fasta_file(path, nm, sq)) return -1;
    *n = nm.size();
    *names = (char **)malloc(std::max<size_t>(1, nm.size()) * sizeof(char *));
    *seqs = (uint8_t **)malloc(std::max<size_t>(1, nm.size()) * sizeof(uint8_t *));
    *lens = (size_t *)malloc(std::max<size_t>(1, nm.size()) * sizeof(size_t));
    for (size_t i = 0; i < nm.size(); i++) {
        (*names)[i] = strdup(nm[i].c_str());
        (*seqs)[i] = (uint8_t *)malloc(std::max<size_t>(1, sq[i].size()));
        if (!sq[i].empty()) memcpy((*seqs)[i], sq[i].data(), sq[i].size());
        (*lens)[i] = sq[i].size();
    }
    return 0;
}

int orc_raw_query(const orc_index *idx, const uint8_t *seq, size_t len, orc_qpair **pairs, size_t *n_pairs,
                  uint64_t **hit_off, orc_sig **hits) {
    std::vector<RawHit> res = raw_query(idx, seq, len);
    std::vector<orc_qpair> qp;
    std::vector<uint64_t> off;
    std::vector<orc_sig> hs;
    for (auto &h : res) {
        qp.push_back(h.q);
        off.push_back(hs.size());
        if (h.sigs)
            for (const Sig &s : *h.sigs) {
                orc_sig o;
                memset(&o, 0, sizeof o);
                o.frg_id = s.frg_id; o.sid = s.sid; o.bgn = s.bgn; o.end = s.end; o.ori = s.ori;
                hs.push_back(o);
            }
    }
    off.push_back(hs.size());
    *pairs = dup_vec(qp);
    *n_pairs = qp.size();
    *hit_off = dup_vec(off);
    *hits = dup_vec(hs);
    return 0;
}

int orc_sparse_aln(orc_hitpair *hits, size_t n, uint32_t max_span, float penalty, int64_t max_gap, int oriented,
                   size_t *n_chains, uint64_t **chain_off, float **scores, orc_hitpair **chain_hits) {
    std::vector<HP> h(n);
    for (size_t i = 0; i < n; i++) h[i] = to_hp(hits[i]);
    std::vector<std::pair<float, std::vector<HP>>> out;
    int rc = sparse_aln(h, max_span, penalty, max_gap >= 0, (uint32_t)std::max<int64_t>(0, max_gap), oriented != 0, out);
    if (rc) return rc;
    for (size_t i = 0; i < n; i++) hits[i] = from_hp(h[i]);
    std::vector<uint64_t> off;
    std::vector<float> sc;
    std::vector<orc_hitpair> ch;
    for (auto &c : out) {
        off.push_back(ch.size());
        sc.push_back(c.first);
        for (auto &x : c.second) ch.push_back(from_hp(x));
    }
    off.push_back(ch.size());
    *n_chains = out.size();
    *chain_off = dup_vec(off);
    *scores = dup_vec(sc);
    *chain_hits = dup_vec(ch);
    return 0;
}

// aln.rs:147-242 (the dead recomputation of the query's own pair counts, :163-170, is skipped:
// its result is never read)
int orc_query_fragment_to_hps(const orc_index *idx, const uint8_t *seq, size_t len, float penalty, int64_t max_count,
                              int64_t max_count_query, int64_t max_count_target, int64_t max_aln_span, int64_t max_gap,
                              int oriented, size_t *n_targets, uint32_t **target_sids, uint64_t **target_chain_off,
                              float **chain_scores, uint64_t **chain_hit_off, orc_hitpair **chain_hits) {
    std::vector<RawHit> raw = raw_query(idx, seq, len);
    std::map<Key, uint32_t> pair_count;
    std::map<std::tuple<uint64_t, uint64_t, uint32_t>, uint32_t> target_count;
    for (auto &h : raw) {
        pair_count[Key(h.q.h0, h.q.h1)] += 1;
        if (h.sigs)
            for (const Sig &s : *h.sigs) target_count[std::make_tuple(h.q.h0, h.q.h1, s.sid)] += 1;
    }
    const uint32_t mc = max_count < 0 ? 128u : (uint32_t)max_count;
    const uint32_t mcq = max_count_query < 0 ? 128u : (uint32_t)max_count_query;
    const uint32_t mct = max_count_target < 0 ? 128u : (uint32_t)max_count_target;
    std::map<uint32_t, std::vector<HP>> by_sid;  // ordered => canonical target order
    for (auto &h : raw) {
        uint32_t count = pair_count[Key(h.q.h0, h.q.h1)];
        if (count > mc) continue;
        if (count > mcq) continue;  // aln.rs:208-211 compares the same count
        if (!h.sigs) continue;
        for (const Sig &s : *h.sigs) {
            uint32_t ct = target_count[std::make_tuple(h.q.h0, h.q.h1, s.sid)];
            if (ct > mct) continue;
            HP hp = {h.q.bgn, h.q.end, h.q.ori, s.bgn, s.end, s.ori};
            by_sid[s.sid].push_back(hp);
        }
    }
    const uint32_t span = max_aln_span < 0 ? 8u : (uint32_t)max_aln_span;
    std::vector<uint32_t> sids;
    std::vector<uint64_t> tco, cho;
    std::vector<float> sc;
    std::vector<orc_hitpair> ch;
    for (auto &kv : by_sid) {
        if (kv.second.size() <= 1) continue;
        std::vector<std::pair<float, std::vector<HP>>> out;
        int rc = sparse_aln(kv.second, span, penalty, max_gap >= 0, (uint32_t)std::max<int64_t>(0, max_gap), oriented != 0, out);
        if (rc) return rc;
        sids.push_back(kv.first);
        tco.push_back(sc.size());
        for (auto &c : out) {
            cho.push_back(ch.size());
            sc.push_back(c.first);
            for (auto &x : c.second) ch.push_back(from_hp(x));
        }
    }
    tco.push_back(sc.size());
    cho.push_back(ch.size());
    *n_targets = sids.size();
    *target_sids = dup_vec(sids);
    *target_chain_off = dup_vec(tco);
    *chain_scores = dup_vec(sc);
    *chain_hit_off = dup_vec(cho);
    *chain_hits = dup_vec(ch);
    return 0;
}

// seq_db.rs:876-944
int orc_adj_list(const orc_index *idx, size_t min_count, const uint32_t *keeps, size_t n_keeps, int has_keeps,
                 orc_adj **out, size_t *n_out) {
    typedef std::tuple<uint32_t, uint32_t, uint32_t, uint64_t, uint64_t, uint8_t> Row;  // sid,bgn,end,(h0,h1,ori)
    std::vector<Row> rows;
    for (auto &kv : idx->frag_map)
        for (const Sig &s : kv.second) rows.push_back(Row(s.sid, s.bgn, s.end, kv.first.first, kv.first.second, s.ori));
    std::vector<orc_adj> res;
    if (rows.size() >= 2) {
        std::sort(rows.begin(), rows.end());
        std::set<uint32_t> keep(keeps, keeps + (has_keeps ? n_keeps : 0));
        std::vector<char> ok(rows.size());
        for (size_t i = 0; i < rows.size(); i++) {
            const Row &v = rows[i];
            size_t cnt = idx->frag_map.at(Key(std::get<3>(v), std::get<4>(v))).size();
            ok[i] = (cnt >= min_count) || (has_keeps && keep.count(std::get<0>(v)));
        }
        for (size_t i = 0; i + 1 < rows.size(); i++) {
            if (!ok[i] || !ok[i + 1]) continue;
            const Row &v = rows[i], &w = rows[i + 1];
            if (std::get<0>(v) != std::get<0>(w) || std::get<2>(v) != std::get<1>(w)) continue;
            orc_adj a;
            memset(&a, 0, sizeof a);
            a.sid = std::get<0>(v);
            a.a0 = std::get<3>(v); a.a1 = std::get<4>(v); a.ori0 = std::get<5>(v);
            a.b0 = std::get<3>(w); a.b1 = std::get<4>(w); a.ori1 = std::get<5>(w);
            res.push_back(a);
            orc_adj b;
            memset(&b, 0, sizeof b);
            b.sid = std::get<0>(v);
            b.a0 = std::get<3>(w); b.a1 = std::get<4>(w); b.ori0 = (uint8_t)(1 - std::get<5>(w));
            b.b0 = std::get<3>(v); b.b1 = std::get<4>(v); b.ori1 = (uint8_t)(1 - std::get<5>(v));
            res.push_back(b);
        }
    }
    *out = dup_vec(res);
    *n_out = res.size();
    return 0;
}

}  // extern "C"
