// pdb_io.hpp — the principal-bundle file of pgr-pbundle-decomp (pgr-bin/src/bin/pgr-pbundle-decomp.rs:158-226 read,
// :367-396 write): 7-byte tag "PDB:0.5" followed by bincode 2 `config::standard()` (variable-length integers, usize as u64)
// of the tuple (w, k, r, min_span: u32; min_branch_size, min_cov: usize; PrincipalBundlesWithId = Vec<(usize, usize,
// Vec<(u64, u64, u8)>)>; VertexToBundleIdMap = HashMap<(u64, u64), (usize, u8, usize)>).  A map is encoded as its length
// and its entries in iteration order; the reference iterates an FxHashMap there, so the entry order of its files is not
// reproducible — entries are written by ascending key here, and any order is accepted when reading.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "pbundle.hpp"

namespace pgrb200 {

struct PdbData {
    uint32_t w = 0, k = 0, r = 0, min_span = 0;
    uint64_t min_branch_size = 0, min_cov = 0;
    std::vector<BundleWithId> bundles;
    VertexMap vmap;
};

namespace pdb_detail {
inline void put_varint(std::vector<uint8_t> &o, uint64_t v) {
    if (v < 251) { o.push_back((uint8_t)v); return; }
    int n; uint8_t tag;
    if (v < (1ull << 16)) { n = 2; tag = 251; } else if (v < (1ull << 32)) { n = 4; tag = 252; } else { n = 8; tag = 253; }
    o.push_back(tag);
    for (int i = 0; i < n; i++) o.push_back((uint8_t)(v >> (8 * i)));
}
struct Reader {
    const uint8_t *p, *e;
    bool ok = true;
    uint8_t u8() { if (p >= e) { ok = false; return 0; } return *p++; }
    uint64_t varint() {
        const uint8_t t = u8();
        if (t < 251) return t;
        const int n = t == 251 ? 2 : t == 252 ? 4 : t == 253 ? 8 : 16;
        if (e - p < n) { ok = false; return 0; }
        uint64_t v = 0;
        for (int i = 0; i < n && i < 8; i++) v |= (uint64_t)p[i] << (8 * i);
        p += n;
        return v;
    }
    // element count, bounded by the remaining bytes before anything is allocated from it
    uint64_t count() { const uint64_t n = varint(); if (!ok || n > (uint64_t)(e - p)) { ok = false; return 0; } return n; }
};
}  // namespace pdb_detail

inline std::vector<uint8_t> encode_pdb(const PdbData &d) {
    using pdb_detail::put_varint;
    std::vector<uint8_t> o = {'P', 'D', 'B', ':', '0', '.', '5'};
    put_varint(o, d.w); put_varint(o, d.k); put_varint(o, d.r); put_varint(o, d.min_span);
    put_varint(o, d.min_branch_size); put_varint(o, d.min_cov);
    put_varint(o, d.bundles.size());
    for (const auto &b : d.bundles) {
        put_varint(o, b.bundle_id); put_varint(o, b.mean_order); put_varint(o, b.vertices.size());
        for (const auto &v : b.vertices) { put_varint(o, v.h0); put_varint(o, v.h1); o.push_back(v.ori); }
    }
    std::vector<std::pair<std::pair<uint64_t, uint64_t>, BundleRef>> ent(d.vmap.begin(), d.vmap.end());
    std::sort(ent.begin(), ent.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    put_varint(o, ent.size());
    for (const auto &e : ent) {
        put_varint(o, e.first.first); put_varint(o, e.first.second);
        put_varint(o, e.second.bundle_id); o.push_back(e.second.direction); put_varint(o, e.second.pos);
    }
    return o;
}

inline bool decode_pdb(const std::vector<uint8_t> &buf, PdbData &d) {
    if (buf.size() < 7 || memcmp(buf.data(), "PDB:0.5", 7) != 0) return false;
    pdb_detail::Reader r{buf.data() + 7, buf.data() + buf.size()};
    d.w = (uint32_t)r.varint(); d.k = (uint32_t)r.varint(); d.r = (uint32_t)r.varint(); d.min_span = (uint32_t)r.varint();
    d.min_branch_size = r.varint(); d.min_cov = r.varint();
    d.bundles.assign((size_t)r.count(), {});
    for (auto &b : d.bundles) {
        if (!r.ok) return false;
        b.bundle_id = (size_t)r.varint(); b.mean_order = (size_t)r.varint();
        b.vertices.assign((size_t)r.count(), {});
        for (auto &v : b.vertices) { v.h0 = r.varint(); v.h1 = r.varint(); v.ori = r.u8(); }
    }
    d.vmap.clear();
    const uint64_t n = r.count();
    for (uint64_t i = 0; i < n && r.ok; i++) {
        const uint64_t h0 = r.varint(), h1 = r.varint();
        BundleRef br;
        br.bundle_id = (size_t)r.varint(); br.direction = r.u8(); br.pos = (size_t)r.varint();
        d.vmap[{h0, h1}] = br;
    }
    return r.ok && r.p == r.e;
}

inline bool write_pdb(const std::string &path, const PdbData &d) {
    const std::vector<uint8_t> o = encode_pdb(d);
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = fwrite(o.data(), 1, o.size(), f) == o.size();
    fclose(f);
    return ok;
}

inline bool read_pdb(const std::string &path, PdbData &d) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    std::vector<uint8_t> buf;
    uint8_t tmp[1 << 16];
    size_t got;
    while ((got = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    fclose(f);
    return decode_pdb(buf, d);
}

}  // namespace pgrb200
