"""CPU: the host-side post-processing of the C++ mirror, compiled without any device call and compared with the Python
restatements on random inputs:
  * query_post.hpp   merge_query_hits                      (pgr-query.rs:166-285)   vs oracle/query_post_oracle.py
  * pbundle.hpp      principal_bundles_with_id, group_smps (ext.rs:552-650, pgr-pbundle-decomp.rs:62-137) vs oracle/pbundle_oracle.py"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pbundle_oracle as pbo  # noqa: E402
import query_post_oracle as qpo  # noqa: E402

HOST = os.path.join(ROOT, "pgr_tk_b200", "host")

QPROG = r'''
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "query_post.hpp"
using namespace pgrb200;
// input: tol, n_targets, then per target: sid n_chains, per chain: n_hits, per hit: qb qe qo tb te to
int main() {
    long tol; size_t nt;
    if (scanf("%ld %zu", &tol, &nt) != 2) return 2;
    std::vector<uint64_t> q_off = {0, nt}, t_off = {0}, c_off = {0};
    std::vector<uint32_t> sids; std::vector<float> sc; std::vector<pgr_hit_pair> hits;
    for (size_t t = 0; t < nt; t++) {
        unsigned sid; size_t nc;
        if (scanf("%u %zu", &sid, &nc) != 2) return 2;
        sids.push_back(sid);
        for (size_t c = 0; c < nc; c++) {
            size_t nh;
            if (scanf("%zu", &nh) != 1) return 2;
            for (size_t h = 0; h < nh; h++) {
                unsigned a[6];
                if (scanf("%u %u %u %u %u %u", a, a + 1, a + 2, a + 3, a + 4, a + 5) != 6) return 2;
                pgr_hit_pair p; memset(&p, 0, sizeof p);
                p.qb = a[0]; p.qe = a[1]; p.qo = (uint8_t)a[2]; p.tb = a[3]; p.te = a[4]; p.to = (uint8_t)a[5];
                hits.push_back(p);
            }
            c_off.push_back(hits.size()); sc.push_back(1.0f);
        }
        t_off.push_back(sc.size());
    }
    pgr_query_result r; memset(&r, 0, sizeof r);
    r.n_queries = 1; r.n_targets = nt; r.n_chains = sc.size(); r.n_hits = hits.size();
    r.q_target_off = q_off.data(); r.target_sid = sids.data(); r.target_chain_off = t_off.data(); r.chain_score = sc.data();
    r.chain_hit_off = c_off.data(); r.hits = hits.data();
    for (const auto &tr : merge_query_hits(r, 0, tol))
        for (const auto &g : tr.regions) {
            printf("%u %u %u %u %u %zu", tr.sid, g.bgn, g.end, g.len, g.orientation, g.aln.size());
            for (const auto &h : g.aln) printf(" %u,%u,%u,%u,%u,%u", h.qb, h.qe, h.qo, h.tb, h.te, h.to);
            printf("\n");
        }
    return 0;
}
'''

PPROG = r'''
#include <cstdio>
#include "pbundle.hpp"
using namespace pgrb200;
// input: n_bundles, per bundle: n, per vertex: h0 h1 ori; n_seqs, per seq: n, per smp: h0 h1 bgn end ori; cutoff merge
int main() {
    size_t nb;
    if (scanf("%zu", &nb) != 1) return 2;
    std::vector<std::vector<BundleVertex>> pb(nb);
    for (auto &b : pb) { size_t n; if (scanf("%zu", &n) != 1) return 2; b.resize(n); for (auto &v : b) { unsigned o; if (scanf("%lu %lu %u", &v.h0, &v.h1, &o) != 3) return 2; v.ori = (uint8_t)o; } }
    size_t ns;
    if (scanf("%zu", &ns) != 1) return 2;
    std::vector<std::vector<Smp>> smps(ns);
    for (auto &s : smps) { size_t n; if (scanf("%zu", &n) != 1) return 2; s.resize(n); for (auto &v : s) { unsigned o; if (scanf("%lu %lu %u %u %u", &v.h0, &v.h1, &v.bgn, &v.end, &o) != 5) return 2; v.ori = (uint8_t)o; } }
    size_t cutoff, merge;
    if (scanf("%zu %zu", &cutoff, &merge) != 2) return 2;
    std::vector<BundleWithId> out; VertexMap vmap;
    principal_bundles_with_id(pb, smps, out, vmap);
    for (const auto &b : out) { printf("B %zu %zu", b.bundle_id, b.mean_order); for (const auto &v : b.vertices) printf(" %lu,%lu,%u", v.h0, v.h1, (unsigned)v.ori); printf("\n"); }
    for (size_t s = 0; s < ns; s++)
        for (const auto &p : group_smps_by_principle_bundle_id(smps[s], vmap, cutoff, merge)) {
            printf("P %zu", s);
            for (const auto &e : p) printf(" %lu,%zu,%u,%zu", e.smp.h0, e.bid, e.d, e.bpos);
            printf("\n");
        }
    return 0;
}
'''


def _compile(tmp_path, name, prog):
    src = tmp_path / (name + ".cpp")
    src.write_text(prog)
    exe = str(tmp_path / name)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", HOST, "-o", exe, str(src)])
    return exe


def test_merge_query_hits_cpp_equals_oracle(tmp_path):
    exe = _compile(tmp_path, "qpost", QPROG)
    rng = np.random.default_rng(3)
    for trial in range(25):
        tol = int(rng.choice([50, 1000, 100000]))
        targets, lines = [], []
        sids = sorted(rng.choice(50, size=int(rng.integers(1, 6)), replace=False).tolist())
        for sid in sids:
            alns = []
            for _ in range(int(rng.integers(1, 6))):
                n = int(rng.integers(1, 9))
                t0 = int(rng.integers(0, 5000))
                q0 = int(rng.integers(0, 5000))
                rev = int(rng.random() < 0.4)
                hp = []
                for i in range(n):
                    tb = t0 + (n - i if rev else i) * 300 + int(rng.integers(0, 50))
                    hp.append(((q0 + i * 300, q0 + i * 300 + 250, 0), (tb, tb + 250 + int(rng.integers(0, 30)), rev if rng.random() < 0.9 else 1 - rev)))
                alns.append((1.0, hp))
            targets.append((sid, alns))
        inp = ["%d %d" % (tol, len(targets))]
        for sid, alns in targets:
            inp.append("%d %d" % (sid, len(alns)))
            for _, hp in alns:
                inp.append(str(len(hp)) + " " + " ".join("%d %d %d %d %d %d" % (q + t) for q, t in hp))
        got = subprocess.run([exe], input="\n".join(inp) + "\n", capture_output=True, text=True, check=True).stdout
        exp = []
        for sid, rgns in qpo.merge_query_hits(targets, tol):
            for b, e, ln, ori, aln in rgns:
                exp.append("%d %d %d %d %d %d " % (sid, b, e, ln, ori, len(aln)) + " ".join("%d,%d,%d,%d,%d,%d" % (q + t) for q, t in aln))
        assert got == "".join(l + "\n" for l in exp), trial


def test_bundle_ids_and_grouping_cpp_equals_oracle(tmp_path):
    exe = _compile(tmp_path, "pbund", PPROG)
    rng = np.random.default_rng(8)
    for trial in range(25):
        keys = [(int(rng.integers(1, 1 << 40)), int(rng.integers(1, 1 << 40))) for _ in range(30)]
        perm = rng.permutation(24)
        pb, at = [], 0
        for n in (9, 7, 5, 3):
            pb.append([(keys[int(i)][0], keys[int(i)][1], int(rng.integers(0, 2))) for i in perm[at:at + n]])
            at += n
        smps = {}
        for s in range(int(rng.integers(1, 5))):
            pos, lst = 0, []
            for _ in range(int(rng.integers(5, 40))):
                kk = keys[int(rng.integers(0, 30))]
                ln = int(rng.integers(100, 3000))
                lst.append((kk[0], kk[1], pos, pos + ln, int(rng.integers(0, 2))))
                pos += ln
            smps[s] = lst
        cutoff, merge = int(rng.choice([0, 500, 2500])), int(rng.choice([0, 2000, 10000]))
        inp = [str(len(pb))]
        for b in pb:
            inp.append(str(len(b)) + " " + " ".join("%d %d %d" % v for v in b))
        inp.append(str(len(smps)))
        for s in range(len(smps)):
            inp.append(str(len(smps[s])) + " " + " ".join("%d %d %d %d %d" % v for v in smps[s]))
        inp.append("%d %d" % (cutoff, merge))
        got = subprocess.run([exe], input="\n".join(inp) + "\n", capture_output=True, text=True, check=True).stdout
        pbid, vmap = pbo.principal_bundles_with_id(pb, smps)
        exp = ["B %d %d " % (b, o if o < 2 ** 63 else 2 ** 64 - 1) + " ".join("%d,%d,%d" % v for v in verts) for b, o, verts in pbid]
        exp = [l.rstrip() if l.endswith(" ") else l for l in exp]
        for s in range(len(smps)):
            for p in pbo.group_smps(smps[s], vmap, cutoff, merge):
                exp.append("P %d " % s + " ".join("%d,%d,%d,%d" % (e[0][0], e[1], e[2], e[3]) for e in p))
        assert got == "".join(l + "\n" for l in exp), trial


GPROG = r'''
#include <cstdio>
#include "mapg_gfa.hpp"
using namespace pgrb200;
// input: k; n_keys, per key: h0 h1 n, per sig: frg sid bgn end ori; n_adj, per pair: sid a0 a1 o0 b0 b1 o1; n_vmap, per: h0 h1 bundle ori pos
int main(int argc, char **argv) {
    unsigned k; size_t nk;
    if (scanf("%u %zu", &k, &nk) != 2) return 2;
    IndexCsr csr; csr.offsets.push_back(0);
    for (size_t i = 0; i < nk; i++) {
        uint64_t h0, h1; size_t n;
        if (scanf("%lu %lu %zu", &h0, &h1, &n) != 3) return 2;
        csr.keys.push_back(h0); csr.keys.push_back(h1);
        for (size_t j = 0; j < n; j++) { pgr_frag_sig s; memset(&s, 0, sizeof s); unsigned o; if (scanf("%u %u %u %u %u", &s.frg_id, &s.sid, &s.bgn, &s.end, &o) != 5) return 2; s.ori = (uint8_t)o; csr.sigs.push_back(s); }
        csr.offsets.push_back(csr.sigs.size());
    }
    size_t na;
    if (scanf("%zu", &na) != 1) return 2;
    std::vector<pgr_adj_pair> adj(na);
    for (auto &a : adj) { memset(&a, 0, sizeof a); unsigned o0, o1; if (scanf("%u %lu %lu %u %lu %lu %u", &a.sid, &a.a0, &a.a1, &o0, &a.b0, &a.b1, &o1) != 7) return 2; a.ori0 = (uint8_t)o0; a.ori1 = (uint8_t)o1; }
    size_t nv;
    if (scanf("%zu", &nv) != 1) return 2;
    VertexMap vm;
    for (size_t i = 0; i < nv; i++) { uint64_t h0, h1; size_t b, p; unsigned o; if (scanf("%lu %lu %zu %u %zu", &h0, &h1, &b, &o, &p) != 5) return 2; vm[{h0, h1}] = {b, (uint8_t)o, p}; }
    std::vector<CompactSeq> seqs = {{0, 20, "a", "f.fa"}, {1, 30, "b", ""}};
    pgr_shmmr_spec sp{80, k, 4, 64, 0};
    return write_gfa(argv[1], adj.data(), adj.size(), csr, k, nv ? &vm : nullptr) && write_mapg_idx(argv[2], sp, seqs, csr) ? 0 : 3;
}
'''


def test_gfa_writers_cpp_equal_oracle(tmp_path):
    import pgr_tk_b200 as pg
    if not os.path.exists(pg.library_path()):
        pg.build_library()
    src = tmp_path / "gfa.cpp"
    src.write_text("#include <cstring>\n" + GPROG)
    exe = str(tmp_path / "gfa")
    libdir = os.path.dirname(pg.library_path())
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", HOST, "-o", exe, str(src), "-L" + libdir, "-lpgr_b200", "-Wl,-rpath," + libdir])
    rng = np.random.default_rng(12)
    for trial in range(10):
        keys = sorted({(int(rng.integers(1, 1 << 50)), int(rng.integers(1, 1 << 50))) for _ in range(12)})
        fmap, frg = {}, 0
        for kk in keys:
            sigs = []
            for _ in range(int(rng.integers(1, 5))):
                b = int(rng.integers(57, 5000))
                sigs.append((frg, int(rng.integers(0, 2)), b, b + int(rng.integers(10, 900)), int(rng.integers(0, 2))))
                frg += 1
            fmap[kk] = sigs
        adj = []
        for _ in range(int(rng.integers(3, 30))):
            v, w = keys[int(rng.integers(0, len(keys)))], keys[int(rng.integers(0, len(keys)))]
            adj.append((int(rng.integers(0, 2)), (v[0], v[1], int(rng.integers(0, 2))), (w[0], w[1], int(rng.integers(0, 2)))))
        vmap = {kk: (int(rng.integers(0, 4)), int(rng.integers(0, 2)), int(rng.integers(0, 9))) for kk in keys[::3]} if trial % 2 else {}
        inp = ["56 %d" % len(keys)]
        for kk in keys:
            inp.append("%d %d %d " % (kk[0], kk[1], len(fmap[kk])) + " ".join("%d %d %d %d %d" % s for s in fmap[kk]))
        inp.append(str(len(adj)))
        for sid, v, w in adj:
            inp.append("%d %d %d %d %d %d %d" % (sid, v[0], v[1], v[2], w[0], w[1], w[2]))
        inp.append(str(len(vmap)))
        for kk, b in vmap.items():
            inp.append("%d %d %d %d %d" % (kk[0], kk[1], b[0], b[1], b[2]))
        g, x = str(tmp_path / "o.gfa"), str(tmp_path / "o.idx")
        subprocess.run([exe, g, x], input="\n".join(inp) + "\n", text=True, check=True)
        assert open(g).read() == pbo.gfa_text(adj, fmap, 56, vmap if vmap else None), trial
        assert open(x).read() == pbo.mapg_idx_text((80, 56, 4, 64, 0), [(0, 20, "a", "f.fa"), (1, 30, "b", None)], fmap)
