"""GPU: the pgrtk-compatible Python surface (pgr_tk_b200/pgrtk_compat.py over the C ABI) against the oracle, in the
reference's own return shapes (pgr-tk/src/lib.rs)."""
import os

import numpy as np
import pytest

import orc
from pgr_tk_b200 import pgrtk_compat as pgrtk

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def test_seq_index_db_surface_matches_oracle():
    fa = os.path.join(GOLDEN, "test_seqs.fa")
    db = pgrtk.SeqIndexDB()
    db.load_from_fastx(fa)
    o = orc.Index(orc.mkspec(), 0)
    o.load_fasta(fa)
    recs = orc.parse_fasta(fa)
    assert db.get_shmmr_spec() == (80, 56, 4, 64, False)
    assert db.get_shmmr_map() == o.as_map()
    assert len(db.seq_info) == 66 and db.seq_info[3] == (recs[3][0], fa, len(recs[3][1]))
    assert bytes(db.get_sub_seq(fa, recs[5][0], 100, 200)) == recs[5][1][100:200]
    pl = db.get_shmmr_pair_list()
    assert len(pl) == 820 and pl == sorted(pl, key=lambda t: (t[0], t[1]))
    q = recs[7][1][500:3000]
    hits = db.query_fragment(q)
    pairs, off, sig = o.raw_query(q)
    exp = [((int(p["h0"]), int(p["h1"])), (int(p["bgn"]), int(p["end"]), int(p["ori"])),
            [(int(h["frg_id"]), int(h["sid"]), int(h["bgn"]), int(h["end"]), int(h["ori"])) for h in sig[int(off[i]):int(off[i + 1])]])
           for i, p in enumerate(pairs)]
    assert hits == exp and len(hits) > 3
    # a mutated query: pairs without a hit stay in the result with an empty signature list (seq_db.rs:1219-1226)
    qm = bytearray(q)
    for i in range(0, len(qm), 97):
        qm[i] = ord("A") if qm[i] != ord("A") else ord("C")
    hm = db.query_fragment(bytes(qm))
    pm, om, _ = o.raw_query(bytes(qm))
    assert len(hm) == len(pm) and [len(h[2]) for h in hm] == [int(om[i + 1] - om[i]) for i in range(len(pm))]
    assert any(len(h[2]) == 0 for h in hm)
    res = db.query_fragment_to_hps(q, 0.025, 128, 128, 128, 8)
    osid, otco, osc, ocho, ohits = o.query_fragment_to_hps(q, 0.025, max_count=128, max_count_query=128, max_count_target=128, max_aln_span=8)
    assert [r[0] for r in res] == [int(s) for s in osid]
    assert [[np.float32(sc) for sc, _ in r[1]] for r in res] == [[np.float32(osc[c]) for c in range(int(otco[t]), int(otco[t + 1]))] for t in range(len(osid))]
    assert sum(len(a) for r in res for _, a in r[1]) == len(ohits)
    adj = db.get_smp_adj_list(0)
    oadj = o.adj_list(0)
    assert len(adj) == len(oadj) and adj[0] == (int(oadj[0]["sid"]), (int(oadj[0]["a0"]), int(oadj[0]["a1"]), int(oadj[0]["ori0"])), (int(oadj[0]["b0"]), int(oadj[0]["b1"]), int(oadj[0]["ori1"])))
    pb = db.get_principal_bundles(0, 2)
    assert pb and all(len(b) > 0 for b in pb)
    rows = db.sort_adj_list_by_weighted_dfs(adj, adj[0][1])
    assert rows[0][0] == adj[0][1] and rows[0][1] is None
    sp = pgrtk.get_shmmr_pairs_from_seq(recs[0][1], 80, 56, 4, 64)
    assert len(sp) == len(orc.sequence_to_shmmrs(0, recs[0][1], orc.mkspec())) - 1


def smp_adj_list_for_seq_oracle(o, seq, sid, min_count):
    """seq_db.rs:946-1000 restated on the oracle's raw_query (same strict '<' pairs, seq_db.rs:1213-1217 / :962-966)"""
    pairs, off, _ = o.raw_query(seq)
    res = [(int(p["h0"]), int(p["h1"]), int(p["bgn"]), int(p["end"]), int(p["ori"])) for p in pairs]
    cnt = [int(off[i + 1] - off[i]) for i in range(len(res))]
    out = []
    if len(res) < 2:
        return out
    for i in range(len(res) - 1):
        v, w = res[i], res[i + 1]
        if cnt[i] == 0 or cnt[i + 1] == 0 or cnt[i] < min_count or cnt[i + 1] < min_count or v[3] != w[2]:
            continue
        out.append((sid, (v[0], v[1], v[4]), (w[0], w[1], w[4])))
        out.append((sid, (w[0], w[1], 1 - w[4]), (v[0], v[1], 1 - v[4])))
    return out


def test_mapg_gfa_both_methods(tmp_path):
    """generate_mapg_gfa(method="from_fragmap") and the per-sequence method (generate_smp_adj_list_for_seq, ext.rs:696-722)"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
    import pbundle_oracle as po
    fa = os.path.join(GOLDEN, "test_seqs.fa")
    db = pgrtk.SeqIndexDB()
    db.load_from_fastx(fa, w=48, k=56, r=4, min_span=12)
    o = orc.Index(orc.mkspec(48, 56, 4, 12), 0)
    o.load_fasta(fa)
    recs = orc.parse_fasta(fa)
    fmap = o.as_map()
    for min_count, keeps in ((0, None), (2, None), (3, [0, 5, 7]), (10**6, [1])):
        got = db.get_smp_adj_list_by_seq(min_count, keeps)
        exp = []
        for sid, (_, seq) in enumerate(recs):
            mc = 0 if keeps is not None and sid in keeps else min_count
            exp += smp_adj_list_for_seq_oracle(o, seq, sid, mc)
        assert got == exp, (min_count, keeps)
        p = str(tmp_path / "g.gfa")
        db.generate_mapg_gfa(min_count, p, "by_seq", keeps)
        assert open(p).read() == po.gfa_text(exp, fmap, 56)
    oadj = o.adj_list(2)
    exp2 = [(int(a["sid"]), (int(a["a0"]), int(a["a1"]), int(a["ori0"])), (int(a["b0"]), int(a["b1"]), int(a["ori1"]))) for a in oadj]
    p = str(tmp_path / "f.gfa")
    db.generate_mapg_gfa(2, p, "from_fragmap")
    assert open(p).read() == po.gfa_text(exp2, fmap, 56)
    assert len(db.get_smp_adj_list_by_seq(0)) > 100
