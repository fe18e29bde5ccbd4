"""GPU parity: raw_query_fragment, query_fragment_to_hps + sparse_aln, frag_map_to_adj_list vs the oracle (canonical forms)."""
import os

import numpy as np
import pytest

import orc
import pgr_tk_b200 as pg

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COMP = str.maketrans("ACGT", "TGCA")


def rand_seq(rng, L):
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=L)].tobytes()


def mutate(rng, s, rate):
    s = bytearray(s)
    for p in np.nonzero(rng.random(len(s)) < rate)[0]:
        s[p] = b"ACGT"[rng.integers(0, 4)]
    return bytes(s)


def revcomp(s):
    return s.decode().translate(COMP)[::-1].encode()


def pangenome(rng, n_hap, L, snp=0.002, with_dup=True):
    anc = rand_seq(rng, L)
    haps = []
    for h in range(n_hap):
        s = mutate(rng, anc, snp)
        if with_dup and h % 3 == 0:  # a tandem duplication and an inversion
            a = int(rng.integers(L // 4, L // 2))
            s = s[:a] + s[a - 8000:a] + s[a:]
            b = int(rng.integers(L // 2, 3 * L // 4))
            s = s[:b] + revcomp(s[b:b + 6000]) + s[b + 6000:]
        haps.append(s)
    return haps


def build_pair(haps, spec_t=(80, 56, 4, 64), mode=0):
    g = pg.ShmmrIndex(pg.ShmmrSpec(*spec_t), mode)
    o = orc.Index(orc.mkspec(*spec_t), mode)
    g.add_batch(list(range(len(haps))), haps)
    o.add_batch(list(range(len(haps))), haps)
    return g, o


def fields_equal(a, b, names):
    return all(np.array_equal(a[n], b[n]) for n in names)


def test_raw_query_matches_oracle():
    rng = np.random.default_rng(21)
    haps = pangenome(rng, 6, 120000)
    g, o = build_pair(haps)
    for qi in range(4):
        a = int(rng.integers(0, 80000))
        q = mutate(rng, haps[qi][a:a + 30000], 0.001)
        if qi % 2:
            q = revcomp(q)
        gp, goff, gh = g.raw_query(q)
        op, ooff, oh = o.raw_query(q)
        assert fields_equal(gp, op, ["h0", "h1", "bgn", "end", "ori"])
        assert np.array_equal(goff, ooff)
        assert fields_equal(gh, oh, ["frg_id", "sid", "bgn", "end", "ori"])
        assert len(gh) > 0


def assert_query_equal(g, o, queries, penalty, **kw):
    r = g.query_batch(queries, penalty, **kw)
    qto, tsid, tco, csc, cho, hits = r
    assert len(qto) == len(queries) + 1
    n_nonempty = 0
    for qi, q in enumerate(queries):
        osid, otco, osc, ocho, ohits = o.query_fragment_to_hps(q, penalty, **kw)
        t0, t1 = int(qto[qi]), int(qto[qi + 1])
        assert np.array_equal(tsid[t0:t1], osid), qi
        c0, c1 = int(tco[t0]), int(tco[t1])
        assert np.array_equal(tco[t0:t1 + 1] - tco[t0], otco)
        assert np.array_equal(csc[c0:c1].view(np.uint32), osc.view(np.uint32)), qi   # bit-exact f32 scores
        h0, h1 = int(cho[c0]), int(cho[c1])
        assert np.array_equal(cho[c0:c1 + 1] - cho[c0], ocho)
        assert fields_equal(hits[h0:h1], ohits, ["qb", "qe", "qo", "tb", "te", "to"]), qi
        n_nonempty += len(osid) > 0
    return n_nonempty


def test_query_fragment_to_hps_default_params():
    rng = np.random.default_rng(23)
    haps = pangenome(rng, 8, 150000)
    g, o = build_pair(haps)
    queries = []
    for qi in range(12):
        h = int(rng.integers(0, len(haps)))
        a = int(rng.integers(0, 100000))
        q = mutate(rng, haps[h][a:a + 20000], 0.001)
        queries.append(revcomp(q) if qi % 2 else q)
    queries += [b"", b"ACGT" * 50, rand_seq(rng, 20000)]   # no shimmers / no hits
    # pgr-query defaults (pgr-query.rs:144-164)
    n = assert_query_equal(g, o, queries, 0.025, max_count=128, max_count_query=128, max_count_target=128, max_aln_span=8)
    assert n >= 12
    # every Option = None
    assert_query_equal(g, o, queries, 0.25)


def test_query_other_callers_parameter_sets():
    rng = np.random.default_rng(29)
    haps = pangenome(rng, 6, 100000, snp=0.004)
    g, o = build_pair(haps, (48, 56, 4, 12))
    queries = [mutate(rng, haps[i % 6][5000:45000], 0.002) for i in range(6)]
    # pgr-get-sv-candidate-regions (counts 1, oriented, max_gap), pgr-web (span 0), ec.rs (32/33, oriented)
    assert_query_equal(g, o, queries, 0.05, max_count=1, max_count_query=1, max_count_target=1, max_aln_span=4, max_gap=50000, oriented=True)
    assert_query_equal(g, o, queries, 0.25, max_count=128, max_count_query=128, max_count_target=128, max_aln_span=0)
    assert_query_equal(g, o, queries, 0.1, max_count=32, max_count_query=32, max_count_target=32, max_aln_span=33, oriented=True)
    assert_query_equal(g, o, queries, 0.025, max_count=2, max_count_query=2, max_count_target=2, max_aln_span=8, max_gap=2000)


def test_sparse_aln_on_reference_test_hits():
    rows = np.loadtxt(os.path.join(GOLDEN, "test_hits"), dtype=np.int64)
    hits = np.zeros(len(rows), dtype=pg.HITPAIR)
    hits["qb"], hits["qe"], hits["qo"] = rows[:, 0], rows[:, 1], rows[:, 2]
    hits["tb"], hits["te"], hits["to"] = rows[:, 3], rows[:, 4], rows[:, 5]
    for (span, pen, gap, ori) in [(8, 0.5, None, False), (8, 0.025, None, False), (0, 0.25, None, False), (4, 0.1, 100000, True)]:
        gs, goff, gch, gsorted = pg.sparse_aln(hits, span, pen, gap, ori)
        os_, ooff, och, osorted = orc.sparse_aln(hits.copy(), span, pen, gap, ori)
        assert np.array_equal(gs.view(np.uint32), os_.view(np.uint32))
        assert np.array_equal(goff, ooff)
        assert fields_equal(gch, och, ["qb", "qe", "qo", "tb", "te", "to"])
    gs, goff, gch, _ = pg.sparse_aln(hits, 8, 0.5)
    assert len(gs) == 52 and int(np.diff(goff.astype(np.int64)).max()) == 8377
    with pytest.raises(pg.PgrError) as e:
        pg.sparse_aln(hits[:1], 8, 0.5)
    assert e.value.code == -7


def test_sparse_aln_duplicate_hit_pairs():
    # identical HitPairs collapse in the reference's maps (v_s / best_pre_v keyed by value)
    rng = np.random.default_rng(31)
    base = np.zeros(40, dtype=pg.HITPAIR)
    qb = np.sort(rng.integers(0, 100000, size=40))
    base["qb"], base["qe"] = qb, qb + rng.integers(100, 900, size=40)
    base["tb"] = base["qb"] + rng.integers(-50, 50, size=40) + 5000
    base["te"] = base["tb"] + (base["qe"] - base["qb"])
    hits = np.concatenate([base, base[5:15], base[7:9]])
    gs, goff, gch, _ = pg.sparse_aln(hits, 8, 0.05)
    os_, ooff, och, _ = orc.sparse_aln(hits.copy(), 8, 0.05)
    assert np.array_equal(gs.view(np.uint32), os_.view(np.uint32)) and np.array_equal(goff, ooff)
    assert fields_equal(gch, och, ["qb", "qe", "qo", "tb", "te", "to"])


def test_adj_list_matches_oracle():
    rng = np.random.default_rng(37)
    haps = pangenome(rng, 10, 90000, snp=0.003)
    g, o = build_pair(haps, (48, 56, 4, 12))
    names = ["sid", "ori0", "ori1", "a0", "a1", "b0", "b1"]
    for min_count, keeps in [(0, None), (2, None), (5, None), (5, [0, 3]), (100, [1]), (100, None), (3, [])]:
        ga = g.adj_list(min_count, keeps)
        oa = o.adj_list(min_count, keeps)
        assert len(ga) == len(oa), (min_count, keeps)
        assert fields_equal(ga, oa, names), (min_count, keeps)
    assert len(g.adj_list(0)) > 1000
    # fixture index
    recs = orc.parse_fasta(os.path.join(GOLDEN, "test_seqs.fa"))
    g2, o2 = build_pair([s for _, s in recs])
    assert fields_equal(g2.adj_list(0), o2.adj_list(0), names)


def test_colliding_sids_and_duplicate_signatures():
    """pgr-mdb restarts sid at 0 for every .agc file (seq_db.rs:543): per-key vectors are then not sid-monotone and can
    hold identical signatures; counts, hit expansion and the HitPair-keyed maps of sparse_aln must still agree."""
    rng = np.random.default_rng(41)
    haps = pangenome(rng, 6, 60000, snp=0.002, with_dup=False)
    g = pg.ShmmrIndex(pg.ShmmrSpec(48, 56, 4, 12), pg.FRG_ID_AGC)
    o = orc.Index(orc.mkspec(48, 56, 4, 12), 1)
    for lo in (0, 2, 4):                      # three "files", sids 0,1 each time; the last one repeats a sequence
        batch = [haps[lo], haps[lo + 1]] if lo < 4 else [haps[0], haps[5]]
        g.add_batch([0, 1], batch)
        o.add_batch([0, 1], batch)
    gk, go, gs = g.export()
    ok_, oo, os_ = o.export()
    assert np.array_equal(gk, ok_) and np.array_equal(go, oo) and fields_equal(gs, os_, ["frg_id", "sid", "bgn", "end", "ori"])
    queries = [mutate(rng, haps[0][3000:33000], 0.001), revcomp(haps[3][10000:40000]), haps[5][:20000]]
    for kw in (dict(max_count=128, max_count_query=128, max_count_target=128, max_aln_span=8),
               dict(max_count=2, max_count_query=2, max_count_target=2, max_aln_span=8),
               dict(max_count=128, max_count_query=128, max_count_target=3, max_aln_span=2, oriented=True)):
        assert_query_equal(g, o, queries, 0.025, **kw)
    names = ["sid", "ori0", "ori1", "a0", "a1", "b0", "b1"]
    for mc in (0, 2, 4):
        assert fields_equal(g.adj_list(mc), o.adj_list(mc), names)


def test_query_long_and_tiny_queries_mixed():
    rng = np.random.default_rng(43)
    haps = pangenome(rng, 5, 200000)
    g, o = build_pair(haps)
    queries = [haps[2][:150000], b"A", haps[1][1000:1400], revcomp(haps[4][50000:190000]), b"N" * 3000, haps[0][100:3000]]
    n = assert_query_equal(g, o, queries, 0.025, max_count=128, max_count_query=128, max_count_target=128, max_aln_span=8)
    assert n >= 3


def test_mdb_resident_raw_query_equals_the_loaded_index(tmp_path):
    """raw_query_fragment_from_mmap_midx (seq_db.rs:1230-1269): the key table in memory, signatures read out of the memory-mapped
    .mdb — same pairs, offsets and signatures as the index loaded into HBM and as the oracle; also on the reference's own fixture
    .mdb, whose keys are in hash-map order"""
    rng = np.random.default_rng(23)
    haps = pangenome(rng, 5, 90000)
    g, o = build_pair(haps)
    path = str(tmp_path / "t.mdb")
    g.write_mdb(path)
    m = pg.MdbMap(path)
    spec, nk, ns = m.info()
    assert (spec.w, spec.k, spec.r, spec.min_span) == (80, 56, 4, 64) and (nk, ns) == g.counts()[:2]
    for qi in range(3):
        a = int(rng.integers(0, 50000))
        q = mutate(rng, haps[qi][a:a + 25000], 0.002)
        if qi == 1:
            q = revcomp(q)
        mp, moff, mh = m.raw_query(q)
        gp, goff, gh = g.raw_query(q)
        op, ooff, oh = o.raw_query(q)
        for other_p, other_off, other_h in ((gp, goff, gh), (op, ooff, oh)):
            assert fields_equal(mp, other_p, ["h0", "h1", "bgn", "end", "ori"])
            assert np.array_equal(moff, other_off)
            assert fields_equal(mh, other_h, ["frg_id", "sid", "bgn", "end", "ori"])
        assert len(mh) > 0
    # the chaining entry point over the map (query_fragment_to_hps_from_mmap_file, ext.rs:285-342) == the batch on the loaded index
    queries = [mutate(rng, haps[i % 5][5000 * i:5000 * i + 20000], 0.002) for i in range(6)] + [b"ACGT" * 50, b""]
    for kw in (dict(max_count=128, max_count_query=128, max_count_target=128, max_aln_span=8), dict(max_aln_span=3, max_gap=5000, oriented=True), {}):
        a = m.query_batch(queries, 0.025, **kw)
        b = g.query_batch(queries, 0.025, **kw)
        assert len(a) == len(b) and all(x.tobytes() == y.tobytes() for x, y in zip(a, b)), kw
    assert len(a[1]) > 0
    m.close()
    # the reference's fixture: keys in the file's (hash-map) order, per-key vectors in file order
    import os
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    fm = pg.MdbMap(os.path.join(golden, "test_seqs_frag.mdb"))
    fi = pg.ShmmrIndex.read_mdb(os.path.join(golden, "test_seqs_frag.mdb"))
    recs = orc.parse_fasta(os.path.join(golden, "test_seqs.fa"))
    for _, s in recs[:5]:
        mp, moff, mh = fm.raw_query(s)
        ip, ioff, ih = fi.raw_query(s)
        assert fields_equal(mp, ip, ["h0", "h1", "bgn", "end", "ori"]) and np.array_equal(moff, ioff)
        assert fields_equal(mh, ih, ["frg_id", "sid", "bgn", "end", "ori"])
    fm.close()
    fi.close()
    with pytest.raises(pg.PgrError):
        pg.MdbMap(str(tmp_path / "missing.mdb"))
