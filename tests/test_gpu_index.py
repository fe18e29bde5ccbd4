"""GPU parity: ShmmrFragMap build (pairs, frg_id numbering, stable sort, CSR) and .mdb I/O vs the oracle / the fixture."""
import os

import numpy as np
import pytest

import orc
import pgr_tk_b200 as pg

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rand_seq(rng, L, alphabet=b"ACGT"):
    a = np.frombuffer(alphabet, dtype=np.uint8)
    return a[rng.integers(0, len(a), size=L)].tobytes()


def related_seqs(rng, n, L, snp=0.002):
    anc = bytearray(rand_seq(rng, L))
    out = []
    for _ in range(n):
        s = bytearray(anc)
        for p in np.nonzero(rng.random(L) < snp)[0]:
            s[p] = b"ACGT"[rng.integers(0, 4)]
        out.append(bytes(s))
    return out


def assert_same_csr(gidx, oidx):
    gk, go, gs = gidx.export()
    ok, oo, os_ = oidx.export()
    assert np.array_equal(gk, ok)
    assert np.array_equal(go, oo)
    for f in ("frg_id", "sid", "bgn", "end", "ori"):
        assert np.array_equal(gs[f], os_[f]), f


def test_fixture_mdb_equal_as_map(tmp_path):
    recs = orc.parse_fasta(os.path.join(GOLDEN, "test_seqs.fa"))
    seqs = [s for _, s in recs]
    idx = pg.ShmmrIndex(pg.ShmmrSpec(80, 56, 4, 64), pg.FRG_ID_FASTX)
    idx.add_batch(list(range(len(seqs))), seqs)
    _, ref_map, _ = orc.read_mdb_py(os.path.join(GOLDEN, "test_seqs_frag.mdb"))
    assert idx.as_map() == ref_map
    assert idx.counts() == (55, 820, 886 + 66)
    # canonical .mdb is byte-identical to the oracle's canonical .mdb
    p1, p2 = str(tmp_path / "g.mdb"), str(tmp_path / "o.mdb")
    idx.write_mdb(p1)
    o = orc.Index(orc.mkspec(), 0)
    o.load_fasta(os.path.join(GOLDEN, "test_seqs.fa"))
    o.write_mdb(p2)
    assert open(p1, "rb").read() == open(p2, "rb").read()
    # reading the reference's own (hash-map ordered) fixture gives the same map
    rd = pg.ShmmrIndex.read_mdb(os.path.join(GOLDEN, "test_seqs_frag.mdb"))
    assert rd.as_map() == ref_map
    s = rd.spec()
    assert (s.w, s.k, s.r, s.min_span, s.sketch) == (80, 56, 4, 64, 0)


@pytest.mark.parametrize("mode", [0, 1])
def test_incremental_batches_and_modes(mode):
    rng = np.random.default_rng(3)
    seqs = related_seqs(rng, 12, 60000) + [b"", b"ACGT" * 10, rand_seq(rng, 500)] + related_seqs(rng, 5, 30000)
    spec = pg.ShmmrSpec(48, 56, 4, 12)
    g = pg.ShmmrIndex(spec, mode)
    o = orc.Index(orc.mkspec(48, 56, 4, 12), mode)
    sid = 0
    for lo, hi in [(0, 7), (7, 15), (15, 20)]:
        sids = list(range(sid, sid + hi - lo)) if mode == 0 else list(range(hi - lo))  # pgr-mdb restarts sids (seq_db.rs:543)
        g.add_batch(sids, seqs[lo:hi])
        o.add_batch(sids, seqs[lo:hi])
        sid += hi - lo
    assert_same_csr(g, o)
    nk, ns, nf = g.counts()
    assert ns > 1000 and nk < ns


def test_stage_commit_equals_add_batch():
    rng = np.random.default_rng(5)
    seqs = related_seqs(rng, 6, 40000)
    spec = pg.ShmmrSpec()
    a = pg.ShmmrIndex(spec, 0)
    a.add_batch([0, 1, 2], seqs[:3])
    a.add_batch([3, 4, 5], seqs[3:])
    b = pg.ShmmrIndex(spec, 0)
    n0 = b.stage_batch([0, 1, 2], seqs[:3])
    b.commit_batch(0)
    n1 = b.stage_batch([3, 4, 5], seqs[3:])
    b.commit_batch(n0)
    assert a.counts() == b.counts() and a.counts()[2] == n0 + n1
    ka, oa, sa = a.export()
    kb, ob, sb = b.export()
    assert np.array_equal(ka, kb) and np.array_equal(oa, ob) and np.array_equal(sa, sb)


def test_mdb_roundtrip_and_partition(tmp_path):
    rng = np.random.default_rng(9)
    seqs = related_seqs(rng, 8, 80000)
    g = pg.ShmmrIndex(pg.ShmmrSpec(), 0)
    g.add_batch(list(range(8)), seqs)
    ref = g.as_map()
    p = str(tmp_path / "x.mdb")
    g.write_mdb(p)
    r = pg.ShmmrIndex.read_mdb(p)
    assert r.as_map() == ref
    # stable partition by key range keeps the map intact and reports consistent counts
    keys, _, _ = g.export()
    sp = [int(keys[len(keys) // 3, 0]), int(keys[2 * len(keys) // 3, 0])]
    counts = g.partition(sp)
    assert int(counts.sum()) == g.counts()[1]
    assert g.as_map() == ref
