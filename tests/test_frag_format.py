"""CPU: the fragment store (.sdx/.frg) codec and the fragment-compression oracle against the REFERENCE'S OWN FIXTURE
(pgr-db/test/test_data/test_seqs_frag.{sdx,frg}, written by the reference from test_seqs.fa with 80/56/4/64):
  * the fixture decodes (bincode 2 standard config, raw deflate per 256-fragment chunk) to 952 fragments that reconstruct
    exactly the 66 sequences of test_seqs.fa (seq_db.rs:685-735);
  * re-encoding the decoded chunks gives the inflated payloads back byte for byte;
  * oracle/frag_oracle.py (match_reads, deltas_to_aln_segs, seq_to_compressed restated) reproduces every one of the 952
    fragments - alignment segments, base fragment ids, orientation flags - and the .sdx sequence table and chunk lengths."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import frag_format as ff  # noqa: E402
import frag_oracle as fo  # noqa: E402

import orc  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_fixture():
    cs, addr, seqs = ff.read_sdx(os.path.join(GOLDEN, "test_seqs_frag.sdx"))
    payloads = ff.read_frg_chunks(os.path.join(GOLDEN, "test_seqs_frag.frg"), addr)
    return cs, addr, seqs, payloads, ff.decode_chunks(payloads)


def oracle_db(recs, spec_t=(80, 56, 4, 64), source="test_seqs.fa"):
    spec = orc.mkspec(*spec_t)
    db = fo.CompactSeqDB(spec_t[1])
    for sid, (name, seq) in enumerate(recs):
        mm, _ = orc.shmmrs_batch([sid], [seq], spec, False)
        sh = [(int(x) >> 8, (int(y) & 0xFFFFFFFF) >> 1) for x, y in zip(mm["x"], mm["y"])]
        db.seq_to_compressed(source, name, sid, seq, sh)
    return db


def test_fixture_decodes_to_the_fasta():
    cs, addr, seqs, payloads, frags = load_fixture()
    recs = orc.parse_fasta(os.path.join(GOLDEN, "test_seqs.fa"))
    assert cs == 256 and len(addr) == 4 and len(seqs) == 66 and len(frags) == 952
    kinds = [f[0] for f in frags]
    assert kinds.count(ff.FRAG_ALN) == 633 and kinds.count(ff.FRAG_INTERNAL) == 187 and kinds.count(ff.FRAG_PREFIX) == 66
    for i, s in enumerate(seqs):
        assert s["id"] == i and s["name"] == recs[i][0] and s["len"] == len(recs[i][1]) and s["source"] == "test_seqs.fa"
        assert ff.get_seq(frags, 56, s) == recs[i][1]
    for i, p in enumerate(payloads):
        assert ff.enc_chunk(frags[i * cs:(i + 1) * cs]) == p
    # chunk table: (offset, compressed length, bases held) - seq_db.rs:832-860
    off = 0
    for i, (o, ln, bases) in enumerate(addr):
        assert o == off
        off += ln
        tot = 0
        for f in frags[i * cs:(i + 1) * cs]:
            tot += (f[3] - 56) if f[0] == ff.FRAG_ALN else (len(f[1]) - 56 if f[0] == ff.FRAG_INTERNAL else len(f[1]))
        assert tot == bases


def test_oracle_reproduces_the_fixture_fragments():
    _, _, seqs, _, frags = load_fixture()
    recs = orc.parse_fasta(os.path.join(GOLDEN, "test_seqs.fa"))
    db = oracle_db(recs)
    assert len(db.frags) == len(frags)
    assert db.frags == frags
    assert [(s["seq_frag_range"], s["len"], s["name"]) for s in db.seqs] == [(s["seq_frag_range"], s["len"], s["name"]) for s in seqs]


def test_match_reads_known_cases():
    a = b"ACGTACGTTTGACCAGTAGGATCCATTAGACCAGGATTTACCAGGGATTTAGGACCATAGGACCCATTTAG" * 3
    m = fo.match_reads(a, a, True, 0.1, 0, 0, 32)
    assert m["deltas"] == [] and m["end0"] == len(a) and m["end1"] == len(a)
    assert fo.deltas_to_aln_segs(m["deltas"], m["end0"], m["end1"], a, a) == [(ff.SEG_FULL,)]
    b = a[:50] + b"T" + a[50:120] + a[123:]                 # one insertion, one 3-base deletion
    m = fo.match_reads(a, b, True, 0.1, 0, 0, 32)
    segs = fo.deltas_to_aln_segs(m["deltas"], m["end0"], m["end1"], a, b)
    assert ff.reconstruct_from_segs(a, segs) == b
    assert fo.match_reads(a, bytes(reversed(a)), True, 0.1, 0, 0, 32) is None   # too divergent for d_max / the band


def test_oracle_round_trip_on_mutated_haplotypes():
    """size-independent property: whatever the alignments decide, every sequence is reconstructed from the store"""
    import numpy as np
    rng = np.random.default_rng(9)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = np.zeros(256, dtype=np.uint8)
    comp[[65, 67, 71, 84]] = [84, 71, 67, 65]
    anc = acgt[rng.integers(0, 4, size=40000)]
    recs = []
    for h in range(8):
        s = anc.copy()
        m = np.nonzero(rng.random(len(s)) < 0.003 * (1 + h % 3))[0]
        s[m] = acgt[rng.integers(0, 4, size=len(m))]
        for _ in range(h):                                                    # a few indels
            a = int(rng.integers(1000, len(s) - 1000))
            s = np.concatenate([s[:a], s[a + int(rng.integers(1, 30)):]]) if rng.random() < 0.5 else np.concatenate([s[:a], acgt[rng.integers(0, 4, size=int(rng.integers(1, 30)))], s[a:]])
        if h % 2:
            a = int(rng.integers(5000, 15000))
            s = np.concatenate([s[:a], comp[s[a:a + 6000]][::-1], s[a + 6000:]])
        recs.append(("h%d" % h, s.tobytes()))
    recs.append(("tiny", b"ACGTTGCA" * 10))
    for spec_t in ((80, 56, 4, 64), (48, 56, 4, 12), (24, 24, 2, 8)):
        db = oracle_db(recs, spec_t, source="mem")
        assert all(ff.get_seq(db.frags, spec_t[1], s) == recs[i][1] for i, s in enumerate(db.seqs))
        kinds = [f[0] for f in db.frags]
        assert kinds.count(ff.FRAG_ALN) > 20
        # chunk encoding round trip
        pay = ff.enc_chunk(db.frags[:256])
        assert ff.decode_chunks([pay]) == db.frags[:256]


def _records_to_tuples(fr, sg, seqs):
    out = []
    for f in fr:
        if f["kind"] == ff.FRAG_ALN:
            segs = []
            for x in sg[int(f["seg_off"]):int(f["seg_off"]) + int(f["n_segs"])]:
                t = int(x["type"])
                segs.append((t,) if t == ff.SEG_FULL else ((t, int(x["a"]), int(x["b"])) if t == ff.SEG_MATCH else (t, int(x["a"]))))
            out.append((ff.FRAG_ALN, int(f["ref_frag"]), bool(f["reversed"]), int(f["len"]), segs))
        else:
            out.append((int(f["kind"]), bytes(seqs[int(f["sid"])][int(f["bgn"]):int(f["end"])])))
    return out


def test_cpp_fragment_oracle_equals_the_fixture_and_the_python_oracle():
    """oracle/frag_oracle.cpp (the restatement that runs at benchmark sizes / the CPU baseline) against the reference's fixture"""
    _, _, _, _, frags = load_fixture()
    recs = orc.parse_fasta(os.path.join(GOLDEN, "test_seqs.fa"))
    seqs = [s for _, s in recs]
    for nthreads in (1, 4):
        fr, sg = orc.compress_fragments(list(range(len(seqs))), seqs, orc.mkspec(80, 56, 4, 64), nthreads=nthreads)
        assert _records_to_tuples(fr, sg, seqs) == frags
    # another spec, short / shimmer-free sequences
    seqs2 = seqs[:12] + [b"ACGT" * 30, b"", seqs[3][:900]]
    fr, sg = orc.compress_fragments(list(range(len(seqs2))), seqs2, orc.mkspec(48, 56, 4, 12), nthreads=2)
    exp = oracle_db([("s%d" % i, s) for i, s in enumerate(seqs2)], (48, 56, 4, 12)).frags
    assert _records_to_tuples(fr, sg, seqs2) == exp


def test_row_parallel_scheme_equals_the_sequential_decisions():
    """frags.cu decides a shimmer-pair row in two steps - every entry against the row's first entry in parallel, then the
    entries that failed there sequentially - instead of the reference's one sequential walk (seq_db.rs:258-318).  The two
    are the same function of (row order, sequence ids, fragment lengths, which pairs align): property test on random rows
    with an arbitrary `aligns(i, j)` relation."""
    import numpy as np
    rng = np.random.default_rng(31)
    for trial in range(300):
        m = int(rng.integers(1, 14))
        sid = np.sort(rng.integers(0, 6, size=m))
        long_ = rng.random(m) < 0.8                       # frg_len > 128
        ok = rng.random((m, m)) < rng.choice([0.2, 0.6, 0.95])

        def sequential():
            kind, ref = ["I"] * m, [None] * m
            for j in range(m):
                if not long_[j]:
                    continue
                for i in range(j):
                    if sid[i] >= sid[j]:                  # frag_map holds earlier sequences only
                        break
                    if kind[i] != "I":
                        continue
                    if ok[i, j]:
                        kind[j], ref[j] = "A", i
                        break
            return kind, ref

        def two_step():
            kind, ref = ["I"] * m, [None] * m
            for j in range(1, m):                         # pass 0a: independent of each other
                if long_[j] and sid[0] < sid[j]:
                    if ok[0, j]:
                        kind[j], ref[j] = "A", 0
                    else:
                        kind[j] = "P"
            for j in range(1, m):                         # pass 0b: the pending ones, in order
                if kind[j] != "P":
                    continue
                kind[j] = "I"
                for i in range(1, j):
                    if sid[i] >= sid[j]:
                        break
                    if kind[i] != "I":
                        continue
                    if ok[i, j]:
                        kind[j], ref[j] = "A", i
                        break
            return kind, ref
        assert sequential() == two_step(), trial
