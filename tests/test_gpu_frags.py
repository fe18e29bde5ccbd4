"""GPU: fragment compression (pgr_b200_index_compress_fragments: seq_to_compressed's alignment branch, match_reads,
deltas_to_aln_segs on the device, one thread per shimmer-pair row) against the reference's OWN fixture and the oracle:
  * test_seqs.fa -> every one of the 952 fragments equals what the reference stored in test_seqs_frag.frg;
  * synthetic haplotypes with SNPs, indels, inversions, tandem copies, N runs and shimmer-free sequences -> equal to
    oracle/frag_oracle.py (itself pinned to the fixture)."""
import os
import sys

import numpy as np
import pytest

import orc
import pgr_tk_b200 as pg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import frag_format as ff  # noqa: E402
from test_frag_format import load_fixture, oracle_db  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def gpu_frags(seqs, spec_t=(80, 56, 4, 64)):
    """-> the fragments as oracle-style tuples"""
    g = pg.ShmmrIndex(pg.ShmmrSpec(*spec_t), 0)
    sids = list(range(len(seqs)))
    g.add_batch(sids, seqs)
    fr, sg = g.compress_fragments(sids, seqs)
    out = []
    for f in fr:
        s = seqs[int(f["sid"])]
        if f["kind"] == ff.FRAG_ALN:
            segs = []
            for x in sg[int(f["seg_off"]):int(f["seg_off"]) + int(f["n_segs"])]:
                t = int(x["type"])
                segs.append((t,) if t == ff.SEG_FULL else ((t, int(x["a"]), int(x["b"])) if t == ff.SEG_MATCH else (t, int(x["a"]))))
            out.append((ff.FRAG_ALN, int(f["ref_frag"]), bool(f["reversed"]), int(f["len"]), segs))
        else:
            out.append((int(f["kind"]), bytes(s[int(f["bgn"]):int(f["end"])])))
    return out


def test_fixture_fragments_equal_the_reference_frg():
    _, _, _, _, ref = load_fixture()
    recs = orc.parse_fasta(os.path.join(GOLDEN, "test_seqs.fa"))
    got = gpu_frags([s for _, s in recs])
    assert len(got) == len(ref) == 952
    assert got == ref


def mutate(rng, s, snp, indel):
    out = []
    i = 0
    while i < len(s):
        r = rng.random()
        if r < snp:
            out.append(ACGT[rng.integers(0, 4)]); i += 1
        elif r < snp + indel:
            if rng.random() < 0.5:
                i += int(rng.integers(1, 6))
            else:
                out.extend(ACGT[rng.integers(0, 4, size=int(rng.integers(1, 6)))])
        else:
            out.append(s[i]); i += 1
    return np.array(out, dtype=np.uint8)


def test_synthetic_haplotypes_equal_the_oracle():
    rng = np.random.default_rng(404)
    comp = np.zeros(256, dtype=np.uint8)
    comp[[65, 67, 71, 84]] = [84, 71, 67, 65]
    anc = ACGT[rng.integers(0, 4, size=60000)]
    seqs = []
    for h in range(10):
        s = mutate(rng, anc, 0.004 * (h % 3), 0.0008 * (h % 4))
        if h % 3 == 1:
            a = int(rng.integers(10000, 20000))
            s = np.concatenate([s[:a], comp[s[a:a + 9000]][::-1], s[a + 9000:]])          # inversion: reversed fragments
        if h % 4 == 2:
            a = int(rng.integers(30000, 40000))
            s = np.concatenate([s[:a], s[a - 7000:a], s[a:]])                             # tandem copy: same key twice in one sequence
        if h == 5:
            s[25000:25040] = ord("N")
        if h == 7:
                s[5000:9000] = np.frombuffer(s[5000:9000].tobytes().lower(), dtype=np.uint8)    # lower case is valid sequence
        seqs.append(s.tobytes())
    seqs += [b"ACGT" * 20, b"", ACGT[rng.integers(0, 4, size=700)].tobytes(), seqs[0][:30000]]
    for spec_t in ((80, 56, 4, 64), (48, 56, 4, 12)):
        recs = [("s%d" % i, s) for i, s in enumerate(seqs)]
        exp = oracle_db(recs, spec_t).frags
        got = gpu_frags(seqs, spec_t)
        assert len(got) == len(exp)
        bad = [i for i, (a, b) in enumerate(zip(got, exp)) if a != b]
        assert not bad, (spec_t, bad[:5], got[bad[0]][:4], exp[bad[0]][:4])
        kinds = [f[0] for f in exp]
        assert kinds.count(ff.FRAG_ALN) > 100 and kinds.count(ff.FRAG_INTERNAL) > 50
        assert any(f[0] == ff.FRAG_ALN and f[2] for f in exp)                              # reverse-complemented alignments occur


def test_large_pageable_batch_through_the_staged_upload():
    """>= 4 MB of pageable sequences: the raw bytes reach the device through the page-locked ring (upload_raw_staged) and the
    index through the packed transport; records and segments equal the C++ fragment oracle (soft-masked stretches included:
    match_reads compares raw bytes, the index only base classes)"""
    rng = np.random.default_rng(505)
    anc = ACGT[rng.integers(0, 4, size=2_200_000)]
    seqs = []
    for h in range(3):
        s = anc.copy()
        pos = rng.integers(0, len(s), size=len(s) // 400)
        s[pos] = ACGT[rng.integers(0, 4, size=len(pos))]
        if h == 1:
            s[700_000:760_000] |= 0x20
        if h == 2:
            s = np.concatenate([s[:900_000], s[950_000:]])
        seqs.append(s.tobytes())
    assert sum(map(len, seqs)) >= 4 << 20
    spec_t = (80, 56, 4, 64)
    g = pg.ShmmrIndex(pg.ShmmrSpec(*spec_t), 0)
    g.add_batch([0, 1, 2], seqs)
    fr, sg = g.compress_fragments([0, 1, 2], seqs)
    g.close()
    efr, esg = orc.compress_fragments([0, 1, 2], seqs, orc.mkspec(*spec_t), nthreads=3)
    assert len(fr) == len(efr) and len(sg) == len(esg)
    for f in ("kind", "sid", "bgn", "end", "len", "reversed", "ref_frag", "n_segs"):
        assert np.array_equal(fr[f], efr[f]), f

    def segs_in_id_order(frags, segs):   # the segment arrays are laid out differently (index order / id order): compare per fragment
        idx = np.concatenate([np.arange(int(o), int(o) + int(c)) for o, c in zip(frags["seg_off"], frags["n_segs"]) if c] or [np.zeros(0, dtype=np.int64)])
        return segs[idx.astype(np.int64)]
    a, b = segs_in_id_order(fr, sg), segs_in_id_order(efr, esg)
    assert len(a) == len(b) == len(sg)
    for f in ("type", "a", "b"):
        assert np.array_equal(a[f], b[f]), f
