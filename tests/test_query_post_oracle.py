"""CPU checks of oracle/query_post_oracle.py (pgr-query.rs:166-430 restated) on hand-derived cases."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import query_post_oracle as q  # noqa: E402


def hp(qb, qe, qo, tb, te, to):
    return ((qb, qe, qo), (tb, te, to))


def test_merge_and_orientation_rules():
    fwd1 = [hp(0, 10, 0, 100, 110, 0), hp(10, 20, 0, 110, 120, 0), hp(20, 30, 0, 120, 130, 0)]
    fwd2 = [hp(40, 50, 0, 5000, 5010, 0), hp(50, 60, 0, 5010, 5020, 0), hp(60, 70, 0, 5020, 5030, 0)]
    short = [hp(1, 2, 0, 3, 4, 0), hp(2, 3, 0, 4, 5, 0)]                       # <= 2 anchors: dropped
    rev = [hp(80, 90, 0, 900, 910, 1), hp(90, 100, 0, 890, 900, 1), hp(100, 110, 0, 880, 890, 1), hp(110, 120, 0, 870, 880, 1),
           hp(120, 130, 0, 860, 870, 1), hp(130, 140, 0, 850, 860, 1), hp(140, 150, 0, 840, 850, 1)]
    targets = [(7, [(3.0, rev)]), (3, [(10.0, fwd1), (5.0, fwd2), (1.0, short)])]
    m = q.merge_query_hits(targets, 100000)
    assert [sid for sid, _ in m] == [3, 7]                                        # canonical: ascending sid
    assert [(r[0], r[1], r[2], r[3], len(r[4])) for r in m[0][1]] == [(100, 5030, 4930, 0, 6)]
    assert [(r[0], r[1], r[2], r[3], len(r[4])) for r in m[1][1]] == [(840, 910, 70, 1, 7)]
    m = q.merge_query_hits(targets, 1000)
    assert [(r[0], r[1]) for r in m[0][1]] == [(100, 130), (5000, 5030)]
    # the orientation counters run across the chains of one target (pgr-query.rs:170-171): after 7 reverse anchors a
    # forward chain of 3 is still labelled reverse
    m = q.merge_query_hits([(1, [(9.0, rev), (2.0, fwd1)])], 10)
    assert [(r[0], r[3]) for r in m[0][1]] == [(100, 1), (840, 1)]


def test_names_and_lines():
    assert q.file_stem("/a/b/test_seqs.fa") == "test_seqs" and q.file_stem("x.fa.gz") == "x.fa" and q.file_stem("/p/.hidden") == ".hidden"
    assert q.reverse_complement(b"ACGTNacgtn-") == b"-nacgtNACGT"
    m = [(2, [(5, 25, 20, 1, [hp(30, 40, 0, 15, 25, 1), hp(10, 20, 0, 5, 15, 1), hp(20, 30, 0, 9, 19, 1)])])]
    text, subs = q.hit_lines(4, "qname", 1234, m, {2: ("ctgA", "/data/hap.fa.gz")})
    assert text.splitlines()[1] == "004\tqname\t10\t40\t1234\t3\t/data/hap.fa.gz\tctgA\t5\t25\t1\thap.fa::ctgA_5_25_1"
    assert subs == [("hap.fa::ctgA_5_25_1", 2, 5, 25, 1)]
    text, _ = q.hit_lines(4, "qname", 1234, m, {2: ("ctgA", "/data/hap.fa.gz")}, bed=True)
    assert text.splitlines()[1] == "ctgA\t5\t25\tqname\t#AAAAAA\t1\t1234\t3\t4\t/data/hap.fa.gz\t10\t40\thap.fa::ctgA_5_25_1"
