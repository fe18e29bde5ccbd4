"""Property test: the local (stateless) rule the CUDA kernels implement == the reference state machine (oracle),
whenever no pushed position holds a reverse-complement palindrome (those route to the sequential replay kernel)."""
import os

import numpy as np
import pytest

import local_rule as lr
import orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rand_seq(rng, L, alphabet):
    a = np.frombuffer(alphabet, dtype=np.uint8)
    return a[rng.integers(0, len(a), size=L)].tobytes()


def tandem(rng, L, unit_len, noise):
    unit = rand_seq(rng, unit_len, b"ACGT")
    s = bytearray((unit * (L // unit_len + 1))[:L])
    for i in range(L):
        if rng.random() < noise:
            s[i] = b"ACGT"[rng.integers(0, 4)]
    return bytes(s)


def check(seq, w, k, r, ms, padding):
    spec = orc.mkspec(w, k, r, ms, False)
    ref = orc.sequence_to_shmmrs(7, seq, spec, padding)
    xs, ys, pal = lr.sequence_to_shmmrs_local(7, seq, w, k, r, ms, padding)
    if pal:
        return False
    assert list(ref["x"]) == list(xs), (w, k, r, ms, padding, len(seq))
    assert list(ref["y"]) == list(ys), (w, k, r, ms, padding, len(seq))
    return True


def test_local_rule_matches_state_machine_random():
    rng = np.random.default_rng(12345)
    n_ok = 0
    ks = [5, 7, 8, 11, 16, 24, 31, 56]
    ws = [1, 2, 3, 4, 8, 12, 24, 48, 80, 128]
    rs = [1, 2, 3, 4, 6, 12]
    mss = [0, 4, 12, 24, 64]
    alphabets = [b"ACGT", b"AC", b"ACGTN", b"ACGTacgtNn-", b"A"]
    for it in range(1500):
        k = ks[rng.integers(len(ks))]
        w = ws[rng.integers(len(ws))]
        r = rs[rng.integers(len(rs))]
        ms = mss[rng.integers(len(mss))]
        edge = [0, 1, k - 1, k, k + 1, k + 2, w + k - 1, w + k, w + k + 1, 2 * w + k, 2 * w + k - 1, 3 * w + k]
        L = edge[rng.integers(len(edge))] if rng.random() < 0.3 else int(rng.integers(100, 3000))
        if rng.random() < 0.15:
            seq = tandem(rng, L, int(rng.integers(1, 40)), 0.02)
        else:
            seq = rand_seq(rng, L, alphabets[rng.integers(len(alphabets))])
        n_ok += check(seq, w, k, r, ms, bool(rng.integers(2)))
    assert n_ok > 700  # palindrome-bearing cases are skipped, the rest must match


def test_local_rule_on_fixture():
    recs = orc.parse_fasta(os.path.join(GOLDEN, "test_seqs.fa"))
    n = 0
    for w, k, r, ms in [(80, 56, 4, 64), (48, 56, 4, 12), (24, 24, 12, 24), (31, 31, 1, 0)]:
        for _, s in recs[:12]:
            n += check(s, w, k, r, ms, False)
    assert n >= 40


def test_sketch_local():
    rng = np.random.default_rng(5)
    for it in range(200):
        k = [8, 16, 24, 56][rng.integers(4)]
        r = [1, 2, 4][rng.integers(3)]
        ms = [0, 12, 64][rng.integers(3)]
        seq = rand_seq(rng, int(rng.integers(0, 3000)), [b"ACGT", b"ACGTN", b"AC"][rng.integers(3)])
        ref = orc.sequence_to_shmmrs(3, seq, orc.mkspec(80, k, r, ms, True))
        xs, ys = lr.sketch_local(3, seq, k, r, ms)
        assert list(ref["x"]) == list(xs) and list(ref["y"]) == list(ys)


def test_prefix_candidates_plus_tie_rule_equal_exact_selection():
    """The tile kernel selects on a PREFIX of the key (24 bits) and repairs the result with one rule (phase 5): a
    candidate (selected on the prefixes) that has no other candidate with the same prefix within distance < w is
    selected for sure; one that has is tested exactly.  This is that argument as a property test on the numpy model,
    with prefixes short enough (3-6 bits) that ties are everywhere:
      * every exactly selected position is a prefix candidate,
      * every prefix candidate without a same-prefix candidate within w-1 positions is exactly selected,
    so testing only the tied candidates exactly gives the exact selection."""
    rng = np.random.default_rng(77)
    for trial in range(60):
        n = int(rng.integers(50, 600))
        w = int(rng.integers(2, 130))
        bits = int(rng.integers(3, 7))
        x = rng.integers(0, 1 << 20, size=n, dtype=np.int64)
        if trial % 3 == 0:                               # runs of equal keys (tandem repeats)
            x[rng.integers(0, n, size=n // 3)] = x[rng.integers(0, n)]
        if n < w:
            continue
        pre = x >> (20 - bits)
        exact = lr.window_select(x, w)
        cand = lr.window_select(pre, w)
        assert not (exact & ~cand).any()                 # candidates are a superset
        idx = np.nonzero(cand)[0]
        tied = np.zeros(n, dtype=bool)
        for a, i in enumerate(idx):                      # neighbouring candidates within distance < w with the same prefix
            for j in idx[max(0, a - w):a + w + 1]:
                if j != i and abs(int(j) - int(i)) < w and pre[j] == pre[i]:
                    tied[i] = True
                    break
        untied = cand & ~tied
        assert not (untied & ~exact).any()               # an untied candidate is exactly selected
        repaired = untied | (tied & exact)               # tied candidates get the exact test
        assert np.array_equal(repaired, exact)
