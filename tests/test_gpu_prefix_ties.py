"""GPU: directed test of the level-0 tie machinery.  The tile kernel selects window minima on a 24-bit prefix of the
key and resolves prefix ties between neighbouring candidates exactly (phase 5), including a partner that sits in the
halo of the tile.  Random sequence produces such ties about once per 10^5 minimizers, so this test PLANTS them: pairs
of 56-mers whose keys share the top 24 bits (found by hashing 8M random 56-mers) are placed 56..79 bases apart - inside
one window - all along a contig, a third of them straddling the tile boundaries (output stride 7936), plus exact
duplicates.  The level-0 list (spec r = 1, min_span = 0) must equal the oracle's."""
import numpy as np
import pytest

import orc
import pgr_tk_b200 as pg
from local_rule import u64hash

pytestmark = pytest.mark.gpu
U64 = np.uint64
K = 56
TILE_STRIDE = 7936   # 254 key blocks x 32 - 2 x 96 halo positions (shmmr_kernels.cuh)


def _brev56(v):
    out = np.zeros_like(v)
    for i in range(K):
        out |= ((v >> U64(i)) & U64(1)) << U64(K - 1 - i)
    return out


def _colliding_kmers(rng, n=8_000_000, frac=1000):
    mask = U64((1 << K) - 1)
    f0 = rng.integers(0, 1 << K, size=n, dtype=np.uint64)
    f1 = rng.integers(0, 1 << K, size=n, dtype=np.uint64)
    r0, r1 = _brev56(~f0 & mask), _brev56(~f1 & mask)
    fwd = ~(r0 < f0)
    h = u64hash(np.where(fwd, f0, r0)) ^ u64hash(np.where(fwd, f1, r1) ^ U64(0xAD12CF59))
    pre = ((h >> U64(32)) & U64(0xFFFFFF)).astype(np.int64)      # top 24 bits of x = hash << 8
    small = np.nonzero(pre < (1 << 24) // frac)[0]
    order = small[np.argsort(pre[small], kind="stable")]
    p = pre[order]
    same = np.nonzero((p[1:] == p[:-1]) & (h[order][1:] != h[order][:-1]))[0]
    return [(int(order[i]), int(order[i + 1])) for i in same], f0, f1


def _kmer_bases(f0, f1, idx):
    bits = np.arange(K - 1, -1, -1, dtype=np.uint64)             # oldest base first
    code = ((U64(f0[idx]) >> bits) & U64(1)) | (((U64(f1[idx]) >> bits) & U64(1)) << U64(1))
    return np.frombuffer(b"ACGT", dtype=np.uint8)[code.astype(np.int64)]


def test_planted_prefix_ties_match_oracle():
    rng = np.random.default_rng(2024)
    pairs, f0, f1 = _colliding_kmers(rng)
    assert len(pairs) >= 300
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    L = 2_000_000
    seq = acgt[rng.integers(0, 4, size=L)].copy()
    planted = 0
    pos = 1000
    for j, (ia, ib) in enumerate(pairs[:560]):
        delta = int(rng.integers(K, 80))                          # both k-mers inside one window of 80
        if j % 3 == 0:                                            # straddle a tile boundary of the kernel
            t = (pos // TILE_STRIDE + 1) * TILE_STRIDE
            pos = t - int(rng.integers(1, delta))
        if pos + delta + 200 > L:
            break
        a, b = (ia, ib) if j % 2 else (ib, ia)
        seq[pos - K + 1: pos + 1] = _kmer_bases(f0, f1, a)        # k-mer ending at pos
        seq[pos + delta - K + 1: pos + delta + 1] = _kmer_bases(f0, f1, b if j % 7 else a)   # every 7th: an exact duplicate
        planted += 1
        pos += int(rng.integers(300, 1500))
    assert planted >= 500
    seqs = [seq.tobytes(), seq[::-1].copy().tobytes(), seq[3000:200_000].tobytes()]
    for spec_t in ((80, 56, 1, 0), (80, 56, 4, 64)):
        got, goff = pg.get_shmmrs_from_seqs([0, 1, 2], seqs, pg.ShmmrSpec(*spec_t))
        exp, eoff = orc.shmmrs_batch([0, 1, 2], seqs, orc.mkspec(*spec_t), False, nthreads=4)
        assert list(goff) == list(eoff)
        assert np.array_equal(got, exp)
    # the planted k-mers really are selected minimizers in most cases (the test is not vacuous)
    lvl0, _ = pg.get_shmmrs_from_seqs([0], seqs[:1], pg.ShmmrSpec(80, 56, 1, 0))
    pre = (lvl0["x"] >> np.uint64(40)).astype(np.int64)
    ties = np.count_nonzero((pre[1:] == pre[:-1]) & (lvl0["x"][1:] != lvl0["x"][:-1]))
    assert ties >= 50, ties
