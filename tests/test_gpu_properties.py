"""GPU: size-independent properties at sizes the oracle does not finish in seconds (hundreds of Mbases):
ordering, min_span, strand/rid fields, idempotence, host-batch (chunked, overlapped) == device-resident, and an exact
oracle comparison on a sample of the sequences."""
import numpy as np
import pytest

import orc
import pgr_tk_b200 as pg

pytestmark = pytest.mark.gpu


def big_batch(seed, n, L):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    return [acgt[rng.integers(0, 4, size=L, dtype=np.uint8)] for _ in range(n)]


@pytest.mark.parametrize("spec_t", [(80, 56, 4, 64), (48, 56, 4, 12)])
def test_large_batch_properties(spec_t):
    seqs = big_batch(2024, 48, 5_000_000)          # 240 Mbases: several chunks of the overlapped host path
    w, k, r, ms = spec_t
    spec = pg.ShmmrSpec(*spec_t)
    rids = list(range(100, 100 + len(seqs)))
    mm, off = pg.get_shmmrs_from_seqs(rids, seqs, spec)
    # device-resident pipeline gives the same bytes (and is idempotent)
    ctx = pg.Ctx(0)
    ctx.upload(seqs, rids)
    n1 = ctx.shmmrs(spec)
    a, aoff = ctx.shmmrs_download()
    n2 = ctx.shmmrs(spec)
    b, boff = ctx.shmmrs_download()
    ctx.close()
    assert n1 == n2 == len(mm)
    assert np.array_equal(a, mm) and np.array_equal(aoff, off) and np.array_equal(a, b) and np.array_equal(aoff, boff)
    pos = ((mm["y"] & np.uint64(0xFFFFFFFF)) >> np.uint64(1)).astype(np.int64)
    rid = (mm["y"] >> np.uint64(32)).astype(np.int64)
    assert np.all((mm["x"] & np.uint64(0xFF)) == k)                       # MM128.span() is always k
    for i in range(len(seqs)):
        s, e = int(off[i]), int(off[i + 1])
        assert e - s > 5_000_000 // 1000
        assert np.all(rid[s:e] == rids[i])
        p = pos[s:e]
        assert np.all(np.diff(p) > 0) and p[0] >= k and p[-1] < 5_000_000
        # min_span filter (shmmrutils.rs:536-555): an interior shimmer is farther than min_span from both neighbours
        if len(p) >= 3:   # every gap touches an interior shimmer, and interior shimmers keep both gaps > min_span
            assert np.all(np.diff(p) > ms)
    # density sanity (random ACGT): 1 per ~329 bases for 80/56/4/64, ~1 per 140 for 48/56/4/12
    dens = 240_000_000 / len(mm)
    assert (250 < dens < 420) if w == 80 else (100 < dens < 190)
    # exact oracle comparison on a sample of the sequences
    for i in (0, 17, 47):
        exp = orc.sequence_to_shmmrs(rids[i], seqs[i], orc.mkspec(*spec_t))
        assert np.array_equal(mm[int(off[i]):int(off[i + 1])], exp)


def test_one_long_contig_matches_oracle():
    """a single 60 Mb contig (the reference processes it with one thread; here ~15.6k tiles over 740 CTAs)"""
    seq = big_batch(7, 1, 60_000_000)[0]
    spec = pg.ShmmrSpec()
    got = pg.sequence_to_shmmrs(5, seq, spec)
    exp = orc.sequence_to_shmmrs(5, seq, orc.mkspec())
    assert np.array_equal(got, exp)
