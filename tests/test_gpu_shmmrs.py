"""GPU parity: sequence_to_shmmrs through the C ABI vs the CPU oracle (bit-exact MM128 lists)."""
import os

import numpy as np
import pytest

import orc
import pgr_tk_b200 as pg

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ospec(s):
    return orc.mkspec(s.w, s.k, s.r, s.min_span, bool(s.sketch))


def rand_seq(rng, L, alphabet=b"ACGT"):
    a = np.frombuffer(alphabet, dtype=np.uint8)
    return a[rng.integers(0, len(a), size=L)].tobytes()


def assert_batch_equal(seqs, spec, padding=False, rids=None):
    rids = list(range(len(seqs))) if rids is None else rids
    got, goff = pg.get_shmmrs_from_seqs(rids, seqs, spec, padding)
    exp, eoff = orc.shmmrs_batch(rids, seqs, ospec(spec), padding, nthreads=8)
    assert list(goff) == list(eoff), (spec.w, spec.k, spec.r, spec.min_span, padding)
    assert np.array_equal(got["x"], exp["x"])
    assert np.array_equal(got["y"], exp["y"])
    return len(got)


def test_fixture_sequences_default_spec():
    recs = orc.parse_fasta(os.path.join(GOLDEN, "test_seqs.fa"))
    n = assert_batch_equal([s for _, s in recs], pg.ShmmrSpec(80, 56, 4, 64))
    assert n == 886


def test_fixture_sequences_other_specs():
    recs = orc.parse_fasta(os.path.join(GOLDEN, "test_seqs.fa"))
    seqs = [s for _, s in recs]
    for w, k, r, ms in [(48, 56, 4, 12), (24, 24, 12, 24), (31, 31, 1, 0), (80, 56, 1, 0), (128, 56, 2, 10), (33, 16, 3, 5), (97, 41, 5, 0)]:
        assert_batch_equal(seqs, pg.ShmmrSpec(w, k, r, ms))


def test_single_sequence_api_and_rid():
    recs = orc.parse_fasta(os.path.join(GOLDEN, "test_seqs.fa"))
    spec = pg.ShmmrSpec()
    got = pg.sequence_to_shmmrs(1234, recs[3][1], spec)
    exp = orc.sequence_to_shmmrs(1234, recs[3][1], ospec(spec))
    assert np.array_equal(got, exp)
    assert np.all((got["y"] >> 32) == 1234)


def test_boundary_known_answer_padding():
    recs = orc.parse_fasta(os.path.join(GOLDEN, "boundary_seqs.fa"))
    spec = pg.ShmmrSpec(24, 24, 12, 24)
    for _, s in recs:
        got = pg.sequence_to_shmmrs(0, s, spec, padding=True)
        assert len(got) == 2
        assert np.array_equal(got, orc.sequence_to_shmmrs(0, s, ospec(spec), True))


def test_rc_match_sketch():
    recs = orc.parse_fasta(os.path.join(GOLDEN, "test_rev.fa"))
    for sketch in (True, False):
        spec = pg.ShmmrSpec(80, 56, 4, 64, sketch)
        s0 = pg.sequence_to_shmmrs(0, recs[0][1], spec)
        s1 = pg.sequence_to_shmmrs(0, recs[1][1], spec)
        assert len(s0) > 0
        assert list(s0["x"] >> 8) == list((s1["x"] >> 8)[::-1])
        assert np.array_equal(s0, orc.sequence_to_shmmrs(0, recs[0][1], ospec(spec)))


def test_long_random_contigs_multi_tile():
    rng = np.random.default_rng(1)
    seqs = [rand_seq(rng, L) for L in (1_000_000, 7936, 7937, 8128, 15872, 15873, 200_003, 31, 56, 57, 135, 136, 137, 0, 1)]
    for spec in (pg.ShmmrSpec(80, 56, 4, 64), pg.ShmmrSpec(48, 56, 4, 12)):
        assert_batch_equal(seqs, spec)


def test_random_specs_edge_lengths():
    rng = np.random.default_rng(7)
    ks = [5, 7, 8, 11, 16, 24, 31, 56]
    ws = [1, 2, 3, 8, 24, 32, 33, 48, 64, 65, 80, 96, 97, 128]
    for it in range(40):
        k = ks[rng.integers(len(ks))]
        w = ws[rng.integers(len(ws))]
        r = int(rng.integers(1, 13))
        ms = int([0, 4, 12, 64][rng.integers(4)])
        edge = [0, 1, k - 1, k, k + 1, w + k - 1, w + k, w + k + 1, 2 * w + k, 3 * w + k, 3 * w + 31, 3 * w + 33]
        seqs = [rand_seq(rng, L) for L in edge] + [rand_seq(rng, int(rng.integers(100, 30000))) for _ in range(6)]
        assert_batch_equal(seqs, pg.ShmmrSpec(w, k, r, ms), padding=bool(rng.integers(2)))


def test_low_complexity_and_ties():
    rng = np.random.default_rng(11)
    seqs = [b"A" * 5000, b"AC" * 4000, b"ACG" * 3000, rand_seq(rng, 20000, b"AC"),
            (rand_seq(rng, 37) * 400), rand_seq(rng, 3000) + b"T" * 3000 + rand_seq(rng, 3000)]
    for w, k, r, ms in [(80, 56, 4, 64), (48, 56, 4, 12), (80, 31, 4, 0), (33, 7, 2, 0), (16, 5, 3, 0)]:
        assert_batch_equal(seqs, pg.ShmmrSpec(w, k, r, ms))


def test_invalid_bytes_and_palindromes_take_replay_path():
    rng = np.random.default_rng(13)
    s1 = bytearray(rand_seq(rng, 50000))
    s1[20000:20100] = b"N" * 100
    s2 = rand_seq(rng, 10000) + b"AT" * 60 + rand_seq(rng, 10000)   # (AT)n >= 56: reverse-complement palindromes
    s3 = rand_seq(rng, 30000, b"ACGTacgtNn")
    s4 = bytes([0, 1, 2, 3]) * 2000
    seqs = [bytes(s1), s2, s3, s4, rand_seq(rng, 40000)]
    for spec in (pg.ShmmrSpec(80, 56, 4, 64), pg.ShmmrSpec(24, 24, 12, 24), pg.ShmmrSpec(80, 56, 4, 64, True)):
        assert_batch_equal(seqs, spec)
    ctx = pg.Ctx(0)
    ctx.upload(seqs)
    ctx.shmmrs(pg.ShmmrSpec())
    c = ctx.counters()
    assert c[2] == 0   # nothing needs the whole-sequence replay any more
    assert c[4] >= 4   # invalid bytes and the (AT)n palindromes are handled by local patches
    ctx.close()


def test_invalid_byte_patches():
    """bytes outside ACGTacgt do not update the k-mer registers but the stale k-mer is still pushed (shmmrutils.rs:461-476)"""
    rng = np.random.default_rng(107)
    r = lambda L, a=b"ACGT": rand_seq(rng, L, a)
    seqs = [
        r(30000) + b"N" * 5000 + r(30000),                 # one long N run
        b"N" * 300 + r(9000) + b"N" * 200,                 # at both ends
        b"N" * 4000,                                       # nothing valid at all
        b"NNNN" + r(40) + b"N" + r(30) + b"NN" + r(5000),  # fewer than k valid bases before position k
        r(8000) + b"n" + r(8000) + b"-" + r(31) + b"R" + r(8000),          # isolated single bytes
        bytes([0, 1, 2, 3]) * 1500 + r(3000),              # raw codes 0..3 are valid for the reference's LUT
        r(20000, b"ACGTacgtN"),                            # dense N (every ~9th byte)
        r(20000, b"ACGT" * 30 + b"N"),                     # sparse N (every ~120th byte)
        b"AT" * 200 + b"N" * 10 + b"AT" * 200 + r(4000),   # palindromes next to N
        r(50000),                                          # clean
    ]
    for w, k, r_, ms in [(80, 56, 4, 64), (48, 56, 4, 12), (24, 24, 12, 24), (8, 12, 2, 0), (128, 31, 3, 7), (5, 7, 1, 0)]:
        for padding in (False, True):
            assert_batch_equal(seqs, pg.ShmmrSpec(w, k, r_, ms), padding=padding)
    ctx = pg.Ctx(0)
    ctx.upload(seqs)
    ctx.shmmrs(pg.ShmmrSpec())
    c = ctx.counters()
    assert c[2] == 0 and c[4] >= 8
    ctx.close()


COMP = bytes.maketrans(b"ACGT", b"TGCA")


def revcomp(s):
    return s.translate(COMP)[::-1]


def test_palindrome_patches_inverted_repeats_and_runs():
    """pushed positions with fmmer == rmmer (shmmrutils.rs:477) are re-derived by an exact local replay"""
    rng = np.random.default_rng(101)
    u = rand_seq(rng, 4000)
    seqs = [
        rand_seq(rng, 20000) + u + revcomp(u) + rand_seq(rng, 20000),                 # perfect inverted repeat: one palindromic centre
        rand_seq(rng, 300) + u + revcomp(u) + rand_seq(rng, 100),                     # near both ends
        u[:200] + revcomp(u[:200]),                                                   # whole sequence is one palindrome
        rand_seq(rng, 9000) + b"AT" * 400 + rand_seq(rng, 9000),                      # long run of consecutive skips
        b"AT" * 300 + rand_seq(rng, 5000) + b"TA" * 100,                              # runs at the very start / end
        b"".join(rand_seq(rng, int(rng.integers(150, 900))) + x + revcomp(x) for x in [rand_seq(rng, 70) for _ in range(40)]),  # many clusters
        rand_seq(rng, 30000),                                                         # clean
    ]
    for w, k, r, ms in [(80, 56, 4, 64), (48, 56, 4, 12), (24, 24, 12, 24), (128, 56, 2, 0), (33, 40, 3, 5), (80, 56, 1, 0)]:
        assert_batch_equal(seqs, pg.ShmmrSpec(w, k, r, ms))
    ctx = pg.Ctx(0)
    ctx.upload(seqs)
    ctx.shmmrs(pg.ShmmrSpec())
    c = ctx.counters()
    assert c[2] == 0 and c[4] >= 6
    ctx.close()


def test_palindrome_patches_small_even_k_random():
    """even small k makes palindromic k-mers frequent on random sequence: clusters merge, replays run long"""
    rng = np.random.default_rng(103)
    for it in range(12):
        k = [6, 8, 10, 12, 16][rng.integers(5)]
        w = [4, 8, 24, 33, 48, 80, 128][rng.integers(7)]
        r = int(rng.integers(1, 6))
        seqs = [rand_seq(rng, int(rng.integers(50, 40000))) for _ in range(6)] + [rand_seq(rng, 9000, b"AT"), rand_seq(rng, 9000, b"ACGTTTTAAA")]
        assert_batch_equal(seqs, pg.ShmmrSpec(w, k, r, int([0, 5, 30][rng.integers(3)])), padding=bool(rng.integers(2)))


def test_sketch_mode_random():
    rng = np.random.default_rng(17)
    seqs = [rand_seq(rng, int(L)) for L in (0, 10, 1023, 1024, 1025, 50000, 3000)] + [rand_seq(rng, 5000, b"ACGTN")]
    for k, r, ms in [(56, 4, 64), (16, 1, 0), (24, 2, 12)]:
        assert_batch_equal(seqs, pg.ShmmrSpec(80, k, r, ms, True))


def test_lowercase_is_fast_path():
    rng = np.random.default_rng(19)
    s = rand_seq(rng, 60000, b"ACGTacgt")
    ctx = pg.Ctx(0)
    ctx.upload([s])
    ctx.shmmrs(pg.ShmmrSpec())
    got, off = ctx.shmmrs_download()
    assert ctx.counters()[2] == 0
    assert np.array_equal(got, orc.sequence_to_shmmrs(0, s, orc.mkspec()))
    ctx.close()


def test_long_invalid_runs_are_crossed_in_closed_form():
    """Mb-scale gaps: the replay thread jumps over the all-invalid blocks (fill segments) instead of walking them, and a run
    entered with fmmer == rmmer (leading N: all-zero registers) pushes nothing; every alignment of the run ends to the
    32-base blocks, runs closer than k / w / the cluster gap to each other, runs at both sequence ends"""
    rng = np.random.default_rng(211)
    r = lambda L, a=b"ACGT": rand_seq(rng, L, a)
    seqs = [
        r(40000) + b"N" * 300_000 + r(40000),                           # one gap
        b"N" * 100_000 + r(30000) + b"N" * 70_001,                      # telomere-style gaps at both ends
        r(5003) + b"N" * 9_999 + r(17) + b"N" * 20_000 + r(70) + b"N" * 12_345 + r(300) + b"N" * 5_000 + r(8000),   # gaps closer than k, w, cluster gap
        r(20000, b"acgt") + b"n" * 50_000 + r(20000, b"ACGTacgt"),      # soft-masked flanks, lower-case gap
        b"AT" * 40 + b"N" * 40_000 + b"AT" * 40 + r(6000),              # gap entered with palindromic registers ((AT)n, k even)
        r(10000) + b"N" * 32 + r(10000) + b"N" * 64 + r(10000) + b"N" * 31 + r(10000) + b"N" * 33 + r(9999) + b"N" * 96,
        b"".join(r(int(rng.integers(50, 3000))) + b"N" * int(rng.integers(1, 4000)) for _ in range(60)),          # scaffold-like
        r(100_000),                                                     # clean
    ]
    for i in range(32):     # every alignment of run start and length modulo 32
        seqs.append(r(3000 + i) + b"N" * (2000 + 7 * i) + r(3000))
    for w, k, r_, ms in [(80, 56, 4, 64), (48, 56, 4, 12), (24, 24, 12, 24), (128, 31, 3, 7), (5, 7, 1, 0), (1, 9, 2, 0)]:
        assert_batch_equal(seqs, pg.ShmmrSpec(w, k, r_, ms))
    assert_batch_equal(seqs[:8], pg.ShmmrSpec(80, 56, 4, 64), padding=True)
    ctx = pg.Ctx(0)
    ctx.upload(seqs)
    ctx.shmmrs(pg.ShmmrSpec())
    c = ctx.counters()
    assert c[2] == 0          # no whole-sequence replay
    assert c[5] >= 8          # the long runs were jumped over ...
    assert c[6] < 100_000     # ... and only their two ends materialised (the reference's level-0 list holds every position of a
                              # saturated run, ~0.6 M entries here; the middle cannot survive the min_span filter)
    ctx.close()


def test_assembly_like_decoration():
    """bench_synth.decorate_assembly_like (N gaps, soft masking, microsatellites, inverted repeats) on a few Mb"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench_synth as S
    rng = np.random.default_rng(5)
    seqs = []
    for i, L in enumerate((3_000_000, 1_000_003, 400_000)):
        a = S.rand_seq(rng, L).copy()
        S.decorate_assembly_like(a, 100 + i)
        seqs.append(a.tobytes())
    for spec in (pg.ShmmrSpec(80, 56, 4, 64), pg.ShmmrSpec(48, 56, 4, 12)):
        assert_batch_equal(seqs, spec)
    ctx = pg.Ctx(0)
    ctx.upload(seqs)
    ctx.shmmrs(pg.ShmmrSpec())
    assert ctx.counters()[2] == 0
    ctx.close()


def test_sketch_mode_tiled_paths():
    """sketch mode through the tiled kernels (sketch_kernels.cuh): sequences of many tiles, every tile-edge length, soft masking,
    N runs of every alignment (kb to Mb: the all-invalid bitmap walk), raw codes 0..3, junk bytes, leading / trailing runs,
    a sequence made of invalid bytes only, reverse-complement palindromes, every r"""
    rng = np.random.default_rng(23)

    def with_runs(L, runs):
        a = np.frombuffer(rand_seq(rng, L), dtype=np.uint8).copy()
        for s, n, ch in runs:
            a[s:s + n] = ch
        return a.tobytes()

    half = rand_seq(rng, 400)
    comp = bytes({65: 84, 67: 71, 71: 67, 84: 65}[c] for c in reversed(half))
    seqs = [
        rand_seq(rng, 300_000),
        rand_seq(rng, 8127), rand_seq(rng, 8128), rand_seq(rng, 8129), rand_seq(rng, 16256), rand_seq(rng, 16257), rand_seq(rng, 57), rand_seq(rng, 56), b"",
        rand_seq(rng, 100_000, b"ACGTacgt"),
        with_runs(200_000, [(0, 5000, ord("N")), (50_001, 31, ord("N")), (60_000, 32, ord("n")), (70_016, 64, ord("N")), (90_000, 1, ord("N")),
                            (120_003, 40_000, ord("N")), (199_000, 1000, ord("N"))]),
        with_runs(2_500_000, [(100_000, 2_000_000, ord("N")), (2_200_000, 3, ord("-"))]),
        with_runs(50_000, [(1000, 1, 0), (2000, 1, 1), (3000, 2, 2), (4000, 1, 3), (5000, 3, 4), (6000, 1, 255), (7000, 10, ord("*"))]),
        b"N" * 70_000,
        half + comp + rand_seq(rng, 20_000) + comp + half,
        b"AT" * 5000 + rand_seq(rng, 9000) + b"ACGT" * 3000,
    ]
    for k, r, ms in [(56, 4, 64), (56, 1, 0), (31, 12, 5), (16, 2, 0), (24, 7, 24), (5, 3, 0)]:
        assert_batch_equal(seqs, pg.ShmmrSpec(80, k, r, ms, True))


def test_sketch_mode_device_store_any_order():
    """the marked-block pass looks sequences up by store offset: a device store whose sequences are not in offset order"""
    import torch
    rng = np.random.default_rng(29)
    seqs = [np.frombuffer(rand_seq(rng, L, b"ACGTN"), dtype=np.uint8) for L in (40_000, 9_000, 70_001)]
    slack = 16384
    offs, off = [], slack
    for s in reversed(seqs):          # laid out in reverse order
        offs.append(off)
        off += (len(s) + 31) & ~31
    offs = offs[::-1]
    store = torch.zeros(off + slack, dtype=torch.uint8, device="cuda")
    for s, o in zip(seqs, offs):
        store[o:o + len(s)] = torch.from_numpy(s.copy()).cuda()
    ctx = pg.Ctx(0)
    ctx.set_device_seqs(store.data_ptr(), offs, [len(s) for s in seqs])
    spec = pg.ShmmrSpec(80, 56, 3, 10, True)
    ctx.shmmrs(spec)
    got, goff = ctx.shmmrs_download()
    exp, eoff = orc.shmmrs_batch([0, 1, 2], [s.tobytes() for s in seqs], ospec(spec), False, nthreads=3)
    assert list(goff) == list(eoff) and np.array_equal(got, exp)
    ctx.close()
