"""GPU: the sharded (multi-GPU) ShmmrFragMap build of the library (pgr_b200_mindex_*, pgr_tk_b200/csrc/shard.cu) gives the
same canonical map — byte-identical .mdb — as the single-GPU build and as the oracle.  Shards may share a device
(devices=[0, 0, ...]): the sampling, splitters, stable partition, count matrix, exchange layout, owner sort and slice
concatenation are then exercised on a 1-GPU box (the exchange runs through device-to-device copies; NCCL itself needs
distinct devices and is covered by test_gpu_distributed.py / bench.py --gpus N)."""
import os

import numpy as np
import pytest

import orc
import pgr_tk_b200 as pg

pytestmark = pytest.mark.gpu

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def pangenome(n_hap=9, L=120_000, seed=5, snp=0.003):
    rng = np.random.default_rng(seed)
    anc = ACGT[rng.integers(0, 4, size=L)]
    seqs = []
    for h in range(n_hap):
        s = anc.copy()
        m = rng.random(L) < snp
        s[m] = ACGT[rng.integers(0, 4, size=int(m.sum()))]
        a = int(rng.integers(0, L // 2))
        seqs.append(s[a: a + int(rng.integers(L // 3, L // 2))].tobytes())
    return seqs


def same_csr(a, b):
    ak, ao, asg = a
    bk, bo, bsg = b
    return (np.array_equal(ak, bk) and np.array_equal(ao, bo) and
            all(np.array_equal(asg[f], bsg[f]) for f in ("frg_id", "sid", "bgn", "end", "ori")))


@pytest.mark.parametrize("n_shards", [1, 2, 3, 5])
@pytest.mark.parametrize("mode", [pg.FRG_ID_FASTX, pg.FRG_ID_AGC])
def test_sharded_equals_single_and_oracle(tmp_path, n_shards, mode):
    seqs = pangenome()
    seqs.insert(4, b"")                     # an empty sequence consumes two fragment ids in FASTX numbering
    seqs.insert(7, b"ACGT" * 10)            # shorter than k: no shimmer
    sids = list(range(len(seqs)))
    spec = pg.ShmmrSpec(80, 56, 4, 64)
    single = pg.ShmmrIndex(spec, mode)
    single.add_batch(sids, seqs)
    m = pg.ShardedIndex(spec, mode, devices=[0] * n_shards)
    m.add_batch(sids, seqs)
    assert m.n_shards() == n_shards
    assert same_csr(m.export(), single.export())
    assert m.counts() == single.counts()
    o = orc.Index(orc.mkspec(80, 56, 4, 64), mode)
    o.add_batch(sids, seqs)
    assert same_csr(m.export(), o.export())
    p1, p2 = str(tmp_path / "single.mdb"), str(tmp_path / "sharded.mdb")
    single.write_mdb(p1)
    m.write_mdb(p2)
    assert open(p1, "rb").read() == open(p2, "rb").read()
    st = m.stats()
    assert sum(s["n_tuples_owned"] for s in st) == single.counts()[1]
    assert sum(s["n_tuples_local"] for s in st) == single.counts()[1]
    if n_shards > 1:
        assert sum(s["bytes_sent"] for s in st) == sum(s["bytes_recv"] for s in st) > 0
        # every shard owns one key range: slices are disjoint and ascending
        last = None
        for g in range(n_shards):
            k, _, _ = m.shard(g).export()
            if len(k):
                first = (int(k[0, 0]), int(k[0, 1]))
                assert last is None or last < first
                last = (int(k[-1, 0]), int(k[-1, 1]))
    m.close()
    single.close()


def test_sharded_several_batches_interleave(tmp_path):
    """one add_batch per input file, as pgr-make-frgdb does: the blocks of different batches interleave on the shards and the
    owner's sort restores insertion order inside every key (minor key = insertion ordinal of the sequence)"""
    seqs = pangenome(n_hap=10, L=90_000, seed=8)
    batches = [seqs[0:3], seqs[3:4], [], seqs[4:10]]
    spec = pg.ShmmrSpec(48, 56, 4, 12)
    single = pg.ShmmrIndex(spec, pg.FRG_ID_FASTX)
    m = pg.ShardedIndex(spec, pg.FRG_ID_FASTX, devices=[0, 0, 0])
    sid = 0
    for b in batches:
        ids = list(range(sid, sid + len(b)))
        single.add_batch(ids, b)
        m.add_batch(ids, b)
        sid += len(b)
    assert same_csr(m.export(), single.export())
    assert m.counts() == single.counts()
    m.close()
    single.close()


def test_sharded_more_shards_than_sequences_and_empty():
    spec = pg.ShmmrSpec(80, 56, 4, 64)
    seqs = pangenome(n_hap=2, L=60_000, seed=3)
    single = pg.ShmmrIndex(spec, 0)
    single.add_batch([0, 1], seqs)
    m = pg.ShardedIndex(spec, 0, devices=[0] * 4)
    m.add_batch([0, 1], seqs)
    assert same_csr(m.export(), single.export())
    m.close()
    e = pg.ShardedIndex(spec, 0, devices=[0, 0])
    e.add_batch([], [])
    assert e.counts() == (0, 0, 0)
    e.close()
    single.close()


def test_sharded_needs_visible_devices():
    spec = pg.ShmmrSpec(80, 56, 4, 64)
    with pytest.raises(pg.PgrError):
        pg.ShardedIndex(spec, 0, n_gpus=pg.device_count() + 1)


def test_sharded_nccl_two_gpus(tmp_path):
    """the same through NCCL (one process, one host thread per GPU) when the box has two GPUs"""
    if pg.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    seqs = pangenome(n_hap=12, L=200_000, seed=21)
    sids = list(range(len(seqs)))
    spec = pg.ShmmrSpec(80, 56, 4, 64)
    single = pg.ShmmrIndex(spec, 0)
    single.add_batch(sids, seqs)
    m = pg.ShardedIndex(spec, 0, n_gpus=2)
    m.add_batch(sids, seqs)
    assert same_csr(m.export(), single.export())
    m.close()
    single.close()
