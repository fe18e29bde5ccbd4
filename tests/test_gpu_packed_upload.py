"""GPU parity of the packed host-to-device transport (pack_upload.cuh): batches large enough to take it (>= 4 MB) give the
oracle's shimmers bit for bit, whatever the bytes are — lower case, N runs, raw codes 0..3 (bases for the reference's LUT,
shmmrutils.rs:426), junk, lengths that are no multiple of 32, empty sequences — and so does the index built from them."""
import numpy as np
import pytest

import orc
import pgr_tk_b200 as pg

pytestmark = pytest.mark.gpu


def messy(rng, L, kind):
    a = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=L)].copy()
    if kind >= 1:   # soft masking + N runs
        for _ in range(max(1, L // 200_000)):
            s = int(rng.integers(0, L)); e = min(L, s + int(rng.integers(1, 50_000)))
            a[s:e] |= 0x20
        for _ in range(max(1, L // 300_000)):
            s = int(rng.integers(0, L)); e = min(L, s + int(rng.integers(1, 3_000)))
            a[s:e] = ord("N")
    if kind >= 2:   # raw codes and junk bytes sprinkled in
        idx = rng.integers(0, L, size=max(1, L // 5_000))
        a[idx] = rng.integers(0, 4, size=len(idx), dtype=np.uint8)
        idx = rng.integers(0, L, size=max(1, L // 20_000))
        a[idx] = rng.integers(0, 256, size=len(idx), dtype=np.uint8)
    return a.tobytes()


def test_packed_batch_equals_oracle():
    rng = np.random.default_rng(77)
    lens = [2_000_003, 0, 1_500_000, 31, 32, 33, 777_777, 1, 64, 900_001, 0, 1_234_567, 5, 0]
    seqs = [messy(rng, L, i % 3) if L else b"" for i, L in enumerate(lens)]
    assert sum(lens) >= 4 << 20
    for spec, padding in ((pg.ShmmrSpec(80, 56, 4, 64), False), (pg.ShmmrSpec(48, 56, 4, 12), True), (pg.ShmmrSpec(31, 17, 2, 0), False)):
        rids = list(range(100, 100 + len(seqs)))
        got, goff = pg.get_shmmrs_from_seqs(rids, seqs, spec, padding)
        exp, eoff = orc.shmmrs_batch(rids, seqs, orc.mkspec(spec.w, spec.k, spec.r, spec.min_span), padding, nthreads=8)
        assert list(goff) == list(eoff)
        assert np.array_equal(got, exp)


def test_packed_many_small_sequences():
    rng = np.random.default_rng(78)
    seqs = [messy(rng, int(rng.integers(1, 40_000)), i % 3) for i in range(400)]
    assert sum(map(len, seqs)) >= 4 << 20
    spec = pg.ShmmrSpec(80, 56, 4, 64)
    got, goff = pg.get_shmmrs_from_seqs(list(range(len(seqs))), seqs, spec)
    exp, eoff = orc.shmmrs_batch(list(range(len(seqs))), seqs, orc.mkspec(80, 56, 4, 64), False, nthreads=8)
    assert list(goff) == list(eoff) and np.array_equal(got, exp)


def test_packed_index_equals_oracle():
    rng = np.random.default_rng(79)
    base = messy(rng, 2_500_000, 1)
    seqs = [base, base[:1_000_000] + messy(rng, 1_200_000, 2) + base[1_000_000:2_000_000]]
    spec = pg.ShmmrSpec(80, 56, 4, 64)
    idx = pg.ShmmrIndex(spec, pg.FRG_ID_FASTX, 0)
    idx.add_batch([0, 1], seqs)
    gk, go, gs = idx.export()
    idx.close()
    o = orc.Index(orc.mkspec(80, 56, 4, 64), 0)
    o.add_batch([0, 1], seqs, nthreads=2)
    ek, eo, es = o.export()
    assert np.array_equal(gk, ek) and np.array_equal(go, eo)
    assert all(np.array_equal(gs[f], es[f]) for f in ("frg_id", "sid", "bgn", "end", "ori"))


def test_hybrid_feeding_of_a_page_locked_batch_equals_the_direct_copy():
    """a page-locked batch of several ring slots (hybrid feeding: some slots packed, some copied as they are) gives the same
    shimmers as the direct transport, and the oracle's on its first sequences"""
    import ctypes as C
    rng = np.random.default_rng(81)
    lens = [33_000_001, 5_000_017, 41_000_000, 64, 12_345_678, 9_000_000]
    total = sum(lens)
    hb = pg.host_alloc(total + 64 * len(lens))
    ptrs, off = [], 0
    for i, L in enumerate(lens):
        v = hb.array[off:off + L]
        v[:] = np.frombuffer(messy(rng, L, i % 3), dtype=np.uint8)
        ptrs.append(hb.ptr + off)
        off += (L + 63) & ~63
    spec = pg.ShmmrSpec(80, 56, 4, 64)
    Lb = pg.lib()

    def run():
        n = len(lens)
        cp = (C.c_void_p * n)(*ptrs)
        cl = (C.c_size_t * n)(*lens)
        rids = np.arange(n, dtype=np.uint32)
        offs = np.zeros(n + 1, dtype=np.uint64)
        out = C.c_void_p()
        rc = Lb.pgr_b200_shmmrs_batch(n, rids.ctypes.data, cp, cl, C.byref(spec), 0, C.byref(out), offs.ctypes.data)
        assert rc == 0, Lb.pgr_b200_last_error()
        k = int(offs[-1])
        mm = np.frombuffer((C.c_char * (k * 16)).from_address(out.value), dtype=pg.MM128, count=k).copy()
        Lb.pgr_b200_free(out)
        return mm, offs

    for _ in range(2):   # the second call finds the ring warm and its pack-time estimate settled
        a, ao = run()
    assert Lb.pgr_b200_last_transport() == pg.TRANSPORT_PACKED or Lb.pgr_b200_pool_threads() < 10
    prev = pg.set_transport(pg.TRANSPORT_DIRECT)
    try:
        b, bo = run()
        assert Lb.pgr_b200_last_transport() == pg.TRANSPORT_DIRECT
    finally:
        pg.set_transport(prev)
    assert np.array_equal(ao, bo) and np.array_equal(a, b)
    o1 = (lens[0] + 63) & ~63
    seq1 = bytes(hb.array[o1:o1 + lens[1]])
    exp = orc.sequence_to_shmmrs(1, seq1, orc.mkspec(80, 56, 4, 64))
    k0, k1 = int(ao[1]), int(ao[2])
    assert k1 - k0 == len(exp) and np.array_equal(a[k0:k1], exp)
    hb.free()
