"""world_size-2 gloo tests (CPU) of the multi-GPU index build: the scheme the library implements in CUDA + NCCL
(pgr_tk_b200/csrc/shard.cu) restated in numpy (tests/dist_model.py: fragment-id bases, splitters as quantiles of an
all-gathered sample, the all-to-all plan) — consecutive blocks + stable partition + stable sort must reproduce the
single-process map — and the host-side plumbing of pgr_tk_b200/distributed.py (block rule, NCCL-id broadcast).
Tuples come from the oracle here (no GPU); the CUDA kernels and NCCL are covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_seqs():
    rng = np.random.default_rng(99)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    anc = acgt[rng.integers(0, 4, size=60000)]
    seqs = []
    for h in range(7):
        s = anc.copy()
        m = rng.random(len(s)) < 0.003
        s[m] = acgt[rng.integers(0, 4, size=int(m.sum()))]
        seqs.append(s.tobytes())
    seqs.insert(3, b"")            # an empty sequence consumes two fragment ids
    return seqs


def _worker(rank, world, port, q):
    import orc
    from pgr_tk_b200 import api, distributed as PD
    import dist_model as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seqs = _make_seqs()
        n = len(seqs)
        cut = PD.shard_blocks([len(s) for s in seqs], world)          # consecutive blocks of about equal bases
        lo, hi = cut[rank], cut[rank + 1]
        # the NCCL id travels over the process group as 128 opaque bytes
        uid = PD.broadcast_id(bytes(range(128)) if rank == 0 else None)
        assert uid == bytes(range(128))
        spec = orc.mkspec(80, 56, 4, 64)
        # shard: shimmers -> fragment count -> global base
        _, offs = orc.shmmrs_batch(list(range(lo, hi)), seqs[lo:hi], spec)
        cnt = np.diff(offs.astype(np.int64))
        n_frags = int(np.where(cnt == 0, 2, cnt + 1).sum())
        base, total = D.frag_bases(n_frags)
        shard = orc.Index(spec, 0)
        shard.add_batch(list(range(lo, hi)), seqs[lo:hi])
        keys, koff, sigs = shard.export()
        # tuples in key order, insertion order inside a key; frg ids shifted to the global numbering
        t = np.zeros(len(sigs), dtype=api.TUPLE)
        rep = np.repeat(np.arange(len(keys)), np.diff(koff.astype(np.int64)))
        t["h0"], t["h1"] = keys[rep, 0], keys[rep, 1]
        for f in ("sid", "bgn", "end"):
            t[f] = sigs[f]
        t["ori"] = sigs["ori"]
        t["frg_id"] = sigs["frg_id"] + base
        # the partition works on tuples in insertion order; here they come in key order, which the stable sort undoes
        splitters = D.choose_splitters(D.sample_h0(t["h0"]))
        assert len(splitters) == world - 1 and np.all(np.diff(splitters.astype(np.float64)) >= 0)
        dest = np.searchsorted(splitters, t["h0"], side="right")
        order = np.argsort(dest, kind="stable")
        counts = np.bincount(dest, minlength=world)
        send = torch.from_numpy(t[order].view(np.uint8).copy().reshape(-1))
        recv, recv_counts = D.exchange_records(send, counts)
        got = recv.numpy().view(api.TUPLE)
        assert len(got) == sum(recv_counts)
        # owner: stable sort by key
        o2 = np.lexsort((got["h1"], got["h0"]))   # lexsort is stable
        got = got[o2]
        # reference: the single-process map restricted to this rank's key range
        full = orc.Index(spec, 0)
        full.add_batch(list(range(n)), seqs)
        fk, fo, fs = full.export()
        frep = np.repeat(np.arange(len(fk)), np.diff(fo.astype(np.int64)))
        fdest = np.searchsorted(splitters, fk[frep, 0], side="right")
        sel = fdest == rank
        ok = (np.array_equal(got["h0"], fk[frep, 0][sel]) and np.array_equal(got["h1"], fk[frep, 1][sel]) and
              all(np.array_equal(got[f], fs[f][sel]) for f in ("frg_id", "sid", "bgn", "end")) and
              np.array_equal(got["ori"], fs["ori"][sel].astype(np.uint32)))
        q.put((rank, bool(ok), int(sel.sum()), total))
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_reproduces_single_process_map():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(ok for _, ok, _, _ in res), res
    assert all(n > 0 for _, _, n, _ in res), res          # both ranks own a non-empty key range
    assert res[0][3] == res[1][3] and res[0][3] > 0       # same global fragment total
