#!/usr/bin/env python
"""bench_configs.py — the other BASELINE.json configs (1, 3, 4, 5) measured on the B200 path, with the CPU oracle timed
beside them on a bounded sample.  bench.py (config 2) is the driver's contract; this script produces the supplementary
lines kept under profiles/.  Run:  python bench_configs.py [--scale 1.0] [--configs 1,3,4,5]
Multi-GPU index build (config 3):  torchrun --nproc-per-node N --master-addr 127.0.0.1 bench_configs.py --configs 3
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGT", b"TGCA"):
    COMP[a] = b


def rand_seq(rng, L):
    return ACGT[rng.integers(0, 4, size=L, dtype=np.uint8)]


def derive_haplotype(rng, anc, snp, indel, n_sv):
    """SNPs at rate snp, indels (geometric length, mean 3) at rate indel, n_sv inversions and n_sv tandem duplications"""
    s = anc.copy()
    L = len(s)
    m = np.nonzero(rng.random(L) < snp)[0]
    s[m] = ACGT[rng.integers(0, 4, size=len(m), dtype=np.uint8)]
    pos = np.sort(np.nonzero(rng.random(L) < indel)[0])
    if len(pos):
        lens = rng.geometric(1.0 / 3.0, size=len(pos))
        is_del = rng.random(len(pos)) < 0.5
        pieces, last = [], 0
        for p, ln, d in zip(pos, lens, is_del):
            if p < last:
                continue
            pieces.append(s[last:p])
            if d:
                last = min(L, p + ln)
            else:
                pieces.append(rand_seq(rng, ln))
                last = p
        pieces.append(s[last:])
        s = np.concatenate(pieces)
    for _ in range(n_sv):
        L = len(s)
        ln = int(rng.integers(1000, 50000))
        a = int(rng.integers(0, max(1, L - ln)))
        s[a:a + ln] = COMP[s[a:a + ln]][::-1]
    for _ in range(n_sv):
        L = len(s)
        ln = int(rng.integers(1000, 50000))
        a = int(rng.integers(0, max(1, L - ln)))
        s = np.concatenate([s[:a + ln], s[a:a + ln], s[a + ln:]])
    return s


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def emit(d):
    print(json.dumps(d), flush=True)


def config1(pg, orc, scale):
    """one synthetic 1 Mb contig, 80/56/4/64: canonical .mdb byte-identical to the oracle's"""
    import tempfile
    rng = np.random.default_rng(1)
    seq = rand_seq(rng, int(1_000_000 * scale)).tobytes()
    spec = pg.ShmmrSpec(80, 56, 4, 64)
    with tempfile.TemporaryDirectory() as td:
        t0 = time.perf_counter()
        o = orc.Index(orc.mkspec(80, 56, 4, 64), 0)
        o.add_batch([0], [seq])
        o.write_mdb(os.path.join(td, "o.mdb"))
        t_cpu = time.perf_counter() - t0
        g = pg.ShmmrIndex(spec, 0)
        g.add_batch([0], [seq])   # warm
        g.close()
        t0 = time.perf_counter()
        g = pg.ShmmrIndex(spec, 0)
        t1 = time.perf_counter()
        g.add_batch([0], [seq])
        t2 = time.perf_counter()
        g.write_mdb(os.path.join(td, "g.mdb"))
        t_gpu = time.perf_counter() - t0
        parts = {"index_new_ms": (t1 - t0) * 1e3, "add_batch_ms": (t2 - t1) * 1e3, "finalize_write_mdb_ms": (t0 + t_gpu - t2) * 1e3}
        same = open(os.path.join(td, "o.mdb"), "rb").read() == open(os.path.join(td, "g.mdb"), "rb").read()
        nk, ns, nf = g.counts()
    emit({"config": 1, "workload": "pgr-make-frgdb index part on one %d-base contig, 80/56/4/64" % len(seq), "mdb_byte_identical": same,
          "n_keys": nk, "n_sigs": ns, "gpu_ms_host_to_mdb": t_gpu * 1e3, "gpu_parts": parts, "cpu_oracle_ms": t_cpu * 1e3})
    assert same


def frags_equal(fa, sa, fb, sb):
    """two (FRAGMENT[], ALNSEG[]) results describe the same fragments (segment arrays may be laid out in different orders)"""
    if len(fa) != len(fb):
        return False
    for f in ("kind", "reversed", "sid", "bgn", "end", "len", "ref_frag", "n_segs"):
        if not np.array_equal(fa[f], fb[f]):
            return False

    def in_id_order(fr, sg):
        cnt = fr["n_segs"].astype(np.int64)
        tot = int(cnt.sum())
        if tot == 0:
            return sg[:0]
        first = np.cumsum(cnt) - cnt                                   # position of each fragment's first segment in id order
        idx = np.repeat(fr["seg_off"].astype(np.int64) - first, cnt) + np.arange(tot)
        return sg[idx]
    ga, gb = in_id_order(fa, sa), in_id_order(fb, sb)
    return len(ga) == len(gb) and all(np.array_equal(ga[f], gb[f]) for f in ("type", "a", "b"))


def config3(pg, orc, scale, dist_info):
    """full ShmmrFragMap build on 94 haplotypes x 50 Mb derived from one ancestor"""
    import torch
    rank, world, local = dist_info
    n_hap, L = 94, int(50_000_000 * scale)
    import bench_synth as S
    lo, hi = (n_hap * rank) // world, (n_hap * (rank + 1)) // world
    t0 = time.perf_counter()
    haps, _, _, _ = S.pangenome(L, range(lo, hi), threads=min(16, host_cores()))
    gen_s = time.perf_counter() - t0
    bases_local = sum(len(h) for h in haps)
    spec = pg.ShmmrSpec(80, 56, 4, 64)
    # pinned staging so that the H2D copies run at PCIe rate (what a loader with pinned buffers would hand over)
    hb = pg.host_alloc(bases_local + 64 * len(haps))
    ptrs, lens, off = [], [], 0
    for h in haps:
        hb.array[off:off + len(h)] = h
        ptrs.append(hb.ptr + off)
        lens.append(len(h))
        off += (len(h) + 63) & ~63
    views = [hb.array[p - hb.ptr: p - hb.ptr + ln] for p, ln in zip(ptrs, lens)]
    if world == 1:
        times = []
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            g = pg.ShmmrIndex(spec, 0)
            g.add_batch(list(range(n_hap)), views)
            g.finalize()
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
            if it < 2:
                g.close()
        nk, ns, nf = g.counts()
        # parity: the first two haplotypes alone, GPU vs oracle (same map); and size-independent properties on the full index
        o = orc.Index(orc.mkspec(80, 56, 4, 64), 0)
        t0 = time.perf_counter()
        o.add_batch([0, 1], [haps[0].tobytes(), haps[1].tobytes()], nthreads=2)
        cpu_s = time.perf_counter() - t0
        g2 = pg.ShmmrIndex(spec, 0)
        g2.add_batch([0, 1], views[:2])
        gk, go, gs = g2.export()
        ok_, oo, os_ = o.export()
        parity = bool(np.array_equal(gk, ok_) and np.array_equal(go, oo) and all(np.array_equal(gs[f], os_[f]) for f in ("frg_id", "sid", "bgn", "end", "ori")))
        keys, offs, sigs = g.export()
        props = {
            "keys_strictly_ascending": bool(np.all((keys[1:, 0] > keys[:-1, 0]) | ((keys[1:, 0] == keys[:-1, 0]) & (keys[1:, 1] > keys[:-1, 1])))),
            "h0_le_h1": bool(np.all(keys[:, 0] <= keys[:, 1])),
            "sigs_total": int(offs[-1]) == ns,
            "per_key_sid_nondecreasing": bool(np.all((np.diff(sigs["sid"].astype(np.int64)) >= 0) | np.isin(np.arange(1, ns), offs[1:-1].astype(np.int64)))),
            "bgn_lt_end": bool(np.all(sigs["bgn"] < sigs["end"])),
            "frg_ids_unique": int(len(np.unique(sigs["frg_id"]))) == ns,
        }
        # fragment compression (the .frg content of pgr-make-frgdb, SURVEY f-1): every internal fragment aligned on the GPU
        frag_info = None
        if os.environ.get("PGR_B200_BENCH_FRAGS", "1") != "0":
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fr, sg = g.compress_fragments(list(range(n_hap)), views)
            fr_s = time.perf_counter() - t0
            kinds = np.bincount(fr["kind"], minlength=4)
            raw = int(fr["len"][fr["kind"] != 0].sum())
            # CPU beside it + parity at size: the first three haplotypes through the C++ oracle (all cores for a sequence's pairs,
            # inserts sequential, as the reference's par_iter does) and through the GPU on their own
            ns3 = min(3, n_hap)
            t0 = time.perf_counter()
            ofr, osg = orc.compress_fragments(list(range(ns3)), [v for v in views[:ns3]], orc.mkspec(80, 56, 4, 64), nthreads=host_cores())
            cpu_fr_s = time.perf_counter() - t0
            g3 = pg.ShmmrIndex(spec, 0)
            g3.add_batch(list(range(ns3)), views[:ns3])
            gfr, gsg = g3.compress_fragments(list(range(ns3)), views[:ns3])
            g3.close()
            frag_parity = frags_equal(gfr, gsg, ofr, osg)
            b3 = sum(len(v) for v in views[:ns3])
            frag_info = {"ms": fr_s * 1e3, "gbases_per_s": bases_local / fr_s / 1e9, "parity_%d_haplotypes_vs_oracle" % ns3: bool(frag_parity),
                         "cpu_baseline": {"value": b3 / cpu_fr_s / 1e9, "unit": "Gbases/s", "cores": host_cores(), "kind": "port",
                                          "sample": "%d of %d haplotypes: shimmers + seq_to_compressed (oracle/frag_oracle.cpp)" % (ns3, n_hap)},
                         "n_fragments": int(len(fr)), "aln_segments_fragments": int(kinds[0]), "internal_raw": int(kinds[2]),
                         "alignment_segments": int(len(sg)), "bases_kept_raw": raw, "compression": bases_local / max(raw, 1),
                         "note": "host sequences in, per-fragment records + segments out (H2D of the bases and D2H inside)"}
        emit({"config": 3, "workload": "ShmmrFragMap build, %d haplotypes x %d bases (%.2f Gbases), 80/56/4/64" % (n_hap, L, bases_local / 1e9),
              "fragment_compression": frag_info,
              "n_gpus": 1, "value": bases_local / min(times) / 1e9, "unit": "Gbases/s", "ms": min(times) * 1e3, "times_ms": [t * 1e3 for t in times],
              "n_keys": nk, "n_sigs": ns, "n_frags": nf, "timed": "host (pinned) sequences -> sorted CSR in HBM (H2D inside)",
              "parity_two_haplotypes_vs_oracle": parity, "properties": props,
              "cpu_baseline": {"value": (len(haps[0]) + len(haps[1])) / cpu_s / 1e9, "unit": "Gbases/s", "cores": 2, "kind": "port",
                               "sample": "2 of 94 haplotypes, one sequence per thread + single-threaded inserts (seq_db.rs:461,326-340)"},
              "synth_gen_s": gen_s})
        assert parity and all(props.values()), props
        assert frag_info is None or frag_info["parity_%d_haplotypes_vs_oracle" % min(3, n_hap)], "GPU fragments differ from the oracle"
        return g, haps
    import torch.distributed as dist
    from pgr_tk_b200 import distributed as D
    best = None
    for it in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g, info = D.build_index_distributed(spec, list(range(lo, hi)), views, pg.FRG_ID_FASTX, device=local)
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        if best is None or dt < best[0]:
            best = (dt, info)
        if it < 2:
            g.close()
    tb = torch.tensor([bases_local, info["n_keys"], info["n_sigs"]], dtype=torch.int64, device="cuda")
    dist.all_reduce(tb)
    if rank == 0:
        emit({"config": 3, "workload": "ShmmrFragMap build, %d haplotypes x %d bases, 80/56/4/64, NCCL all-to-all merge" % (n_hap, L),
              "n_gpus": world, "value": int(tb[0]) / best[0] / 1e9, "unit": "Gbases/s", "ms": best[0] * 1e3, "n_keys": int(tb[1]),
              "n_sigs": int(tb[2]), "rank0_stages_ms": best[1], "timed": "host (pinned) shards -> per-rank sorted CSR slices (H2D + all-to-all inside), max over ranks"})
    return g, haps


def config4(pg, orc, scale, g, haps):
    """10k x 20 kb queries against the config-3 index"""
    import torch
    rng = np.random.default_rng(11)
    n_q, qlen = int(10000 * min(1.0, scale * 4)), 20000
    queries = []
    for i in range(n_q):
        h = int(rng.integers(0, len(haps)))
        a = int(rng.integers(0, len(haps[h]) - qlen))
        q = haps[h][a:a + qlen].copy()
        m = np.nonzero(rng.random(qlen) < 1e-3)[0]
        q[m] = ACGT[rng.integers(0, 4, size=len(m), dtype=np.uint8)]
        if rng.random() < 0.5:
            q = COMP[q][::-1].copy()
        queries.append(q)
    kw = dict(max_count=128, max_count_query=128, max_count_target=128, max_aln_span=8)
    # time the C ABI call itself (host query buffers in, host result arrays out); the numpy wrapper's copies are not part of it
    import ctypes as C
    from pgr_tk_b200 import api
    L = pg.lib()
    ptrs = (C.c_void_p * n_q)(*[q.ctypes.data for q in queries])
    lens = (C.c_size_t * n_q)(*[q.size for q in queries])
    prm = api.QueryParams(0.025, 128, 128, 128, 8, -1, 0)
    times = []
    for it in range(4):
        res_p = C.POINTER(api.QueryResult)()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = L.pgr_b200_query_batch(g.h, n_q, ptrs, lens, C.byref(prm), C.byref(res_p))
        times.append(time.perf_counter() - t0)
        assert rc == 0, L.pgr_b200_last_error()
        L.pgr_b200_query_result_free(res_p)
    times = times[1:]
    res = g.query_batch(queries, 0.025, **kw)
    qto, tsid, tco, csc, cho, hits = res
    # parity AT SIZE: the oracle answers the first n_chk queries against the SAME 94-haplotype map (the GPU index written as
    # .mdb and read back by the oracle: the map itself is checked against the oracle's own build in bench.py's index_build)
    import tempfile
    n_chk = int(os.environ.get("PGR_B200_BENCH_QUERY_CHECK", "1000"))
    n_chk = min(n_chk, n_q)
    with tempfile.TemporaryDirectory() as td:
        g.write_mdb(os.path.join(td, "g.mdb"))
        o = orc.Index.read_mdb(os.path.join(td, "g.mdb"))
    t0 = time.perf_counter()
    parity, n_cmp_hits = True, 0
    for qi in range(n_chk):
        osid, otco, osc, ocho, ohits = o.query_fragment_to_hps(queries[qi].tobytes(), 0.025, **kw)
        a, b = int(qto[qi]), int(qto[qi + 1])
        c0, c1 = int(tco[a]), int(tco[b])
        h0, h1 = int(cho[c0]), int(cho[c1])
        ok = np.array_equal(tsid[a:b], osid) and np.array_equal(tco[a:b + 1] - tco[a], otco) and np.array_equal(csc[c0:c1].view(np.uint32), osc.view(np.uint32))
        ok = ok and np.array_equal(cho[c0:c1 + 1] - cho[c0], ocho) and all(np.array_equal(hits[h0:h1][f], ohits[f]) for f in ("qb", "qe", "qo", "tb", "te", "to"))
        parity = parity and bool(ok)
        n_cmp_hits += h1 - h0
    cpu_s = time.perf_counter() - t0
    emit({"config": 4, "workload": "%d x %d bp queries vs the %d-haplotype index, pgr-query defaults (0.025,128,128,128,8)" % (n_q, qlen, len(haps)),
          "n_gpus": 1, "value": n_q / min(times), "unit": "queries/s", "gbases_per_s": n_q * qlen / min(times) / 1e9, "ms": min(times) * 1e3, "times_ms": [t * 1e3 for t in times],
          "targets": int(len(tsid)), "chains": int(len(csc)), "hit_pairs": int(len(hits)),
          "parity_%d_queries_vs_oracle_same_index" % n_chk: bool(parity), "hit_pairs_compared": int(n_cmp_hits),
          "algorithmic_bytes": int(n_q * qlen + len(hits) * 41), "achieved_gbs": (n_q * qlen + len(hits) * 41) / min(times) / 1e9,
          "cpu_baseline": {"value": n_chk / cpu_s, "unit": "queries/s", "cores": 1, "kind": "port", "sample": "%d queries against the same 94-haplotype map, one thread (the reference runs one rayon task per query)" % n_chk}})
    assert parity


def config5(pg, orc, scale):
    """MAP-graph adjacency on 94 haplotypes of a 300 kb repetitive locus, 48/56/4/12"""
    import torch
    rng = np.random.default_rng(13)
    units = [rand_seq(rng, 15000) for _ in range(12)]
    flank_l, flank_r = rand_seq(rng, 60000), rand_seq(rng, 60000)
    haps = []
    for h in range(94):
        r = np.random.default_rng(1300 + h)
        parts = [flank_l]
        for u in units:
            for _ in range(int(r.integers(1, 5))):
                v = u.copy()
                if r.random() < 0.05:
                    v = COMP[v][::-1].copy()
                parts.append(v)
        parts.append(flank_r)
        s = np.concatenate(parts)
        m = np.nonzero(r.random(len(s)) < 2e-3)[0]
        s[m] = ACGT[r.integers(0, 4, size=len(m), dtype=np.uint8)]
        haps.append(s)
    spec_t = (48, 56, 4, 12)
    times = []
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = pg.ShmmrIndex(pg.ShmmrSpec(*spec_t), 0)
        g.add_batch(list(range(94)), haps)
        adj = g.adj_list(0)
        times.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    o = orc.Index(orc.mkspec(*spec_t), 0)
    o.add_batch(list(range(94)), [h.tobytes() for h in haps], nthreads=host_cores())
    oadj = o.adj_list(0)
    cpu_s = time.perf_counter() - t0
    names = ["sid", "ori0", "ori1", "a0", "a1", "b0", "b1"]
    parity = len(adj) == len(oadj) and all(np.array_equal(adj[f], oadj[f]) for f in names)
    adj2, oadj2 = g.adj_list(3, [0, 5]), o.adj_list(3, [0, 5])
    parity = parity and len(adj2) == len(oadj2) and all(np.array_equal(adj2[f], oadj2[f]) for f in names)
    # principal bundles (pgr-pbundle-decomp defaults: min_cov 0, min_branch_size 8): host-side walk over the GPU adjacency list
    t0 = time.perf_counter()
    bundles, flt = g.get_principal_bundles_from_adj_list(adj, 8)
    pb_s = time.perf_counter() - t0
    emit({"config": 5, "workload": "shimmers + ShmmrFragMap + frag_map_to_adj_list, 94 haplotypes of a repetitive locus (%.1f Mbases), 48/56/4/12" % (sum(map(len, haps)) / 1e6),
          "n_gpus": 1, "ms": min(times) * 1e3, "value": sum(map(len, haps)) / min(times) / 1e9, "unit": "Gbases/s", "adj_pairs": int(len(adj)),
          "n_sigs": g.counts()[1], "adjlist_bit_exact_vs_oracle": bool(parity),
          "principal_bundles": {"ms": pb_s * 1e3, "n_bundles": len(bundles), "vertices": int(sum(len(b) for b in bundles)), "filtered_adj_pairs": int(len(flt)),
                                "note": "sequential graph walk on the host as in the reference (seq_db.rs:1063-1186); parity unpinned (petgraph orders), tests compare with oracle/bundles_oracle.py"},
          "cpu_baseline": {"ms": cpu_s * 1e3, "cores": host_cores(), "kind": "port", "sample": "the whole config (oracle, all cores for shimmers)"}})
    assert parity


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--configs", default="1,3,4,5")
    args = ap.parse_args()
    cfgs = [int(x) for x in args.configs.split(",")]
    import torch
    import orc
    import pgr_tk_b200 as pg
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pg.set_default_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if 1 in cfgs and rank == 0 and world == 1:
        config1(pg, orc, args.scale)
    g = haps = None
    if 3 in cfgs:
        g, haps = config3(pg, orc, args.scale, (rank, world, local))
    if 4 in cfgs and world == 1:
        config4(pg, orc, args.scale, g, haps)
    if 5 in cfgs and world == 1:
        config5(pg, orc, args.scale)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
