"""small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel family once, tiny inputs"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import pgr_tk_b200 as pg
rng = np.random.default_rng(3)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
def r(L): return acgt[rng.integers(0, 4, size=L)].tobytes()
u = r(3000)
seqs = [r(30000), r(9000) + b"N" * 50 + r(9000), r(5000) + b"AT" * 100 + r(5000), r(100), b"", r(12000)]
for spec in (pg.ShmmrSpec(80, 56, 4, 64), pg.ShmmrSpec(24, 24, 12, 24), pg.ShmmrSpec(128, 31, 2, 0), pg.ShmmrSpec(80, 56, 4, 64, True)):
    mm, off = pg.get_shmmrs_from_seqs(list(range(len(seqs))), seqs, spec)
    print("shimmers", len(mm))
idx = pg.ShmmrIndex(pg.ShmmrSpec(48, 56, 4, 12), 0)
base = r(40000)
haps = []
for h in range(5):
    s = bytearray(base)
    for p in np.nonzero(rng.random(len(s)) < 0.003)[0]: s[p] = b"ACGT"[rng.integers(0, 4)]
    haps.append(bytes(s))
idx.add_batch(list(range(5)), haps)
print("index", idx.counts())
res = idx.query_batch([haps[1][2000:22000], haps[3][10000:30000], b"ACGT"], 0.025, max_count=128, max_count_query=128, max_count_target=128, max_aln_span=8)
print("query", [len(x) for x in res])
print("adj", len(idx.adj_list(0)), len(idx.adj_list(2, [1])))
print("partition", idx.partition([1 << 40, 1 << 44]))
adj = idx.adj_list(0)
print("bundles", len(idx.get_principal_bundles_from_adj_list(adj, 2)[0]), len(idx.sort_adj_list_by_weighted_dfs(adj, (int(adj[0]["a0"]), int(adj[0]["a1"]), int(adj[0]["ori0"])))))
fr, sg = idx.compress_fragments(list(range(5)), haps)
print("fragments", len(fr), int((fr["kind"] == 0).sum()), len(sg))
# round 2: packed transport (>= 4 MB batch: pack ring, unpack_kernel, hybrid slots), sketch mode through its marked-block pass,
# level-0 patches (cluster finder, thread and warp replay, splice), sharded build on one device, per-sequence adjacency
big = [r(2_300_000) + b"N" * 70 + r(900_000), r(1_500_001).lower(), r(700_000)]
for spec in (pg.ShmmrSpec(80, 56, 4, 64), pg.ShmmrSpec(80, 56, 4, 64, True)):
    mm, off = pg.get_shmmrs_from_seqs([0, 1, 2], big, spec)
    print("packed transport shimmers", len(mm))
gap = r(60000) + b"N" * 40000 + r(30000) + b"AT" * 60 + r(20000) + b"n" * 333 + r(5000)
for spec in (pg.ShmmrSpec(80, 56, 4, 64), pg.ShmmrSpec(80, 56, 4, 64, True), pg.ShmmrSpec(80, 24, 2, 0, True)):
    print("gaps", len(pg.sequence_to_shmmrs(0, gap, spec)))
m = pg.ShardedIndex(pg.ShmmrSpec(48, 56, 4, 12), pg.FRG_ID_FASTX, devices=[0, 0, 0])
m.add_batch(list(range(5)), haps)
print("sharded", [len(x) for x in m.export()])
m.close()
# .mdb-resident look-up and query (temporary device index of the hit keys), grouped query batch
idx.write_mdb("/tmp/sanitize_small.mdb")
mm_ = pg.MdbMap("/tmp/sanitize_small.mdb")
print("mdb map", mm_.info()[1:], len(mm_.raw_query(haps[2][1000:21000])[2]), [len(x) for x in mm_.query_batch([haps[1][2000:22000], b"ACGT"], 0.025)])
mm_.close()
os.environ["PGR_B200_QUERY_GROUPS"] = "2"
