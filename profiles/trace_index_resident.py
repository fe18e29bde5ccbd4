"""PGR_B200_TRACE=1 python profiles/trace_index_resident.py : where the time of a device-resident config-3 build goes"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import pgr_tk_b200 as pg
import bench_synth as S
n_hap, L = int(os.environ.get("HAPS", 94)), 50_000_000
views, ptrs, lens, owner = S.pangenome(L, range(n_hap), alloc=pg.host_alloc, threads=16)
offs, off = [], 16384
for ln in lens:
    offs.append(off); off += (ln + 31) & ~31
store = torch.zeros(off + 16384, dtype=torch.uint8, device="cuda")
for v, o, ln in zip(views, offs, lens):
    store[o:o + ln].copy_(torch.from_numpy(v))
torch.cuda.synchronize()
comm = pg.Comm(pg.comm_unique_id(), 0, 1, 0)
spec = pg.ShmmrSpec(80, 56, 4, 64)
for it in range(3):
    sys.stderr.write("=== rep %d\n" % it)
    t0 = time.perf_counter()
    idx = pg.ShmmrIndex(spec, 0, 0)
    t1 = time.perf_counter()
    info = idx.build_sharded_device(comm, store.data_ptr(), list(range(n_hap)), offs, lens)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    sys.stderr.write("rep %d: new %.2f ms build %.2f ms %s\n" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, {k: round(v, 2) for k, v in info.items() if k.endswith("_ms")}))
    idx.close()
