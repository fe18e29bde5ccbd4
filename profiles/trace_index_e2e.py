"""[PGR_B200_TRACE=1] python profiles/trace_index_e2e.py : host-to-CSR config-3 build on one GPU, wall clock per repetition
(and the library's stage trace when PGR_B200_TRACE is set)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import pgr_tk_b200 as pg
import bench_synth as S
n_hap, L = int(os.environ.get("HAPS", 94)), 50_000_000
views, ptrs, lens, owner = S.pangenome(L, range(n_hap), alloc=pg.host_alloc, threads=16)
comm = pg.Comm(pg.comm_unique_id(), 0, 1, 0)
spec = pg.ShmmrSpec(80, 56, 4, 64)
for it in range(4):
    sys.stderr.write("=== rep %d\n" % it)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx = pg.ShmmrIndex(spec, 0, 0)
    t1 = time.perf_counter()
    info = idx.build_sharded_ptrs(comm, list(range(n_hap)), ptrs, lens)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    sys.stderr.write("rep %d: new %.2f ms build %.2f ms %s\n" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, {k: round(v, 2) for k, v in info.items() if k.endswith("_ms")}))
    idx.close()
# the shimmers alone through the batch call, same bytes
import ctypes as C
Lb = pg.lib()
n = n_hap
cp = (C.c_void_p * n)(*[int(p) for p in ptrs]); cl = (C.c_size_t * n)(*[int(x) for x in lens])
rids = np.arange(n, dtype=np.uint32); offs = np.zeros(n + 1, dtype=np.uint64); out = C.c_void_p()
for it in range(4):
    t0 = time.perf_counter()
    rc = Lb.pgr_b200_shmmrs_batch(n, rids.ctypes.data, cp, cl, C.byref(spec), 0, C.byref(out), offs.ctypes.data)
    t1 = time.perf_counter()
    Lb.pgr_b200_free(out)
    sys.stderr.write("shmmrs_batch rep %d: %.2f ms rc %d\n" % (it, (t1 - t0) * 1e3, rc))
