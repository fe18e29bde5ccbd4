"""[PGR_B200_TRACE=1] python profiles/trace_assembly.py : the assembly-like variant of config 2 (bench_synth.decorate_assembly_like)
on CONTIGS x 5 Mb resident in HBM; per-step CUDA-event stages, and the library's host-side stage trace with PGR_B200_TRACE."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import concurrent.futures as cf
import numpy as np, torch
import pgr_tk_b200 as pg
import bench_synth as S
n, clen = int(os.environ.get("CONTIGS", 1000)), 5_000_000
hb = pg.host_alloc(n * clen)
def gen(i):
    rng = np.random.default_rng(1000 + i)
    hb.array[i * clen:(i + 1) * clen] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=clen, dtype=np.uint8)]
    S.decorate_assembly_like(hb.array[i * clen:(i + 1) * clen], 5000 + i)
with cf.ThreadPoolExecutor(16) as ex:
    list(ex.map(gen, range(n)))
store = torch.zeros(n * clen + 2 * 16384, dtype=torch.uint8, device="cuda")
store[16384:16384 + n * clen].copy_(torch.from_numpy(hb.array))
ctx = pg.Ctx(0)
ctx.set_device_seqs(store.data_ptr(), [16384 + i * clen for i in range(n)], [clen] * n)
spec = pg.ShmmrSpec(80, 56, 4, 64)
for it in range(4):
    sys.stderr.write("=== rep %d\n" % it)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ns = ctx.shmmrs(spec)
    torch.cuda.synchronize()
    sys.stderr.write("rep %d: %.2f ms, %d shimmers, stages %s counters %s\n" % (it, (time.perf_counter() - t0) * 1e3, ns,
                     {k: round(v, 2) for k, v in ctx.timings()}, list(ctx.counters())))
