import sys, time, ctypes as C
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
import pgr_tk_b200 as pg
n, clen = 1000, 5_000_000
bases = n*clen
hb = pg.host_alloc(bases)
t = torch.from_numpy(hb.array)
g = torch.randint(0, 4, (bases,), dtype=torch.uint8, device='cuda')
lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device='cuda')
for o in range(0, bases, 250_000_000):
    t[o:o+250_000_000].copy_(lut[g[o:o+250_000_000].long()])
del g
torch.cuda.synchronize()
ctx = pg.Ctx(0)
ptrs = [hb.ptr + i*clen for i in range(n)]
lens = [clen]*n
spec = pg.ShmmrSpec()
L = pg.lib()
for it in range(4):
    t0 = time.perf_counter(); ctx.upload_ptrs(ptrs, lens); t1 = time.perf_counter()
    ns = ctx.shmmrs(spec); t2 = time.perf_counter()
    offs = np.zeros(n + 1, dtype=np.uint64); out = C.c_void_p()
    rc = L.pgr_b200_ctx_shmmrs_download(ctx.h, C.byref(out), offs.ctypes.data); t3 = time.perf_counter()
    L.pgr_b200_free(out); t4 = time.perf_counter()
    print("upload %.1f ms  shmmrs %.1f ms  download %.1f ms  free %.1f ms  n=%d" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, ns))
    print("   ", ctx.timings())
