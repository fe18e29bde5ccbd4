"""python profiles/hostpack_rate.py : host packing rate of pgr_b200_pack_bases (the host half of the packed transport) against
the thread count, next to a plain read of the same bytes (numpy sum) — shows whether the e2e leg is bound by the cores or by
the host's memory system.  No GPU needed."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pgr_tk_b200 as pg

L = pg.lib()
per = 128 << 20
ncpu = len(os.sched_getaffinity(0))
print("cpus available: %d, isa: %s" % (ncpu, pg.pack_isa()))
tmax = max(32, ncpu)
rng = np.random.default_rng(0)
block = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=per)]
bufs = [block.copy() for _ in range(tmax)]
outs = [np.zeros(per // 32 * 3, dtype=np.uint32) for _ in range(tmax)]
nb = per // 32
def pack(t):
    o = outs[t]
    L.pgr_b200_pack_bases(bufs[t].ctypes.data, per, o.ctypes.data, o[nb:].ctypes.data, o[2 * nb:].ctypes.data)
def read(t):
    bufs[t].view(np.uint64).sum()
for name, fn in (("pack", pack), ("read (numpy sum)", read)):
    for T in (1, 2, 4, 8, 16, 32):
        if T > tmax:
            break
        best = 0
        for rep in range(3):
            th = [threading.Thread(target=fn, args=(t,)) for t in range(T)]
            t0 = time.perf_counter()
            [x.start() for x in th]; [x.join() for x in th]
            best = max(best, T * per / (time.perf_counter() - t0) / 1e9)
        print("%-18s T=%2d  %6.1f GB/s  (%.1f per thread)" % (name, T, best, best / T))
