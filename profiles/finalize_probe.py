import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import pgr_tk_b200 as pg
rng = np.random.default_rng(1)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
anc = acgt[rng.integers(0, 4, size=50_000_000, dtype=np.uint8)]
haps = []
for h in range(3):
    s = anc.copy(); m = np.nonzero(rng.random(len(s)) < 1e-3)[0]; s[m] = acgt[rng.integers(0, 4, size=len(m), dtype=np.uint8)]; haps.append(s)
for it in range(2):
    g = pg.ShmmrIndex(pg.ShmmrSpec(), 0)
    g.add_batch([0, 1, 2], haps)
    t0 = time.perf_counter(); g.finalize(); print("finalize ms", (time.perf_counter() - t0) * 1e3, g.counts(), flush=True)
    g.close()
