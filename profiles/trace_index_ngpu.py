"""torchrun --nproc-per-node N profiles/trace_index_ngpu.py : per-rank host wall clock and CUDA-event stages of the
HBM-resident config-3 build (pgr_b200_index_build_sharded_device), to see where a multi-GPU step spends its time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.distributed as dist
import pgr_tk_b200 as pg
from pgr_tk_b200 import distributed as PD
import bench_synth as S

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n_hap, L = int(os.environ.get("HAPS", 94)), 50_000_000
lo, hi = (n_hap * rank) // world, (n_hap * (rank + 1)) // world
views, ptrs, lens, owner = S.pangenome(L, range(lo, hi), alloc=pg.host_alloc, threads=8)
offs, off = [], 16384
for ln in lens:
    offs.append(off); off += (ln + 31) & ~31
store = torch.zeros(off + 16384, dtype=torch.uint8, device="cuda")
for v, o, ln in zip(views, offs, lens):
    store[o:o + ln].copy_(torch.from_numpy(v))
torch.cuda.synchronize()
comm = PD.init_comm(lr) if world > 1 else pg.Comm(pg.comm_unique_id(), 0, 1, lr)
spec = pg.ShmmrSpec(80, 56, 4, 64)
sids = list(range(lo, hi))
for it in range(5):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx = pg.ShmmrIndex(spec, 0, lr)
    t1 = time.perf_counter()
    info = idx.build_sharded_device(comm, store.data_ptr(), sids, offs, lens)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    sys.stderr.write("rank %d rep %d: new %.2f ms build %.2f ms sync %.2f ms %s\n" % (
        rank, it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, {k: round(v, 2) for k, v in info.items() if k.endswith("_ms")}))
    t4 = time.perf_counter()
    idx.close()
    sys.stderr.write("rank %d rep %d: close %.2f ms\n" % (rank, it, (time.perf_counter() - t4) * 1e3))
comm.close()
if world > 1:
    dist.destroy_process_group()
