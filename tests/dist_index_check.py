"""torchrun script: multi-GPU ShmmrFragMap build vs the oracle.  Launched by tests/test_gpu_distributed.py (and by hand:
torchrun --nproc-per-node N --master-addr 127.0.0.1 tests/dist_index_check.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import orc  # noqa: E402
import pgr_tk_b200 as pg  # noqa: E402
from pgr_tk_b200 import distributed as D  # noqa: E402


def make_seqs(n_hap=12, L=200000):
    rng = np.random.default_rng(4242)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    anc = acgt[rng.integers(0, 4, size=L)]
    seqs = []
    for h in range(n_hap):
        s = anc.copy()
        m = rng.random(L) < 0.002
        s[m] = acgt[rng.integers(0, 4, size=int(m.sum()))]
        seqs.append(s.tobytes())
    seqs.insert(5, b"")
    return seqs


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pg.set_default_device(local)
    seqs = make_seqs()
    n = len(seqs)
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    spec = pg.ShmmrSpec(80, 56, 4, 64)
    idx, info = D.build_index_distributed(spec, list(range(lo, hi)), seqs[lo:hi], pg.FRG_ID_FASTX, device=local)
    gk, go, gs = idx.export()
    # oracle: full map; this rank must hold exactly a contiguous key range of it, with identical per-key vectors
    full = orc.Index(orc.mkspec(80, 56, 4, 64), 0)
    full.add_batch(list(range(n)), seqs, nthreads=4)
    fk, fo, fs = full.export()
    # all ranks' key counts -> where my slice starts in the global key order
    t = torch.tensor([len(gk)], dtype=torch.int64, device="cuda")
    allc = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allc, t)
    allc = [int(x.item()) for x in allc]
    k0 = sum(allc[:rank])
    ok = sum(allc) == len(fk)
    ok = ok and np.array_equal(gk, fk[k0:k0 + len(gk)])
    s0, s1 = int(fo[k0]), int(fo[k0 + len(gk)])
    ok = ok and np.array_equal(go, fo[k0:k0 + len(gk) + 1] - fo[k0])
    for f in ("frg_id", "sid", "bgn", "end", "ori"):
        ok = ok and np.array_equal(gs[f], fs[f][s0:s1])
    flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_INDEX_CHECK", "OK" if int(flag.item()) == 1 else "MISMATCH", "world", world, "keys", allc, info)
    idx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
