"""Multi-GPU ShmmrFragMap build, one process per GPU (torchrun): the plumbing around the library's collective
`pgr_b200_index_build_sharded` (pgr_tk_b200/csrc/shard.cu).

    rank 0: pgr_b200_comm_unique_id  --broadcast (torch.distributed, any backend)-->  every rank: pgr_b200_comm_init_rank
    every rank: pgr_b200_index_build_sharded(its consecutive block of the sequence list)
        = shimmers + tuples -> all-gather of fragment totals -> device-side splitters -> stable partition
          -> ONE all-to-all (grouped ncclSend/ncclRecv over NVLink) -> stable sort + CSR of the rank's key range

torch.distributed only carries the 128-byte NCCL id; the exchange itself runs inside libpgr_b200 on the library's
own communicator.  `shard_blocks` is the block rule the single-process form (pgr_b200_mindex_add_batch) uses too.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import api


def shard_blocks(lens, world):
    """cut a sequence list into `world` consecutive blocks of about equal bases: cut g falls after the first sequence
    that brings the running total to g/world of all bases.  Returns world+1 boundaries."""
    n = len(lens)
    total = int(sum(int(x) for x in lens))
    cut = [0] + [n] * world
    acc, g = 0, 1
    for i, ln in enumerate(lens):
        if g >= world:
            break
        acc += int(ln)
        while g < world and acc * world >= total * g:
            cut[g] = i + 1
            g += 1
    return cut


def broadcast_id(id_bytes, group=None, device="cpu"):
    """rank 0 passes the id bytes, the others None; everybody gets rank 0's bytes"""
    t = torch.zeros(api.COMM_ID_BYTES, dtype=torch.uint8)
    if dist.get_rank(group) == 0:
        t = torch.frombuffer(bytearray(id_bytes), dtype=torch.uint8).clone()
    t = t.to(device)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return bytes(t.cpu().numpy().tobytes())


def init_comm(device=None, group=None):
    """the library's own NCCL communicator over the ranks of `group` (collective)"""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev_index = torch.cuda.current_device() if device is None else device
    tdev = torch.device("cuda", dev_index) if dist.get_backend(group) == "nccl" else "cpu"
    uid = api.comm_unique_id() if rank == 0 else None
    uid = broadcast_id(uid, group, tdev)
    return api.Comm(uid, rank, world, dev_index)


def build_index_distributed(spec, local_sids, local_seqs, frg_id_mode=api.FRG_ID_FASTX, device=None, group=None, comm=None):
    """Every rank passes its consecutive block of the global sequence list (rank order = list order); returns this
    rank's ShmmrIndex slice (the rank's key range, finalized) and the rank's pgr_shard_stats.  Collective."""
    dev_index = torch.cuda.current_device() if device is None else device
    own = comm is None
    if own:
        comm = init_comm(dev_index, group)
    idx = api.ShmmrIndex(spec, frg_id_mode, dev_index)
    info = idx.build_sharded(comm, local_sids, local_seqs)
    nk, ns, _ = idx.counts()
    info.update({"n_keys": nk, "n_sigs": ns})
    if own:
        comm.close()
    return idx, info
