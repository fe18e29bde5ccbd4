"""Multi-GPU ShmmrFragMap build: one process per GPU, sequences sharded in contiguous blocks, ONE exchange step.

    stage (shimmers, per rank) -> all-gather fragment totals (global FASTX frg_id bases, seq_db.rs:203-231)
    -> commit (tuples) -> sampled key splitters -> stable partition -> all-to-all of 40-byte tuples over NCCL/NVLink
    -> per-owner stable sort + CSR (rank r owns the r-th key range, so slices concatenate in canonical key order)

torch.distributed is the plumbing (process group, all_gather, all_to_all_single); partition, sort and CSR are the
library's CUDA kernels.  The host-side pieces (`frag_bases`, `choose_splitters`, `exchange_records`) work on CPU
tensors with the gloo backend too, which is how tests/test_distributed_gloo.py covers them without a GPU.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import api

TUPLE_BYTES = api.TUPLE.itemsize  # 40
SAMPLES_PER_RANK = 4096


def frag_bases(n_frags_local, group=None, device="cpu"):
    """exclusive prefix over ranks of the fragment ids each shard consumes -> (base of this rank, total)"""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    t = torch.tensor([int(n_frags_local)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    vals = [int(x.item()) for x in out]
    return sum(vals[:rank]), sum(vals)


def choose_splitters(local_h0_sample, group=None, device="cpu"):
    """world-1 ascending splitters of the h0 key space from an all-gathered sample (minimizer hashes are skewed low,
    so quantiles of a sample, not top bits)"""
    world = dist.get_world_size(group)
    s = np.zeros(SAMPLES_PER_RANK + 1, dtype=np.int64)
    k = min(len(local_h0_sample), SAMPLES_PER_RANK)
    s[0] = k
    s[1:1 + k] = np.asarray(local_h0_sample[:k], dtype=np.uint64).view(np.int64)
    t = torch.from_numpy(s).to(device)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    allv = []
    for o in out:
        o = o.cpu().numpy()
        allv.append(o[1:1 + int(o[0])].view(np.uint64))
    allv = np.sort(np.concatenate(allv)) if allv else np.zeros(0, dtype=np.uint64)
    if len(allv) == 0:
        return np.zeros(world - 1, dtype=np.uint64)
    q = [(len(allv) * (i + 1)) // world for i in range(world - 1)]
    return allv[np.minimum(q, len(allv) - 1)].astype(np.uint64)


def exchange_records(send, send_counts, group=None):
    """all-to-all of fixed-size records. send: uint8 tensor [n_send * rec] already ordered by destination rank;
    send_counts: records per destination.  Returns (recv uint8 tensor, recv_counts)."""
    world = dist.get_world_size(group)
    dev = send.device
    sc = torch.tensor([int(c) for c in send_counts], dtype=torch.int64, device=dev)
    rc = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.cpu().tolist()]
    recv = torch.empty(sum(recv_counts) * TUPLE_BYTES, dtype=torch.uint8, device=dev)
    dist.all_to_all_single(recv, send, output_split_sizes=[c * TUPLE_BYTES for c in recv_counts],
                           input_split_sizes=[int(c) * TUPLE_BYTES for c in send_counts], group=group)
    return recv, recv_counts


class _DevMem:
    """expose a raw device pointer to torch through __cuda_array_interface__"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def build_index_distributed(spec, local_sids, local_seqs, frg_id_mode=api.FRG_ID_FASTX, device=None, group=None):
    """Every rank passes its contiguous block of the global sequence list; returns this rank's ShmmrIndex slice
    (keys of the rank's key range) and a dict of timings / counts.  Must be called by all ranks of `group`."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev_index = torch.cuda.current_device() if device is None else device
    dev = torch.device("cuda", dev_index)
    idx = api.ShmmrIndex(spec, frg_id_mode, dev_index)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    n_frags = idx.stage_batch(local_sids, local_seqs)
    base, total_frags = frag_bases(n_frags, group, dev)
    idx.commit_batch(base)
    ev[1].record()
    ptr, n = idx.tuples_device()
    if world > 1:
        # sample h0 of every stride-th tuple (first 8 bytes of each 40-byte record)
        if n:
            raw = torch.as_tensor(_DevMem(ptr, n * TUPLE_BYTES), device=dev)
            stride = max(1, n // SAMPLES_PER_RANK)
            sample = raw.view(n, TUPLE_BYTES)[::stride, :8].contiguous().view(torch.int64).flatten().cpu().numpy().view(np.uint64)
        else:
            sample = np.zeros(0, dtype=np.uint64)
        splitters = choose_splitters(sample, group, dev)
        counts = idx.partition(splitters)
        ptr, n = idx.tuples_device()
        send = torch.as_tensor(_DevMem(ptr, n * TUPLE_BYTES), device=dev) if n else torch.empty(0, dtype=torch.uint8, device=dev)
        ev[2].record()
        recv, recv_counts = exchange_records(send, counts, group)
        torch.cuda.synchronize()
        idx.set_tuples_device(recv.data_ptr(), sum(recv_counts))
        del recv
    else:
        ev[2].record()
    idx.finalize()
    ev[3].record()
    torch.cuda.synchronize()
    nk, ns, _ = idx.counts()
    info = {"stage_commit_ms": ev[0].elapsed_time(ev[1]), "partition_ms": ev[1].elapsed_time(ev[2]),
            "exchange_sort_ms": ev[2].elapsed_time(ev[3]), "n_keys": nk, "n_sigs": ns, "total_frags": total_frags}
    return idx, info
