"""numpy/torch model of the library's multi-GPU merge (pgr_tk_b200/csrc/shard.cu) used by the CPU (gloo) tests:
fragment-id bases, splitters as quantiles of an all-gathered sample, the all-to-all plan.  Test infrastructure."""
import numpy as np
import torch
import torch.distributed as dist

TUPLE_BYTES = 40
SAMPLES_PER_RANK = 2048


def frag_bases(n_frags_local, group=None):
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t = torch.tensor([int(n_frags_local)], dtype=torch.int64)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    vals = [int(x.item()) for x in out]
    return sum(vals[:rank]), sum(vals)


def sample_h0(h0, n_sample=SAMPLES_PER_RANK):
    """sample_h0_kernel: entry i = h0[(i * n) // n_sample], ~0 when the shard is empty"""
    n = len(h0)
    if n == 0:
        return np.full(n_sample, np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
    return np.asarray(h0, dtype=np.uint64)[(np.arange(n_sample, dtype=np.uint64) * np.uint64(n)) // np.uint64(n_sample)]


def choose_splitters(local_sample, group=None):
    """splitters_kernel: sort the gathered sample, drop the ~0 fillers, cut at i*m/world"""
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.asarray(local_sample, dtype=np.uint64).view(np.int64).copy())
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    allv = np.sort(np.concatenate([o.numpy().view(np.uint64) for o in out]))
    allv = allv[allv != np.uint64(0xFFFFFFFFFFFFFFFF)]
    m = len(allv)
    if m == 0:
        return np.zeros(world - 1, dtype=np.uint64)
    q = [min(m - 1, (m * (i + 1)) // world) for i in range(world - 1)]
    return allv[q].astype(np.uint64)


def exchange_records(send, send_counts, group=None):
    """all-to-all of 40-byte records already ordered by destination rank -> (recv uint8 tensor, recv_counts)"""
    world = dist.get_world_size(group)
    sc = torch.tensor([int(c) for c in send_counts], dtype=torch.int64)
    rc = torch.zeros(world, dtype=torch.int64)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.tolist()]
    recv = torch.empty(sum(recv_counts) * TUPLE_BYTES, dtype=torch.uint8)
    dist.all_to_all_single(recv, send, output_split_sizes=[c * TUPLE_BYTES for c in recv_counts],
                           input_split_sizes=[int(c) * TUPLE_BYTES for c in send_counts], group=group)
    return recv, recv_counts
