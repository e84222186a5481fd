"""Length-balanced sharding of files / windows over the GPUs of one box, and the
final gather.

The reference parallelises by audio file: ``n_process`` copies of the script claim a
file by creating its empty result TSV (/root/reference/align_utterances.sh:127-137,
src/iterative_utterance_alignment.py:436-447 -- check-then-create, racy).  Here the
assignment is deterministic and computed identically on every rank: longest
processing time first onto the least loaded rank.  Units are independent, so
there is NO collective on the data path; ``torch.distributed`` (NCCL between
GPUs, gloo in the CPU tests) is used only to gather per-unit results.
"""
import heapq

import numpy as np
import torch
import torch.distributed as dist


def lpt_shards(costs, n_shards):
    """Greedy LPT: returns ``n_shards`` lists of unit indices, each sorted ascending.
    Deterministic (ties broken by index), so every rank computes the same answer."""
    costs = np.asarray(costs, dtype=np.float64)
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0.0, r) for r in range(n_shards)]
    heapq.heapify(heap)
    shards = [[] for _ in range(n_shards)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + float(costs[i]), r))
    return [sorted(s) for s in shards]


def shard_imbalance(costs, shards):
    """max shard load / mean shard load (1.0 = perfect)."""
    costs = np.asarray(costs, dtype=np.float64)
    loads = np.array([costs[s].sum() if len(s) else 0.0 for s in shards])
    return float(loads.max() / max(loads.mean(), 1e-300))


def lattice_cost(in_len, tgt_len):
    """Cost of a window on the 2L+1 lattice: T * (2L + 1) cells."""
    return np.asarray(in_len, np.float64) * (2.0 * np.asarray(tgt_len, np.float64) + 1.0)


def my_shard(costs, world_size=None, rank=None):
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    return lpt_shards(costs, world_size)[rank]


def gather_rows(local_values, local_index, n_total, group=None):
    """Gather per-unit fixed-width results from every rank into the global order.

    ``local_values``: tensor [n_local, ...] (CUDA with NCCL, CPU with gloo),
    ``local_index``: global unit index of each local row.  Ragged shard sizes are
    handled by a counts all-gather + padded ``all_gather_into_tensor``.  Every rank
    receives the full [n_total, ...] tensor."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = torch.zeros((n_total,) + tuple(local_values.shape[1:]), dtype=local_values.dtype,
                          device=local_values.device)
        out[torch.as_tensor(local_index, dtype=torch.long, device=local_values.device)] = local_values
        return out
    world = dist.get_world_size(group)
    dev = local_values.device
    n_local = torch.tensor([local_values.shape[0]], dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, n_local, group=group)
    n_max = int(counts.max())
    tail = tuple(local_values.shape[1:])
    padded = torch.zeros((n_max,) + tail, dtype=local_values.dtype, device=dev)
    padded[:local_values.shape[0]] = local_values
    idx = torch.full((n_max,), -1, dtype=torch.int64, device=dev)
    idx[:local_values.shape[0]] = torch.as_tensor(local_index, dtype=torch.int64, device=dev)
    all_vals = torch.zeros((world * n_max,) + tail, dtype=local_values.dtype, device=dev)
    all_idx = torch.zeros(world * n_max, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_vals, padded, group=group)
    dist.all_gather_into_tensor(all_idx, idx, group=group)
    keep = all_idx >= 0
    out = torch.zeros((n_total,) + tail, dtype=local_values.dtype, device=dev)
    out[all_idx[keep]] = all_vals[keep]
    return out


def gather_objects(local_obj, group=None):
    """Gather arbitrary per-rank Python results (TSV rows) on every rank, rank order."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [local_obj]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local_obj, group=group)
    return out
