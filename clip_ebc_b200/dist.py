"""Multi-GPU use of the hot path: whole images are partitioned across the ranks of one node (one process per GPU);
the only collective is a single all-gather of per-image counts at the end.

The reference evaluates on rank 0 only (trainer.py:161-179; eval.py:25-35 / test_nwpu.py:89-116 loop over images one by
one), so this layer is new. Image i goes to rank ``i % world_size``; every rank runs ``sliding_window_predict`` on its
images (device-resident result + fused count) and the counts are exchanged with one ``all_gather`` (NCCL over NVLink on
GPUs, gloo in the CPU tests). Density maps are never exchanged.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world_size: int) -> List[int]:
    """Round-robin partition: image i -> rank i % world_size."""
    return list(range(rank, n_items, world_size))


def gather_counts(local_counts: torch.Tensor, n_items: int, rank: int, world_size: int) -> torch.Tensor:
    """One all-gather of fp32 counts, padded to ceil(n_items / world_size) per rank with NaN, re-ordered to image index.

    local_counts: 1-D fp32 tensor with the counts of shard_indices(n_items, rank, world_size), on the device the
    process group communicates from (CUDA for nccl, CPU for gloo)."""
    per_rank = (n_items + world_size - 1) // world_size
    send = torch.full((per_rank,), float("nan"), dtype=torch.float32, device=local_counts.device)
    send[: local_counts.numel()] = local_counts
    if world_size == 1 or not (dist.is_available() and dist.is_initialized()):
        gathered = send.view(1, per_rank)
    else:
        recv = torch.empty((world_size * per_rank,), dtype=torch.float32, device=local_counts.device)
        dist.all_gather_into_tensor(recv, send)
        gathered = recv.view(world_size, per_rank)
    out = torch.empty((n_items,), dtype=torch.float32, device=local_counts.device)
    for r in range(world_size):
        idx = shard_indices(n_items, r, world_size)
        out[idx] = gathered[r, : len(idx)]
    return out


def predict_counts(predict_one: Callable[[int], torch.Tensor], n_items: int, rank: int, world_size: int) -> torch.Tensor:
    """Run ``predict_one(i) -> count tensor [1]`` on this rank's images and return all n_items counts on every rank."""
    mine = shard_indices(n_items, rank, world_size)
    counts = [predict_one(i).reshape(1) for i in mine]
    local = torch.cat(counts) if counts else torch.empty((0,), dtype=torch.float32)
    return gather_counts(local, n_items, rank, world_size)
