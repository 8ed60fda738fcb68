"""Multi-GPU use of the hot path: whole images are partitioned across the ranks of one node (one process per GPU);
the only collective is a single all-gather of per-image counts at the end. For ONE large image whose latency matters,
``sliding_window_predict_sharded`` partitions the image's windows instead and all-gathers the per-window maps.

The reference evaluates on rank 0 only (trainer.py:161-179; eval.py:25-35 / test_nwpu.py:89-116 loop over images one by
one), so this layer is new. Image i goes to rank ``i % world_size``; every rank runs ``sliding_window_predict`` on its
images (device-resident result + fused count) and the counts are exchanged with one ``all_gather`` (NCCL over NVLink on
GPUs, gloo in the CPU tests). Density maps are never exchanged.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world_size: int) -> List[int]:
    """Round-robin partition: image i -> rank i % world_size."""
    return list(range(rank, n_items, world_size))


def shard_balanced(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Partition for images of different sizes: longest-processing-time greedy on a per-image cost, e.g. its
    window count. Images are taken by decreasing cost (ties: lower index first) and dealt to the least-loaded rank (ties:
    lowest rank), so every rank computes the same partition; each rank's list is returned in ascending image order."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += float(costs[i])
    return [sorted(sh) for sh in shards]


def gather_counts(local_counts: torch.Tensor, n_items: int, rank: int, world_size: int,
                  shards: Optional[List[List[int]]] = None) -> torch.Tensor:
    """One all-gather of fp32 counts, padded to the longest shard with NaN, re-ordered to image index.

    local_counts: 1-D fp32 tensor with the counts of this rank's shard (shard_indices(n_items, rank, world_size), or
    shards[rank] when a partition is given), on the device the process group communicates from (CUDA for nccl, CPU for
    gloo)."""
    if shards is None:
        shards = [shard_indices(n_items, r, world_size) for r in range(world_size)]
    per_rank = max(max((len(sh) for sh in shards), default=0), 1)
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
        idx = shards[r]
        if idx:
            out[idx] = gathered[r, : len(idx)]
    return out


def predict_counts(predict_one: Callable[[int], torch.Tensor], n_items: int, rank: int, world_size: int,
                   costs: Optional[Sequence[float]] = None) -> torch.Tensor:
    """Run ``predict_one(i) -> count tensor [1]`` on this rank's images and return all n_items counts on every rank.

    costs: optional per-image cost (window count) for a balanced partition of images of different sizes; the default is
    the round-robin partition. Counts do not depend on the partition (an image is always processed whole, on one GPU)."""
    shards = shard_balanced(costs, world_size) if costs is not None else None
    if costs is not None:
        assert len(costs) == n_items, f"expected {n_items} costs, got {len(costs)}"
    mine = shards[rank] if shards is not None else shard_indices(n_items, rank, world_size)
    counts = [predict_one(i).reshape(1) for i in mine]
    local = torch.cat(counts) if counts else torch.empty((0,), dtype=torch.float32)
    return gather_counts(local, n_items, rank, world_size, shards)


# ------------------------------------------------------------------------------------------ one image, many GPUs
def window_shard(n_windows: int, rank: int, world_size: int) -> Tuple[int, int, int]:
    """Contiguous range [lo, hi) of the row-major window list owned by `rank`, and the padded per-rank length."""
    per = (n_windows + world_size - 1) // world_size
    lo = min(rank * per, n_windows)
    return lo, min(lo + per, n_windows), per


def gather_windows(local: torch.Tensor, n_windows: int, rank: int, world_size: int) -> torch.Tensor:
    """All-gather of the per-window maps: local [hi - lo, 1, g, g] of window_shard(n_windows, rank, world_size) ->
    [n_windows, 1, g, g] in window order on every rank (n_windows * g * g * 4 bytes in total: 3 MB for the 972 windows of a
    4096 x 3072 image at stride 112)."""
    lo, hi, per = window_shard(n_windows, rank, world_size)
    assert local.shape[0] == hi - lo, f"rank {rank} owns {hi - lo} windows, got {local.shape[0]}"
    if world_size == 1 or not (dist.is_available() and dist.is_initialized()):
        return local
    send = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    send[: hi - lo] = local
    recv = torch.empty((world_size * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send)
    return recv[:n_windows]


def sliding_window_predict_sharded(model, image: torch.Tensor, window_size: Union[int, Tuple[int, int]],
                                   stride: Union[int, Tuple[int, int]], rank: int, world_size: int,
                                   return_count: bool = False):
    """``sliding_window_predict`` of ONE image on all ranks of the process group: the latency path for a single large image
    ("alternative for one huge image"). Rank r runs the windows window_shard(n, r, W) through ``model(x)``,
    the per-window maps are exchanged with one all-gather, and every rank folds all of them in the 1-GPU order
    (ascending window index per cell), so the map is bit-identical to the 1-GPU ``sliding_window_predict`` -- the
    partial sums of a reduce-scatter would not be.

    image: [1, 3, H, W] on this rank's device (every rank passes the same image). Returns the density map [1, 1, H/r, W/r]
    on the device (and its sum with return_count), identical on every rank."""
    from . import ops
    from .eval_utils import _pair

    assert len(image.shape) == 4, f"Image must be a 4D tensor (1, c, h, w), got {image.shape}"
    assert image.shape[0] == 1, f"The batch size must be 1 due to varying image sizes, got {image.shape[0]}"
    window_size = _pair(window_size, "Window size")
    stride = _pair(stride, "Stride")
    assert stride[0] <= window_size[0] and stride[1] <= window_size[1], \
        f"Stride must be smaller than window size, got {stride} and {window_size}"
    model.eval()
    H, W = int(image.shape[-2]), int(image.shape[-1])
    red = int(model.reduction)
    rows, cols = ops.window_origins(H, W, window_size, stride)
    n_win = len(rows) * len(cols)
    lo, hi, _ = window_shard(n_win, rank, world_size)
    gh, gw = window_size[0] // red, window_size[1] // red
    with torch.no_grad():
        if hi > lo:
            crops = torch.cat([image[:, :, rows[w // len(cols)]: rows[w // len(cols)] + window_size[0],
                                     cols[w % len(cols)]: cols[w % len(cols)] + window_size[1]] for w in range(lo, hi)])
            local = model(crops.contiguous())
        else:
            local = torch.empty((0, 1, gh, gw), dtype=torch.float32, device=image.device)
        preds = gather_windows(local.contiguous(), n_win, rank, world_size)
        out = ops.fold_average(preds.contiguous(), [r // red for r in rows], [c // red for c in cols], H // red, W // red,
                               want_count=return_count)
    dens, cnt = out if return_count else (out, None)
    dens = dens[None, None]
    return (dens, cnt) if return_count else dens
