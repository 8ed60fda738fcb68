"""world_size-2 gloo test (CPU) of the multi-GPU host logic: round-robin image sharding + the single count all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from clip_ebc_b200.dist import predict_counts, shard_indices

    mine = shard_indices(n_items, rank, world)
    out = predict_counts(lambda i: torch.tensor([float(i) * 1.5 + 0.25]), n_items, rank, world)
    q.put((rank, mine, out.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [5, 8, 1])
def test_shard_and_gather_world2(n_items):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [i * 1.5 + 0.25 for i in range(n_items)]
    seen = []
    for rank, mine, out in res:
        assert out == expect  # every rank holds all counts, in image order
        assert mine == list(range(rank, n_items, world))
        seen += mine
    assert sorted(seen) == list(range(n_items))  # a partition: every image on exactly one rank


def test_single_process_path():
    from clip_ebc_b200.dist import predict_counts

    out = predict_counts(lambda i: torch.tensor([float(i)]), 4, 0, 1)
    assert out.tolist() == [0.0, 1.0, 2.0, 3.0]


def test_balanced_partition_is_deterministic_and_balanced():
    from clip_ebc_b200.dist import shard_balanced

    costs = [266, 15, 15, 9, 24, 972, 234, 4, 4, 60, 15, 9]  # window counts of images of different sizes
    for world in (1, 2, 3, 4, 8):
        shards = shard_balanced(costs, world)
        assert sorted(i for sh in shards for i in sh) == list(range(len(costs)))  # a partition
        assert all(sh == sorted(sh) for sh in shards)
        loads = [sum(costs[i] for i in sh) for sh in shards]
        assert max(loads) <= max(max(costs), sum(costs) / world * 4 / 3 + 1)  # the LPT bound
        assert shards == shard_balanced(list(costs), world)  # same input, same answer (every rank computes it)
    assert shard_balanced([], 4) == [[], [], [], []]
    assert shard_balanced([1, 1, 1, 1], 2) == [[0, 2], [1, 3]]  # ties: lower index first, lowest rank first


def _worker_balanced(rank, world, port, costs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from clip_ebc_b200.dist import gather_windows, predict_counts, window_shard

    out = predict_counts(lambda i: torch.tensor([float(i) * 2.0 - 1.0]), len(costs), rank, world, costs=costs)
    # one image split by windows: every rank contributes its contiguous range, all ranks end with the full list in order
    n_win = 7
    lo, hi, per = window_shard(n_win, rank, world)
    local = torch.stack([torch.full((1, 2, 3), float(w)) for w in range(lo, hi)]) if hi > lo else torch.empty((0, 1, 2, 3))
    allw = gather_windows(local, n_win, rank, world)
    q.put((rank, out.tolist(), allw[:, 0, 0, 0].tolist(), tuple(allw.shape), (lo, hi, per)))
    dist.barrier()
    dist.destroy_process_group()


def test_balanced_counts_and_window_gather_world2():
    world = 2
    costs = [972.0, 9.0, 15.0, 266.0, 24.0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_balanced, args=(r, world, port, costs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, win_ids, shape, rng in res:
        assert out == [i * 2.0 - 1.0 for i in range(len(costs))]
        assert win_ids == [float(w) for w in range(7)] and shape == (7, 1, 2, 3)
        assert rng == ((0, 4, 4) if rank == 0 else (4, 7, 4))


def test_window_shard_edges():
    from clip_ebc_b200.dist import window_shard

    assert window_shard(4, 0, 8) == (0, 1, 1) and window_shard(4, 5, 8) == (4, 4, 1)  # more ranks than windows
    assert [window_shard(972, r, 8)[:2] for r in (0, 7)] == [(0, 122), (854, 972)]
    assert window_shard(10, 0, 1) == (0, 10, 10)
