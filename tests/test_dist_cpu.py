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
