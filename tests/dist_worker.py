"""Worker of tests/test_dist_gpu.py, launched once per GPU by `python -m torch.distributed.run` (not collected by pytest).

SURVEY.md section 8(e): the reference evaluates on rank 0 only (trainer.py:161-179), so multi-GPU correctness is
"W-GPU counts == 1-GPU counts bit for bit". Every rank predicts its shard of N_IMAGES synthetic images (image i -> rank
i % W, clip_ebc_b200.dist), the per-image counts are exchanged with the path's single NCCL all-gather, and rank 0 also
predicts all images on its own GPU (the 1-rank path). The two count vectors must be identical bit for bit, on every rank.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

N_IMAGES = 16
SIZES = [(448, 672), (672, 448), (448, 448), (560, 784)]  # image i has SIZES[i % 4]: 15 / 15 / 9 / 24 windows at stride 112


def main() -> int:
    out_path = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    from clip_ebc_b200 import get_model, sliding_window_predict
    from clip_ebc_b200.dist import predict_counts, shard_indices, sliding_window_predict_sharded
    from oracle import weights

    reduction, bins, anchors = weights.bins_and_anchors("r8_t4_nwpu")
    sd = weights.make_state_dict(3, variant="stress")
    tf = weights.make_text_features(len(bins), seed=103)
    model = get_model("clip_vit_b_16", input_size=224, reduction=reduction, bins=bins, anchor_points=anchors,
                      prompt_type="word", num_vpt=32, vpt_drop=0.0, deep_vpt=True, text_features=tf)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()

    def image(i):
        h, w = SIZES[i % len(SIZES)]
        return weights.make_image((1, 3, h, w), seed=500 + i)

    def predict_one(i):
        return sliding_window_predict(model, image(i).to(dev), 224, 112, return_device=True, return_count=True)[1]

    counts = predict_counts(predict_one, N_IMAGES, rank, world)  # shard + ONE all-gather (NCCL)
    assert counts.shape == (N_IMAGES,) and counts.is_cuda
    # the 1-rank path on rank 0's GPU, broadcast so that every rank checks its own gathered vector
    single = torch.empty(N_IMAGES, dtype=torch.float32, device=dev)
    if rank == 0:
        single = predict_counts(predict_one, N_IMAGES, 0, 1)
    dist.broadcast(single, src=0)
    same = bool(torch.equal(counts.view(torch.int32), single.view(torch.int32)))  # bit for bit
    flags = torch.tensor([int(same)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    # balanced partition (longest-processing-time on the window counts): the same counts, whoever computes them
    n_windows = [{(448, 672): 15, (672, 448): 15, (448, 448): 9, (560, 784): 24}[SIZES[i % len(SIZES)]] for i in range(N_IMAGES)]
    balanced = predict_counts(predict_one, N_IMAGES, rank, world, costs=n_windows)
    same_b = bool(torch.equal(balanced.view(torch.int32), single.view(torch.int32)))
    # ONE image on all ranks: windows sharded, per-window maps all-gathered, folded in the 1-GPU order on every rank
    big = weights.make_image((1, 3, 672, 896), seed=777).to(dev)  # 5 x 7 = 35 windows at stride 112
    dens_s, cnt_s = sliding_window_predict_sharded(model, big, 224, 112, rank, world, return_count=True)
    dens_1, cnt_1 = sliding_window_predict(model, big, 224, 112, return_device=True, return_count=True)
    same_w = bool(torch.equal(dens_s.view(torch.int32), dens_1.view(torch.int32)) and
                  torch.equal(cnt_s.view(torch.int32), cnt_1.view(torch.int32)) and dens_s.shape == dens_1.shape)
    torch.cuda.synchronize()
    dist.barrier()
    lat = {}
    for name, fn in (("sharded", lambda: sliding_window_predict_sharded(model, big, 224, 112, rank, world)),
                     ("one_gpu", lambda: sliding_window_predict(model, big, 224, 112, return_device=True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lat[name] = float(t.item())
    flags2 = torch.tensor([int(same_b and same_w)], device=dev)
    dist.all_reduce(flags2, op=dist.ReduceOp.MIN)
    flags = torch.minimum(flags, flags2)
    if rank == 0:
        json.dump({"world": world, "n_images": N_IMAGES, "bit_exact_on_every_rank": bool(flags.item()),
                   "balanced_partition_bit_exact": same_b, "one_image_window_sharded_bit_exact": same_w,
                   "one_image_672x896_s112_ms": lat,
                   "my_images_rank0": shard_indices(N_IMAGES, 0, world), "counts": counts.tolist(),
                   "counts_1gpu": single.tolist(), "finite": bool(torch.isfinite(counts).all())}, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if flags.item() == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
