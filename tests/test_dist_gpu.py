"""N-rank correctness on hardware: image-sharded sliding_window_predict + the single NCCL all-gather
of per-image counts gives, on every rank, exactly the bits the 1-GPU path gives. Spawns `torch.distributed.run` with one
process per GPU; skipped when the box has fewer GPUs than ranks (the worker is tests/dist_worker.py)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_counts_equal_single_gpu_counts_bit_for_bit(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    out = tmp_path / "dist.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py"), str(out)]
    res = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:]
    rec = json.load(open(out))
    print(f"\n[{world} ranks] counts {rec['counts'][:4]} ... bit-exact on every rank: {rec['bit_exact_on_every_rank']}")
    assert rec["world"] == world and rec["finite"] and rec["bit_exact_on_every_rank"]
    assert rec["counts"] == rec["counts_1gpu"]
    assert rec["balanced_partition_bit_exact"] and rec["one_image_window_sharded_bit_exact"]
    print(f"one 672x896 image, stride 112 (35 windows): {rec['one_image_672x896_s112_ms']} ms")
    keep = os.environ.get("CLIPEBC_DIST_RECORD_DIR")  # gpurun sessions keep the record under profiles/
    if keep:
        os.makedirs(keep, exist_ok=True)
        json.dump(rec, open(os.path.join(keep, f"dist_bit_exact_{world}gpu.json"), "w"), indent=1)


def test_window_sharded_path_on_one_rank_equals_sliding_window_predict():
    """`dist.sliding_window_predict_sharded` with a world of one (no process group): window crops -> model(x) -> the fold
    entry point gives the bits of the one-call sliding path, on grid-aligned and off-grid windows."""
    from clip_ebc_b200 import get_model, sliding_window_predict
    from clip_ebc_b200.dist import sliding_window_predict_sharded
    from oracle import weights

    reduction, bins, anchors = weights.bins_and_anchors("r8_t4_nwpu")
    tf = weights.make_text_features(len(bins), seed=103)
    model = get_model("clip_vit_b_16", input_size=224, reduction=reduction, bins=bins, anchor_points=anchors,
                      prompt_type="word", num_vpt=32, vpt_drop=0.0, deep_vpt=True, text_features=tf)
    model.load_state_dict(weights.make_state_dict(3, variant="stress"), strict=True)
    model = model.to("cuda:0").eval()
    for shape, stride in (((1, 3, 448, 672), 112), ((1, 3, 300, 500), 200), ((1, 3, 224, 224), 224)):
        img = weights.make_image(shape, seed=91).to("cuda:0")
        d1, c1 = sliding_window_predict(model, img, 224, stride, return_device=True, return_count=True)
        d2, c2 = sliding_window_predict_sharded(model, img, 224, stride, 0, 1, return_count=True)
        assert d1.shape == d2.shape
        assert torch.equal(d1.view(torch.int32), d2.view(torch.int32)), shape
        assert torch.equal(c1.view(torch.int32), c2.view(torch.int32)), shape


def test_two_devices_in_one_process():
    """One process, two GPUs (round-1 ADVICE): a native handle lives on the device it was created on; per-device kernel
    attributes (dynamic shared memory, SM count, cluster occupancy, L2 carve-out) are kept per device. A model on cuda:1 next
    to one on cuda:0 gives the same bits, `.to("cuda:1")` re-creates the handle there, and both keep working afterwards."""
    if torch.cuda.device_count() < 2:
        pytest.skip(f"needs 2 GPUs, this box has {torch.cuda.device_count()}")
    from clip_ebc_b200 import get_model, sliding_window_predict
    from oracle import weights

    reduction, bins, anchors = weights.bins_and_anchors("r8_t4_nwpu")
    sd = weights.make_state_dict(3, variant="stress")
    tf = weights.make_text_features(len(bins), seed=103)

    def build(dev):
        m = get_model("clip_vit_b_16", input_size=224, reduction=reduction, bins=bins, anchor_points=anchors,
                      prompt_type="word", num_vpt=32, vpt_drop=0.0, deep_vpt=True, text_features=tf)
        m.load_state_dict(sd, strict=True)
        return m.to(dev).eval()

    x = weights.make_image((5, 3, 224, 224), seed=800)
    img = weights.make_image((1, 3, 448, 672), seed=801)
    m0, m1 = build("cuda:0"), build("cuda:1")
    y0 = m0(x.to("cuda:0"))
    y1 = m1(x.to("cuda:1"))                      # first use of every kernel on the second device
    assert y1.device.index == 1 and torch.equal(y0.cpu(), y1.cpu())
    d0 = sliding_window_predict(m0, img, 224, 112)
    d1 = sliding_window_predict(m1, img, 224, 112)
    assert torch.equal(d0, d1)
    x2 = weights.make_image((5, 3, 224, 224), seed=802)
    m0.use_cuda_graphs = False
    y2 = m0(x2.to("cuda:0")).cpu()
    m0.use_cuda_graphs = True
    assert not torch.equal(y2, y0.cpu())
    for i in range(4):                           # interleaved calls with alternating inputs: graph capture / replay on both devices
        xi, yi = (x, y0.cpu()) if i % 2 == 0 else (x2, y2)
        assert torch.equal(m0(xi.to("cuda:0")).cpu(), yi)
        assert torch.equal(m1(xi.to("cuda:1")).cpu(), yi)
    assert any("graph" in e for e in m1._graph_cache.values()), "no graph was captured on the second device"
    m0 = m0.to("cuda:1")                         # the handle is re-created on the new device, the old buffers are released
    assert torch.equal(m0(x.to("cuda:1")).cpu(), y0.cpu())
    with pytest.raises(RuntimeError):            # an input on another device than the model is refused, not dereferenced
        m0(x.to("cuda:0"))
    m0 = m0.to("cuda:0")
    assert torch.equal(m0(x.to("cuda:0")).cpu(), y0.cpu())
    # a CLIP-ResNet model on the second device as well
    sd_rn = weights.make_resnet_state_dict(40, "resnet50", "stress")
    tf_rn = weights.make_text_features(len(bins), seed=140, embed=1024)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        rn = get_model("clip_resnet50", input_size=224, reduction=reduction, bins=bins, anchor_points=anchors, text_features=tf_rn)
        rn.load_state_dict(sd_rn, strict=True)
        outs.append(rn.to(dev).eval()(x[:2].to(dev)).cpu())
    assert torch.equal(outs[0], outs[1])
