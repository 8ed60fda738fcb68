"""Throughput of the three ViT backbones behind the same entry point: model(x) on windows of 224x224, >= 1 s timed per backbone
(python profiles/backbones_bench.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clip_ebc_b200 import get_model  # noqa: E402
from oracle import weights  # noqa: E402

dev = torch.device("cuda", 0)
reduction, bins, anchors = weights.bins_and_anchors("r8_t4_nwpu")
for backbone, patch, B in (("clip_vit_b_32", 32, 256), ("clip_vit_b_16", 16, 256), ("clip_vit_l_14", 14, 96)):
    sd = weights.make_state_dict(0, input_size=224, num_vpt=32, deep_vpt=True, variant="default", patch=patch)
    tf = weights.make_text_features(len(bins), seed=100, embed=768 if patch == 14 else 512)
    model = get_model(backbone, input_size=224, reduction=reduction, bins=bins, anchor_points=anchors, prompt_type="word",
                      num_vpt=32, vpt_drop=0.0, deep_vpt=True, text_features=tf)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    xs = [weights.make_image((B, 3, 224, 224), seed=70 + i).to(dev) for i in range(2)]
    for i in range(6):
        model(xs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); model(xs[0]); e1.record(); torch.cuda.synchronize()
    n = max(10, int(1000.0 / e0.elapsed_time(e1)))
    e0.record()
    for i in range(n):
        model(xs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{backbone}: {B} windows in {ms:.2f} ms -> {B / ms * 1e3:.0f} windows/s ({n} passes)")
    del model, xs
    torch.cuda.empty_cache()
