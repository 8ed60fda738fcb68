"""Windows per internal pass for ViT-L/14 (257 live rows of width 1024 per window): model(x) on 292 windows for a few
window_chunk values, >= 1 s per point (python profiles/l14_chunk_sweep.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clip_ebc_b200 import get_model  # noqa: E402
from oracle import weights  # noqa: E402

dev = torch.device("cuda", 0)
reduction, bins, anchors = weights.bins_and_anchors("r8_t4_nwpu")
sd = weights.make_state_dict(0, input_size=224, num_vpt=32, deep_vpt=True, variant="default", patch=14)
tf = weights.make_text_features(len(bins), seed=100, embed=768)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 292
xs = [weights.make_image((B, 3, 224, 224), seed=70 + i).to(dev) for i in range(2)]
for chunk in (0, 48, 64, 73, 96, 146):
    model = get_model("clip_vit_l_14", input_size=224, reduction=reduction, bins=bins, anchor_points=anchors, prompt_type="word",
                      num_vpt=32, vpt_drop=0.0, deep_vpt=True, text_features=tf, window_chunk=chunk)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    for i in range(3):
        model(xs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 16
    e0.record()
    for i in range(n):
        model(xs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"window_chunk {chunk:3d}: {B} windows in {ms:7.2f} ms -> {B / ms * 1e3:6.0f} windows/s", flush=True)
    del model
    torch.cuda.empty_cache()
