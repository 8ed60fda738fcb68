"""Windows per internal pass on the window workloads: model(x) on B windows of 224x224 for a few window_chunk values, >= 1 s per
point.  python profiles/window_chunk_sweep.py <backbone> <bins> <deep 0|1> <B> <chunk> [<chunk> ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clip_ebc_b200 import get_model  # noqa: E402
from oracle import weights  # noqa: E402

backbone, bins_name, deep, B = sys.argv[1], sys.argv[2], bool(int(sys.argv[3])), int(sys.argv[4])
chunks = [int(c) for c in sys.argv[5:]]
patch = {"clip_vit_b_32": 32, "clip_vit_b_16": 16, "clip_vit_l_14": 14}[backbone]
dev = torch.device("cuda", 0)
reduction, bins, anchors = weights.bins_and_anchors(bins_name)
sd = weights.make_state_dict(0, input_size=224, num_vpt=32, deep_vpt=deep, variant="default", patch=patch)
tf = weights.make_text_features(len(bins), seed=100, embed=768 if patch == 14 else 512)
xs = [weights.make_image((B, 3, 224, 224), seed=70 + i).to(dev) for i in range(2)]
for chunk in chunks:
    model = get_model(backbone, input_size=224, reduction=reduction, bins=bins, anchor_points=anchors, prompt_type="word",
                      num_vpt=32, vpt_drop=0.0, deep_vpt=deep, text_features=tf, window_chunk=chunk)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    for i in range(4):
        model(xs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); model(xs[0]); e1.record(); torch.cuda.synchronize()
    n = max(8, int(1200.0 / e0.elapsed_time(e1)))
    e0.record()
    for i in range(n):
        model(xs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{backbone} {bins_name} deep={int(deep)} B={B} window_chunk {chunk:3d}: {ms:7.3f} ms -> {B / ms * 1e3:6.0f} windows/s", flush=True)
    del model
    torch.cuda.empty_cache()
