"""A/B in one process: c_fc -> c_proj with tile-level dependencies (default) against plain grid-level dependencies
(grid_level_deps=True), on 64-window batches (BASELINE configs[1]) and on 2048x1536 images (configs[2]); alternating >= 1.5 s
timed regions, CUDA events.   python profiles/tile_deps_ab.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clip_ebc_b200 import get_model, sliding_window_predict  # noqa: E402
from oracle import weights  # noqa: E402

dev = torch.device("cuda", 0)
reduction, bins, anchors = weights.bins_and_anchors("r8_t4_nwpu")
sd = weights.make_state_dict(0)
tf = weights.make_text_features(len(bins), seed=100)


def build(grid_level):
    m = get_model("clip_vit_b_16", input_size=224, reduction=reduction, bins=bins, anchor_points=anchors, prompt_type="word",
                  num_vpt=32, vpt_drop=0.0, deep_vpt=True, text_features=tf, grid_level_deps=grid_level)
    m.load_state_dict(sd, strict=True)
    return m.to(dev).eval()


models = {"tile-level": build(False), "grid-level": build(True)}
xs = [weights.make_image((64, 3, 224, 224), seed=50 + i).to(dev) for i in range(4)]
img = [weights.make_image((1, 3, 1536, 2048), seed=60 + i).to(dev) for i in range(2)]


def timed(fn, seconds=1.5):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(0); e1.record(); torch.cuda.synchronize()
    n = max(5, int(seconds * 1000.0 / e0.elapsed_time(e1)))
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for rnd in range(3):
    for name, m in models.items():
        a = timed(lambda i: m(xs[i % 4]))
        b = timed(lambda i: sliding_window_predict(m, img[i % 2], 224, 112, return_device=True))
        print(f"round {rnd} {name:10s}: 64 windows {a:.3f} ms = {64 / a * 1e3:7.0f} windows/s | 2048x1536 {b:.2f} ms = {234 / b * 1e3:7.0f} windows/s")
