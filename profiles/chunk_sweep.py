"""Windows per internal pass (`window_chunk`) on the image workloads: images/s of sliding_window_predict for a few chunk sizes,
>= 1.5 s per point (python profiles/chunk_sweep.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clip_ebc_b200 import get_model, sliding_window_predict  # noqa: E402
from oracle import weights  # noqa: E402

dev = torch.device("cuda", 0)
reduction, bins, anchors = weights.bins_and_anchors("r8_t4_nwpu")
sd = weights.make_state_dict(0, input_size=224, num_vpt=32, deep_vpt=True, variant="default")
tf = weights.make_text_features(len(bins), seed=100)
for (H, W, stride) in ((1536, 2048, 112), (3072, 4096, 224)):
    imgs = [weights.make_image((1, 3, H, W), seed=600 + i).to(dev) for i in range(2)]
    for chunk in (0, 64, 96, 117 if stride == 112 else 133, 128, 148, 192, 234, 266):
        model = get_model("clip_vit_b_16", input_size=224, reduction=reduction, bins=bins, anchor_points=anchors, prompt_type="word",
                          num_vpt=32, vpt_drop=0.0, deep_vpt=True, text_features=tf, window_chunk=chunk)
        model.load_state_dict(sd, strict=True)
        model = model.to(dev).eval()
        for i in range(3):
            sliding_window_predict(model, imgs[i % 2], 224, stride, return_device=True)
        torch.cuda.synchronize()
        n = 8
        while True:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                sliding_window_predict(model, imgs[i % 2], 224, stride, return_device=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if ms > 1500:
                break
            n = int(n * max(1.5, 1700 / ms))
        print(f"{W}x{H} s{stride}  window_chunk {chunk:3d}: {ms / n:7.3f} ms per image  {n / ms * 1e3:6.2f} images/s", flush=True)
        del model
        torch.cuda.empty_cache()
