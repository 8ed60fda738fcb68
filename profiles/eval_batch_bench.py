"""Cross-image window batching on small images: per-image sliding_window_predict vs sliding_window_predict_batch.

  python profiles/eval_batch_bench.py [H] [W] [stride] [n_images]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import build_model  # noqa: E402
from clip_ebc_b200 import sliding_window_predict, sliding_window_predict_batch  # noqa: E402
from oracle import weights  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 768
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
stride = int(sys.argv[3]) if len(sys.argv) > 3 else 224
n = int(sys.argv[4]) if len(sys.argv) > 4 else 16
dev = torch.device("cuda", 0)
model, _ = build_model(dev, "r8_deep")
images = [weights.make_image((1, 3, H, W), seed=70 + i).to(dev) for i in range(n)]


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def per_image():
    for im in images:
        sliding_window_predict(model, im, 224, stride, return_device=True, return_count=True)


def batched(group):
    def run():
        for i in range(0, n, group):
            sliding_window_predict_batch(model, images[i:i + group], 224, stride, return_device=True)
    return run


from clip_ebc_b200 import ops  # noqa: E402

ro, co = ops.window_origins(H, W, (224, 224), (stride, stride))
nw = len(ro) * len(co)
t1 = timed(per_image)
print(f"{n} images {H}x{W}, stride {stride}: {nw} windows each")
print(f"  per image      : {t1:8.2f} ms  {n / t1 * 1e3:8.1f} images/s  {n * nw / t1 * 1e3:8.0f} windows/s")
for group in (2, 4, 8, 16):
    if group <= n:
        t = timed(batched(group))
        print(f"  batches of {group:2d}  : {t:8.2f} ms  {n / t * 1e3:8.1f} images/s  {n * nw / t * 1e3:8.0f} windows/s")
