"""Where does the device-resident sliding loop lose time? Host enqueue time of every sliding_window_predict call vs the
GPU time of the whole loop (python profiles/sliding_host_probe.py [steps])."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import build_model  # noqa: E402
from clip_ebc_b200 import sliding_window_predict  # noqa: E402
from oracle import weights  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda", 0)
model, _ = build_model(dev, "r8_deep")
imgs = [weights.make_image((1, 3, 1536, 2048), seed=60 + i).to(dev) for i in range(2)]
for i in range(5):
    sliding_window_predict(model, imgs[i % 2], 224, 112, return_device=True, return_count=True)
torch.cuda.synchronize()
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = []
    t_all = time.perf_counter()
    e0.record()
    for i in range(K):
        t0 = time.perf_counter()
        last = sliding_window_predict(model, imgs[i % 2], 224, 112, return_device=True, return_count=True)
        host.append((time.perf_counter() - t0) * 1e3)
    e1.record()
    t_enq = (time.perf_counter() - t_all) * 1e3
    torch.cuda.synchronize()
    gpu = e0.elapsed_time(e1)
    host_sorted = sorted(host)
    print(f"rep {rep}: gpu {gpu / K:.2f} ms/step, host enqueue total {t_enq:.1f} ms ({t_enq / K:.2f}/step), "
          f"per-call median {host_sorted[K // 2]:.2f} max {host_sorted[-1]:.2f} ms; first 8: "
          + " ".join(f"{h:.1f}" for h in host[:8]))
