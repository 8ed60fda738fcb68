"""Minimal driver for ncu: build the model, run `steps` passes of the bench workload (windows64: model(x) on 64 windows).

  python profiles/prof_step.py [steps] [batch]
Used for the launch list (ncu --metrics gpu__time_duration.sum) and the `--set full` capture of the top kernel.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import build_model  # noqa: E402
from oracle import weights  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
model, _ = build_model(dev, "r8_deep")
x = weights.make_image((batch, 3, 224, 224), seed=50).to(dev)
model.use_cuda_graphs = False  # ncu must see the individual launches
for _ in range(steps):
    out = model(x)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.sum()))
