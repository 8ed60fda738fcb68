"""Per-kernel breakdown of one backbone through model(x) (library profiling hooks: CUDA events around every launch):
python profiles/prof_backbone.py clip_vit_l_14 [windows]"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clip_ebc_b200 import _lib, get_model  # noqa: E402
from oracle import weights  # noqa: E402

backbone = sys.argv[1] if len(sys.argv) > 1 else "clip_vit_l_14"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 96
patch = {"clip_vit_b_32": 32, "clip_vit_b_16": 16, "clip_vit_l_14": 14}.get(backbone, 0)
dev = torch.device("cuda", 0)
reduction, bins, anchors = weights.bins_and_anchors("r8_t4_nwpu")
lib = _lib.load()
if patch:
    sd = weights.make_state_dict(0, input_size=224, num_vpt=32, deep_vpt=True, variant="default", patch=patch)
    tf = weights.make_text_features(len(bins), seed=100, embed=768 if patch == 14 else 512)
else:
    sd = weights.make_resnet_state_dict(0, backbone[5:], "stress")
    tf = weights.make_text_features(len(bins), seed=100, embed=weights.RESNETS[backbone[5:]]["embed"])
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
model.use_cuda_graphs = False
lib.clipebc_profile_enable(1)
for i in range(3):
    model(xs[i % 2])
buf = ctypes.create_string_buffer(1 << 16)
lib.clipebc_profile_dump(buf, len(buf))
lib.clipebc_profile_enable(0)
prof = json.loads(buf.value.decode())
tot = sum(v["ms"] for v in prof.values())
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    tf_s = f"{v['flops'] / v['ms'] / 1e9:7.0f} TF/s" if v["flops"] else ""
    print(f"    {k:28s} {v['ms'] / 3:8.3f} ms {100 * v['ms'] / tot:5.1f}%  x{v['launches'] / 3:5.1f} {tf_s}")
