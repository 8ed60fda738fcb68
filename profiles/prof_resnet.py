"""Minimal driver for ncu: CLIP-ResNet-50 model(x) on 96 windows, `passes` forwards without CUDA graphs.
  python profiles/prof_resnet.py [passes] [backbone]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clip_ebc_b200 import get_model  # noqa: E402
from oracle import weights  # noqa: E402

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
backbone = sys.argv[2] if len(sys.argv) > 2 else "resnet50"
dev = torch.device("cuda", 0)
reduction, bins, anchors = weights.bins_and_anchors("r8_t4_nwpu")
sd = weights.make_resnet_state_dict(0, backbone, "stress")
tf = weights.make_text_features(len(bins), seed=100, embed=weights.RESNETS[backbone]["embed"])
model = get_model("clip_" + backbone, input_size=224, reduction=reduction, bins=bins, anchor_points=anchors, prompt_type="word",
                  num_vpt=32, vpt_drop=0.0, deep_vpt=True, text_features=tf)
model.load_state_dict(sd, strict=True)
model = model.to(dev).eval()
model.use_cuda_graphs = False
x = weights.make_image((96, 3, 224, 224), seed=70).to(dev)
for _ in range(passes):
    out = model(x)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.sum()))
