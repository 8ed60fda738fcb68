"""One configuration of the attention kernel for ncu: python profiles/attn_one.py [B] [impl] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clip_ebc_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
t_live, n_const = 197, 32
qkv = torch.randn(B * t_live, 2304, device="cuda").to(torch.bfloat16)
ckv = torch.randn(n_const, 2304, device="cuda").to(torch.bfloat16)
ops.set_attention_impl(impl)
for _ in range(reps):
    out = ops.attention(qkv, B, t_live, ckv)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
