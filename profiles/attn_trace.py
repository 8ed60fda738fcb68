"""Time line of CTA 0 of the persistent attention kernel (CLIPEBC_ATTN_TRACE=1): python profiles/attn_trace.py"""
import os
import sys

os.environ["CLIPEBC_ATTN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clip_ebc_b200 import ops  # noqa: E402

B, t_live, n_const = 64, 197, 32
qkv = torch.randn(B * t_live, 2304, device="cuda").to(torch.bfloat16)
ckv = torch.randn(n_const, 2304, device="cuda").to(torch.bfloat16)
ops.set_attention_impl(int(sys.argv[1]) if len(sys.argv) > 1 else 4)  # 3 and 4 carry the trace hooks
for _ in range(3):
    ops.attention(qkv, B, t_live, ckv)
torch.cuda.synchronize()
