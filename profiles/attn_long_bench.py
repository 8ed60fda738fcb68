"""Micro-benchmark of the attention kernels at the ViT-L/14 shapes (16 heads; deep: 257 live + 32 constant keys; shallow: 289
live): python profiles/attn_long_bench.py [windows]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clip_ebc_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
heads = 16
for t_live, n_const in [(257, 32), (289, 0)]:
    qkv = torch.randn(B * t_live, 3 * 64 * heads, device="cuda").to(torch.bfloat16)
    ckv = torch.randn(n_const, 3 * 64 * heads, device="cuda").to(torch.bfloat16) if n_const else None
    for _ in range(3):
        ops.attention(qkv, B, t_live, ckv, heads=heads)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.attention(qkv, B, t_live, ckv, heads=heads)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    fl = 4.0 * B * heads * t_live * (t_live + n_const) * 64
    print(f"T={t_live}+{n_const}: {ms * 1e3:7.1f} us  {fl / ms / 1e9:6.0f} TF/s")
