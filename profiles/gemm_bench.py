"""Micro-benchmark of the tcgen05 GEMM on the shapes of the hot path (B windows), per tile width.

  python profiles/gemm_bench.py [B]
CUDA events around `reps` back-to-back launches after warm-up; activations at B=64 are L2-resident between launches,
exactly as inside a forward step (the previous kernel just wrote them).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clip_ebc_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
M = B * 197
Mp = B * 900
dev = "cuda"
shapes = [
    ("qkv", M, 2304, 768, ops.EPI_BIAS_BF16),
    ("out_proj", M, 768, 768, ops.EPI_BIAS_RESID_F32),
    ("c_fc", M, 3072, 768, ops.EPI_BIAS_GELU_BF16),
    ("c_proj", M, 768, 3072, ops.EPI_BIAS_RESID_F32),
    ("projection", Mp, 512, 2304, ops.EPI_BIAS_F32),
    ("conv_like", Mp, 768, 6912, ops.EPI_BIAS_BF16),
]
for name, m, n, k, epi in shapes:
    a = torch.randn(m, k, device=dev).to(torch.bfloat16)
    w = (torch.randn(n, k, device=dev) * 0.03).to(torch.bfloat16)
    bias = torch.randn(n, device=dev)
    resid = torch.randn(m, n, device=dev) if epi == ops.EPI_BIAS_RESID_F32 else None
    out = resid if resid is not None else None
    res = []
    for bn in (128, 192, 256, 0):
        if bn and n % bn:
            continue
        o = ops.gemm(a, w, epi, bias=bias, resid=resid, out=out, block_n=bn)
        for _ in range(3):
            ops.gemm(a, w, epi, bias=bias, resid=resid, out=o, block_n=bn)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            ops.gemm(a, w, epi, bias=bias, resid=resid, out=o, block_n=bn)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res.append(f"bn={bn or 'auto':>4}: {ms * 1e3:6.1f} us {2.0 * m * n * k / ms / 1e9:5.0f} TF/s")
    print(f"{name:11s} M={m} N={n} K={k}\n   " + "\n   ".join(res))
