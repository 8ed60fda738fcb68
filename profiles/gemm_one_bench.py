"""Time one GEMM shape/impl/tile (CUDA events, 50 reps): python profiles/gemm_one_bench.py M N K epi block_n"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clip_ebc_b200 import ops  # noqa: E402

m, n, k, epi, bn = [int(v) for v in sys.argv[1:6]]
a = torch.randn(m, k, device="cuda").to(torch.float16)
w = (torch.randn(n, k, device="cuda") * 0.03).to(torch.float16)
bias = torch.randn(n, device="cuda")
resid = torch.randn(m, n, device="cuda") if epi in (ops.EPI_BIAS_RESID_F32,) else None
o = ops.gemm(a, w, epi, bias=bias, resid=resid, out=resid, block_n=bn)
for _ in range(5):
    ops.gemm(a, w, epi, bias=bias, resid=resid, out=o, block_n=bn)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    ops.gemm(a, w, epi, bias=bias, resid=resid, out=o, block_n=bn)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print(f"M={m} N={n} K={k} epi={epi} bn={bn}: {ms * 1e3:.1f} us  {2.0 * m * n * k / ms / 1e9:.0f} TF/s")
