"""One GEMM shape, one implementation, a few launches (ncu target): python profiles/gemm_one.py impl M N K epi block_n"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clip_ebc_b200 import ops  # noqa: E402

impl, m, n, k, epi, bn = [int(v) for v in sys.argv[1:7]]
ops.set_gemm_impl(impl)
a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
w = (torch.randn(n, k, device="cuda") * 0.03).to(torch.bfloat16)
bias = torch.randn(n, device="cuda")
resid = torch.randn(m, n, device="cuda") if epi in (ops.EPI_BIAS_RESID_F32,) else None
for _ in range(3):
    o = ops.gemm(a, w, epi, bias=bias, resid=resid, out=resid, block_n=bn)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
