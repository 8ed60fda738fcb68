"""Time the LN-folded GEMM pair in isolation (CUDA events, 50 reps, L2-warm):
  python profiles/gemm_ln_bench.py M         -> out_proj / c_proj with and without the statistics epilogue, QKV / c_fc with and
                                               without the LayerNorm epilogue
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clip_ebc_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 12608
dt = torch.float16


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for K in (768, 3072):
    a = torch.randn(M, K, device="cuda").to(dt)
    w = (torch.randn(768, K, device="cuda") * 0.02).to(dt)
    b = torch.randn(768, device="cuda")
    x = torch.randn(M, 768, device="cuda")
    for bn in (192, 256):
        t0 = timeit(lambda: ops.gemm(a, w, ops.EPI_BIAS_RESID_F32, bias=b, resid=x, out=x, block_n=bn))
        print(f"resid K={K} bn={bn}: plain {t0:.1f} us")
    t1 = timeit(lambda: ops.gemm_resid_stats(a, w, x, b))
    print(f"resid K={K} bn=192 +stats (incl. ~6 us of wrapper allocations): {t1:.1f} us")
x = torch.randn(M, 768, device="cuda")
x16, stats = ops.rowstats(x, fp16=True)
g, be = torch.ones(768, device="cuda"), torch.zeros(768, device="cuda")
for N, gelu in ((2304, False), (3072, True)):
    W = torch.randn(N, 768, device="cuda") * 0.03
    b = torch.randn(N, device="cuda")
    wf, cs, bf = ops.fold_ln_linear(W, b, g, be, fp16=True)
    t0 = timeit(lambda: ops.gemm(x16, wf, ops.EPI_BIAS_GELU_BF16 if gelu else ops.EPI_BIAS_BF16, bias=b))
    t1 = timeit(lambda: ops.gemm_ln(x16, wf, bf, stats, cs, 1, gelu=gelu))
    print(f"N={N} gelu={gelu}: plain {t0:.1f} us, LN epilogue {t1:.1f} us")
