"""Does a CUDA graph of model(x) pay for small batches? eager vs torch.cuda.graph replay (python profiles/graph_probe.py)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import build_model  # noqa: E402
from oracle import weights  # noqa: E402

dev = torch.device("cuda", 0)
model, _ = build_model(dev, "r8_deep")
for B in (1, 4, 16, 64):
    x = weights.make_image((B, 3, 224, 224), seed=50).to(dev)
    for _ in range(5):
        ref = model(x)
    torch.cuda.synchronize()

    def timeit(fn, n=50):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    t_eager = timeit(lambda: model(x))
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                model(x)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            out = model(x)
        g.replay()
        torch.cuda.synchronize()
        same = torch.equal(out, ref)
        t_graph = timeit(g.replay)
        print(f"B={B}: eager {t_eager:.3f} ms, graph {t_graph:.3f} ms, identical={same}")
    except Exception as e:  # noqa: BLE001
        print(f"B={B}: eager {t_eager:.3f} ms, graph capture failed: {type(e).__name__}: {str(e)[:200]}")
        torch.cuda.synchronize()
