"""Pretty-print a bench.py JSON line (headline + per-kernel breakdown)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.4f}  e2e {d['e2e']['value']:.1f}  launches {d['gpu_launches']}")
print("clocks", d["clocks"])
r = d["roofline"]
print(f"roofline: {r['achieved']:.1f} / {r['peak']:.1f} {r['unit']} = {r['frac']:.3f}  share_of_step {r['share_of_step']:.3f}")
if d.get("cpu_baseline"):
    print("cpu_baseline", d["cpu_baseline"])
tot = sum(v["ms_per_step"] for v in d["kernels"].values())
for k, v in d["kernels"].items():
    tf = f"{v['tflops']:.0f} TF/s" if v["tflops"] else ""
    gb = f"{v['gbs']:.0f} GB/s" if v["gbs"] else ""
    print(f"  {k:20s} {v['ms_per_step']:.4f} ms {100 * v['ms_per_step'] / tot:5.1f}%  x{v['launches_per_step']:.0f}  {tf:>10s} {gb:>10s}")
print(f"  {'sum':20s} {tot:.4f} ms")
