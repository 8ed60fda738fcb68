"""ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum of every GEMM launch of ONE windows64
step) -> profiles/rNN_gemm_dram.json, the per-use DRAM traffic bench.py reports as `roofline.traffic`.

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:gemm2_tcgen05 --csv --log-file gpurun_out/gemm_dram.csv python profiles/prof_step.py 1
  python profiles/make_gemm_dram.py gpurun_out/gemm_dram.csv profiles/r01i_gemm_dram.json

Launch order of a step (api.cu run_windows), kernel = gemm2_tcgen05_kernel<BLOCK_N, EPI>: patch_embed <256,0>, 12 x [qkv
<256,2>, out_proj <192,4>, c_fc <256,3>, c_proj <192,4>], dec_conv1 (coarse-grid form: <256,2> with N = 6912, the 13th
<*,2> launch of a step; fine-grid form: <256,5>), dec_conv2 <256,8> (upsampled-skip epilogue; fine-grid form <256,6>),
projection+head <256,7>. Pack-time launches (the constant prompt K/V rows: <128,2> with M = 32) are skipped.
"""
import collections
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if not l.startswith("==")]
launches = collections.OrderedDict()
for row in csv.DictReader(lines):
    lid = row["ID"]
    d = launches.setdefault(lid, {"kernel": re.sub(r"\(.*", "", row["Kernel Name"])})
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    name = row["Metric Name"]
    if name.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    elif name == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
    d[name] = v
per_tag = collections.OrderedDict()
n_resid = 0
n_epi2 = 0
for d in launches.values():
    m = re.search(r"gemm2_tcgen05_kernel<(\d+), (\d+)>", d["kernel"])
    if not m:
        continue
    bn, epi = int(m.group(1)), int(m.group(2))
    if epi == 2 and bn == 128:
        continue  # pack time: constant prompt K/V rows
    if epi == 4:
        tag = "out_proj" if n_resid % 2 == 0 else "c_proj"
        n_resid += 1
    elif epi == 2:
        tag = "dec_conv1" if n_epi2 % 13 == 12 else "qkv"
        n_epi2 += 1
    else:
        tag = {0: "patch_embed", 3: "c_fc", 5: "dec_conv1", 6: "dec_conv2", 8: "dec_conv2", 7: "projection+head"}.get(epi)
    if tag is None:
        continue
    t = per_tag.setdefault(tag, {"kernel": f"gemm2_tcgen05_kernel<{bn}, {epi}>", "captured_launches": 0, "r": 0.0, "w": 0.0, "t": 0.0})
    t["captured_launches"] += 1
    t["r"] += d.get("dram__bytes_read.sum", 0.0)
    t["w"] += d.get("dram__bytes_write.sum", 0.0)
    t["t"] += d.get("gpu__time_duration.sum", 0.0)
out = {"source": f"{src}: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                 "(serialised launches, cold L2 per launch)", "per_tag": {}}
for tag, t in per_tag.items():
    n = t["captured_launches"]
    out["per_tag"][tag] = {"kernel": t["kernel"], "captured_launches": n, "dram_read_bytes_per_launch": t["r"] / n,
                           "dram_write_bytes_per_launch": t["w"] / n, "ncu_time_us_per_launch": t["t"] / n}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps({k: (v["captured_launches"], round(v["dram_read_bytes_per_launch"] / 1e6, 1), round(v["dram_write_bytes_per_launch"] / 1e6, 1))
                  for k, v in out["per_tag"].items()}))
