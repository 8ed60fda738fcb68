"""CPU: the measurement contract of bench.py that can be checked without a GPU -- the reference arm prints exactly ONE
JSON line on stdout with the keys the driver reads, for the same metric / unit / workload as the GPU arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "windows_per_sec" and d["unit"] == "windows/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("configs[1]:")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_gpu_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: the GPU arm asserts on a CUDA device instead of measuring something else."""
    import torch

    if torch.cuda.is_available():
        return
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"], cwd=ROOT,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert res.returncode != 0 and "CUDA device" in res.stderr
    assert res.stdout.strip() == ""
