"""N-rank correctness on hardware: image-sharded sliding_window_predict + the single NCCL all-gather
of per-image counts gives, on every rank, exactly the bits the 1-GPU path gives. Spawns `torch.distributed.run` with one
process per GPU; skipped when the box has fewer GPUs than ranks (the worker is tests/dist_worker.py)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_counts_equal_single_gpu_counts_bit_for_bit(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    out = tmp_path / "dist.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py"), str(out)]
    res = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:]
    rec = json.load(open(out))
    print(f"\n[{world} ranks] counts {rec['counts'][:4]} ... bit-exact on every rank: {rec['bit_exact_on_every_rank']}")
    assert rec["world"] == world and rec["finite"] and rec["bit_exact_on_every_rank"]
    assert rec["counts"] == rec["counts_1gpu"]
    keep = os.environ.get("CLIPEBC_DIST_RECORD_DIR")  # gpurun sessions keep the record under profiles/
    if keep:
        os.makedirs(keep, exist_ok=True)
        json.dump(rec, open(os.path.join(keep, f"dist_bit_exact_{world}gpu.json"), "w"), indent=1)
