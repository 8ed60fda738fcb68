#!/usr/bin/env python
"""bench.py -- the measurement contract for the CLIP-EBC hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload windows64|sliding]

A "step" is one pass of the hot path over one batch of synthetic input:
  windows64 (default, BASELINE.json configs[1]): model(x), x = 64 synthetic 224x224 windows, deep VPT 32, reduction 8.
  sliding  (BASELINE.json configs[2]): sliding_window_predict on one synthetic 2048x1536 image, window 224, stride 112
           (234 windows), images sharded over the ranks, one all-gather of the per-image counts at the end.
Rank 0 prints ONE JSON line. `value` is whole-job windows/s with inputs resident in HBM; `e2e` is the same metric
through the public API with pinned HOST buffers (H2D of every step's input and D2H of its result inside the timed
region). `--impl reference` times the CPU restatement of the reference (oracle port; the reference is pure Python and
/root/reference does not exist on the GPU box) on the host cores for the same metric.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WINDOW = 224
BATCH = 64
# algorithmic FLOPs per 224x224 window actually executed by this implementation (2*MAC), see DESIGN.md section 5:
# deep VPT runs Q/out/MLP on 197 live rows, K/V on 197 rows (+32 constant prompt rows precomputed at pack time).
FLOPS_PER_WINDOW_NOMINAL_R8 = 58.33e9


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="windows64", choices=["windows64", "sliding"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(source="measured", bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"])
    return dict(source="fallback", bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.first = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_first_sample(self, timeout_s: float = 10.0):
        """nvidia-smi takes a few hundred ms to initialise NVML, and while it does kernel launches of this process stall
        (measured: a timed region that overlaps the start-up runs up to 2x slower). Start it before the warm-up and
        enter the timed region only once it is polling; samples taken before mark() are dropped."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.02)

    def mark(self):
        self.first = len(self.lines)

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines[self.first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def build_model(device):
    from clip_ebc_b200 import get_model
    from oracle import weights

    reduction, bins, anchors = weights.bins_and_anchors("r8_t4_nwpu")
    sd = weights.make_state_dict(0, input_size=224, num_vpt=32, deep_vpt=True, variant="default")
    tf = weights.make_text_features(len(bins), seed=100)
    model = get_model("clip_vit_b_16", input_size=224, reduction=reduction, bins=bins, anchor_points=anchors,
                      prompt_type="word", num_vpt=32, vpt_drop=0.0, deep_vpt=True, text_features=tf)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval(), (sd, tf, anchors, reduction)


def cpu_port_windows_per_sec(sample_windows: int, repeats: int, parts):
    """The oracle (CPU restatement of the reference, fp32) on the host cores: windows/s on a bounded sample."""
    from oracle import clip_ebc_oracle as O
    from oracle import weights

    sd, tf, anchors, reduction = parts
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    x = weights.make_image((sample_windows, 3, WINDOW, WINDOW), seed=7)
    O.clip_ebc_forward(x, sd, tf, anchors, reduction, 32, True)  # warm-up
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.clip_ebc_forward(x, sd, tf, anchors, reduction, 32, True)
        times.append(time.perf_counter() - t0)
    return sample_windows / statistics.median(times), cores, times


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port), all host threads."""
    if rank != 0:
        return
    from oracle import weights

    reduction, bins, anchors = weights.bins_and_anchors("r8_t4_nwpu")
    sd = weights.make_state_dict(0)
    tf = weights.make_text_features(len(bins), seed=100)
    from oracle import clip_ebc_oracle as O

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    sample = 8  # windows per step: a bounded sample of the 64-window batch (the rows of the batch are independent)
    x = weights.make_image((sample, 3, WINDOW, WINDOW), seed=7)
    for _ in range(max(1, min(args.warmup, 2))):
        O.clip_ebc_forward(x, sd, tf, anchors, reduction, 32, True)
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        O.clip_ebc_forward(x, sd, tf, anchors, reduction, 32, True)
    dt = time.perf_counter() - t0
    wps = sample * steps / dt
    line = {
        "impl": "reference", "metric": "windows_per_sec", "value": wps, "unit": "windows/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": 1000 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same workload as the GPU arm (BASELINE.json configs[1]); the CPU arm times a bounded sample of it per step
        "config": {"workload": "configs[1]: ViT-B/16 deep-VPT(32) forward + decoder + EBC head, batch 64 synthetic 224x224 "
                               "windows, reduction 8 (5 bins)",
                   "arm": f"reference algorithm on the host cores (CPU fp32 port, oracle/clip_ebc_oracle.py); each step = "
                          f"{sample} windows sampled from the 64-window batch",
                   "weights": "seeded random init with the reference's init distributions (oracle/weights.py)"},
        "cpu_baseline": {"value": wps, "unit": "windows/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {sample} windows (oracle/clip_ebc_oracle.py, torch {torch.__version__} fp32)"},
        "e2e": {"value": wps, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly ONE JSON line. Libraries write there too (NCCL prints its version banner to stdout at
    NCCL_DEBUG=VERSION, which the GPU boxes set), so file descriptor 1 is pointed at stderr for the life of the process
    and the JSON line goes to a private duplicate of the original stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    args = parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from clip_ebc_b200 import _lib, sliding_window_predict
    from clip_ebc_b200.dist import gather_counts, shard_indices

    assert torch.cuda.is_available(), "bench.py (impl ours) needs a CUDA device: there is no CPU fallback"
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    model, parts = build_model(dev)
    from oracle import weights

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, args.warmup
    peaks = measured_peaks()
    sampler = ClockSampler(local_rank)

    if args.workload == "windows64":
        # ring of distinct input batches larger than L2 (4 x 38.5 MB = 154 MB > 126 MB); weights + activations of one
        # step (~170 MB bf16 weights, ~250 MB activations) also exceed L2, so no step sees a warm cache.
        ring = [weights.make_image((BATCH, 3, WINDOW, WINDOW), seed=50 + i).to(dev) for i in range(4)]
        units_per_step = BATCH
        step = lambda i: model(ring[i % len(ring)])  # noqa: E731
        workload = ("configs[1]: ViT-B/16 deep-VPT(32) forward + decoder + EBC head, batch 64 synthetic 224x224 windows, "
                    "reduction 8 (5 bins), 16-bit tensor-core GEMMs (fp16 operands) / fp32 accumulate+residual, "
                    "1 process per GPU")
        l2 = "ring of 4 distinct input batches (154 MB > 126 MB L2); per-step weights+activations > L2"
    else:
        H, Wd = 1536, 2048
        n_local = 2
        ring = [weights.make_image((1, 3, H, Wd), seed=60 + i + 10 * rank).to(dev) for i in range(n_local)]
        units_per_step = 13 * 18
        step = lambda i: sliding_window_predict(model, ring[i % n_local], WINDOW, 112, return_device=True,  # noqa: E731
                                                return_count=True)
        workload = ("configs[2]: sliding_window_predict on synthetic 2048x1536 images, window 224, stride 112 "
                    "(234 windows/image), images sharded round-robin over ranks, one count all-gather at the end")
        l2 = "each image touches > 1 GB of activations (>> 126 MB L2)"

    sampler.start()
    for i in range(W):
        last = step(i)
    if args.workload == "sliding":
        # the collective is part of the warm-up too: its first call creates the NCCL communicator and loads torch's
        # fill / index kernels (lazy module loading) -- 70-200 ms that belong to no step
        gather_counts(last[1].reshape(1), world, rank, world)
    barrier()
    sampler.wait_first_sample()

    # ---- timed region (device-resident inputs): CUDA events on the launching stream, max over ranks ---------------
    sampler.mark()
    l0 = lib.clipebc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    last = None
    # bounded run-ahead: at most 2 steps in flight, so the host never sits on a full launch queue (with ~260 launches
    # per image the queue fills after 4 images; a blocked launch thread next to the nvidia-smi sampler made the timed
    # region vary by 2x between runs); the GPU still always has the next step queued
    inflight, step_events = [], []
    for i in range(K):
        if len(inflight) == 2:
            inflight.pop(0).synchronize()
        last = step(i)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        inflight.append(ev)
        step_events.append(ev)
    if args.workload == "sliding":
        # the path's only collective: all-gather of the per-image counts (here: of the last image of every rank)
        counts = gather_counts(last[1].reshape(1), world, rank, world)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if os.environ.get("BENCH_DEBUG"):
        prev, per = e0, []
        for ev in step_events:
            per.append(prev.elapsed_time(ev))
            prev = ev
        print("[bench] per-step ms: " + " ".join(f"{v:.1f}" for v in per), file=sys.stderr)
    launches = lib.clipebc_launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()
    value = world * units_per_step * K / (ms_max / 1000.0)

    # ---- end-to-end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region ----------
    # two-deep pipeline: the H2D copy of step i+1 (copy stream) overlaps the kernels of step i (compute stream).
    e2e = None
    if args.workload == "windows64":
        host_in = [weights.make_image((BATCH, 3, WINDOW, WINDOW), seed=50 + i).pin_memory() for i in range(4)]
        dev_in = [torch.empty((BATCH, 3, WINDOW, WINDOW), device=dev) for _ in range(2)]
        host_out = [torch.empty((BATCH, 1, 28, 28)).pin_memory() for _ in range(2)]
        copy_s, comp_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def e2e_loop(n):
            for i in range(n):
                b = i % 2
                with torch.cuda.stream(copy_s):
                    if i >= 2:
                        copy_s.wait_event(consumed[b])
                    dev_in[b].copy_(host_in[i % 4], non_blocking=True)
                    copied[b].record(copy_s)
                with torch.cuda.stream(comp_s):
                    comp_s.wait_event(copied[b])
                    out = model(dev_in[b])
                    consumed[b].record(comp_s)
                    host_out[b].copy_(out, non_blocking=True)

        e2e_loop(W)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        s0.record(copy_s)
        comp_s.wait_event(s0)
        e2e_loop(K)
        copy_s.wait_stream(comp_s)
        s1.record(copy_s)
        barrier()
        t2 = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": world * BATCH * K / (t2.item() / 1000.0), "unit": "windows/s",
               "h2d_bytes_per_step": BATCH * 3 * WINDOW * WINDOW * 4, "d2h_bytes_per_step": BATCH * 28 * 28 * 4,
               "ms_per_step": t2.item() / K, "note": "pinned host buffers, 2-deep copy/compute pipeline"}
    else:
        host_img = [weights.make_image((1, 3, 1536, 2048), seed=60 + i + 10 * rank).pin_memory() for i in range(2)]
        for i in range(2):
            sliding_window_predict(model, host_img[i % 2], WINDOW, 112)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        s0.record()
        for i in range(K):
            sliding_window_predict(model, host_img[i % 2], WINDOW, 112)  # CPU image in, CPU density out (reference API)
        s1.record()
        barrier()
        t2 = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": world * units_per_step * K / (t2.item() / 1000.0), "unit": "windows/s",
               "h2d_bytes_per_step": 3 * 1536 * 2048 * 4, "d2h_bytes_per_step": 192 * 256 * 4,
               "ms_per_step": t2.item() / K, "images_per_sec": world * K / (t2.item() / 1000.0)}

    clocks = sampler.stop()  # sampled over both timed regions (device-resident and end-to-end)

    # ---- per-kernel breakdown (separate instrumented pass: CUDA events around every launch on its stream) ---------
    lib.clipebc_profile_enable(1)
    prof_steps = 3
    for i in range(prof_steps):
        step(i)
    buf = ctypes.create_string_buffer(1 << 16)
    _lib.check(lib.clipebc_profile_dump(buf, len(buf)), "profile_dump")
    lib.clipebc_profile_enable(0)
    prof = json.loads(buf.value.decode())
    gemm_ms = sum(v["ms"] for k, v in prof.items() if k.startswith("gemm"))
    gemm_flops = sum(v["flops"] for k, v in prof.items() if k.startswith("gemm"))
    gemm_launches = sum(v["launches"] for k, v in prof.items() if k.startswith("gemm"))
    total_ms = sum(v["ms"] for v in prof.values())
    achieved = gemm_flops / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
    peak = peaks["bf16_sustained"]
    # DRAM traffic per launch of the same kernel family from the committed ncu capture (dram__bytes_read + write per
    # launch, weighted by how often each GEMM of the step is launched); cold-L2 figures, so an upper bound in-step.
    traffic, traffic_note = None, "no ncu capture committed"
    tpath = os.path.join(ROOT, "profiles", "r01k_gemm_dram.json")
    if args.workload == "windows64" and os.path.exists(tpath):
        per_tag = json.load(open(tpath))["per_tag"]
        num = den = 0.0
        for k, v in prof.items():
            tag = k.split(":", 1)[1] if k.startswith("gemm:") else None
            if tag in per_tag:
                n = v["launches"] / prof_steps
                num += n * (per_tag[tag]["dram_read_bytes_per_launch"] + per_tag[tag]["dram_write_bytes_per_launch"])
                den += n
        if den > 0:
            traffic = num / den
            traffic_note = ("bytes per launch, launch-weighted mean over the GEMMs of a step, from profiles/r01k_gemm_dram.json "
                            "(ncu dram__bytes_read.sum + dram__bytes_write.sum, cold L2 per launch)")
    roofline = {
        "bound": "tensor", "kernel": "gemm2_tcgen05_kernel (all epilogues)", "achieved": achieved, "peak": peak,
        "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
        "algorithmic_bytes_per_launch": (sum(v["bytes"] for k, v in prof.items() if k.startswith("gemm")) / max(1, gemm_launches)),
        "peak_source": f"bf16_tflops_sustained of {peaks['source']} (kernel timed inside a long step)",
        "avg_launch_ms": gemm_ms / max(1, gemm_launches), "gflop_per_launch": gemm_flops / max(1, gemm_launches) / 1e9,
        "share_of_step": gemm_ms / total_ms if total_ms else None,
    }
    kernels = {k: {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["launches"] / prof_steps,
                   "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["ms"] > 0 and v["flops"] > 0 else None,
                   "gbs": (v["bytes"] / (v["ms"] / 1e3) / 1e9) if v["ms"] > 0 and v["bytes"] > 0 else None}
               for k, v in prof.items()}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        wps, cores, times = cpu_port_windows_per_sec(8, 3, parts)
        cpu_baseline = {"value": wps, "unit": "windows/s", "cores": cores, "kind": "port",
                        "sample": f"8 of the 64 windows of one step, median of 3 runs after 1 warm-up "
                                  f"({statistics.median(times):.2f} s per run), oracle/clip_ebc_oracle.py fp32"}

    if rank == 0:
        line = {
            "metric": "windows_per_sec", "value": value, "unit": "windows/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16",
            "dtype_note": "16-bit tensor-core operands on tcgen05 kind::f16: fp16 by default (11-bit mantissa; bf16 selectable, same "
                          "rate); fp32 accumulate / residual stream / LayerNorm statistics / softmax / head",
            "data": "synthetic",
            "config": {"workload": workload, "l2": l2, "global_batch_windows": world * units_per_step,
                       "parallelism": f"dp{world} (independent windows/images per rank)",
                       "weights": "seeded random init with the reference's init distributions (oracle/weights.py)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "images_per_sec": (world * K / (ms_max / 1000.0)) if args.workload == "sliding" else None,
            "tensor_frac_whole_step_nominal": (value / world) * FLOPS_PER_WINDOW_NOMINAL_R8 / 1e12 / peak
            if args.workload == "windows64" else None,
            "kernels": kernels,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
