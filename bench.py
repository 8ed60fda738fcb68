#!/usr/bin/env python
"""bench.py -- the measurement contract for the CLIP-EBC hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|torch_gpu] [--workload all|<name>]

A "step" is one pass of the hot path over one batch of synthetic input, per rank (weak scaling: one process per GPU,
independent windows / images per rank). One invocation measures every workload of BASELINE.json and prints ONE JSON line
on rank 0; the top-level `value` / `e2e` / `roofline` are those of the headline workload, `workloads` holds all of them:

  windows64  configs[1] (headline): model(x), x = 64 synthetic 224x224 windows, ViT-B/16 deep VPT 32, reduction 8
  sliding    configs[2]: sliding_window_predict on synthetic 2048x1536 images, window 224, stride 112 (234 windows / image),
             images sharded round-robin over the ranks (clip_ebc_b200.dist.shard_indices), ONE all-gather of the per-image
             counts INSIDE the timed region -> windows/s and images/s
  r16, r32   configs[3]: 256 windows, shallow VPT, reduction 16 (9 bins) / 32 (20 bins)
  qnrf224, qnrf112  configs[4]: 4096x3072 images, stride 224 (266 windows) / 112 (972 windows), sharded + count all-gather

Timing rules: W >= 3 warm-up steps; every timed region is stretched to >= --min-seconds (default 2 s) by repeating the K
driver steps `reps` times (`steps` is reported as given, `timed_passes` = K * reps, `ms_per_step` = region / timed_passes),
so that clocks are in the sustained regime; CUDA events on the launching stream, barrier + synchronize on both sides, MAX
over ranks; inputs rotate over a ring larger than L2; nvidia-smi clocks are sampled during the timed regions.
`value` is whole-job windows/s with inputs resident in HBM; `e2e` is the same metric through the public API with pinned HOST
buffers (H2D of every step's input and D2H of its result inside the timed region).
`roofline`: the tcgen05 GEMM kernel family (CUDA events around every launch in an instrumented pass that directly follows the
timed region): executed 2*M*N*K / duration against BOTH measured peaks (burst, sustained); `frac` uses the peak of the regime
the region ran in. `flops_executed` counts what the kernels multiply (padded rows, and with bf16 operands the split-precision
segments, included),
`flops_algorithmic` what SURVEY.md section 8(d) credits (never more than what was run).
`--impl reference` times the CPU restatement of the reference (oracle port; the reference is pure Python and /root/reference
does not exist on the GPU box) on the host cores for the headline metric. `--impl torch_gpu` (also embedded in the main line
as `gpu_library_baseline`) runs the same functional PyTorch forward on the GPU (cuBLAS / cuDNN / SDPA: the library bar).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
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
HEADLINE = "windows64"

# name -> description of a workload. kind "windows": model(x) on `batch` windows; kind "sliding": one image per step and rank.
WORKLOADS = {
    "windows64": dict(kind="windows", model="r8_deep", batch=64, config="configs[1]",
                      text="ViT-B/16 deep-VPT(32) forward + decoder + EBC head, batch 64 synthetic 224x224 windows, reduction 8 (5 bins)"),
    "sliding": dict(kind="sliding", model="r8_deep", H=1536, W=2048, stride=112, config="configs[2]",
                    text="sliding_window_predict on synthetic 2048x1536 images, window 224, stride 112 (234 windows/image), images "
                         "sharded round-robin over ranks, one count all-gather inside the timed region"),
    "r16": dict(kind="windows", model="r16_shallow", batch=256, config="configs[3]",
                text="reduction 16 (9 bins), shallow VPT(32), batch 256 synthetic 224x224 windows"),
    "r32": dict(kind="windows", model="r32_shallow", batch=256, config="configs[3]",
                text="reduction 32 (20 bins), shallow VPT(32), batch 256 synthetic 224x224 windows"),
    "qnrf224": dict(kind="sliding", model="r8_deep", H=3072, W=4096, stride=224, config="configs[4]",
                    text="sliding_window_predict on synthetic 4096x3072 images, window 224, stride 224 (266 windows/image), sharded, "
                         "count all-gather"),
    "qnrf112": dict(kind="sliding", model="r8_deep", H=3072, W=4096, stride=112, config="configs[4]",
                    text="sliding_window_predict on synthetic 4096x3072 images, window 224, stride 112 (972 windows/image), sharded, "
                         "count all-gather"),
}
MODELS = {"r8_deep": dict(bins="r8_t4_nwpu", deep=True), "r16_shallow": dict(bins="r16_t8_qnrf", deep=False),
          "r32_shallow": dict(bins="r32_t19_qnrf", deep=False)}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--workload", default="all", choices=["all"] + list(WORKLOADS))
    ap.add_argument("--min-seconds", type=float, default=2.0, help="minimum length of every timed region")
    ap.add_argument("--operand-dtype", default="fp16", choices=["fp16", "bf16"],
                    help="16-bit tensor-core operand format of the GEMMs (same kind::f16 rate); the headline line also carries a "
                         "bf16 measurement of the headline workload (`bf16_operands`)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-library-baseline", action="store_true")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(source="MEASURED_PEAKS.json", bf16_burst=d["bf16_tflops"],
                    bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), hbm=d["hbm_gbs"])
    # B200_PROFILING.md fallback figures
    return dict(source="fallback (B200_PROFILING.md)", bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0)


# ------------------------------------------------------------------------------------------------ FLOP accounting
def flops_per_window(model_key: str, split_precision: bool = False) -> dict:
    """FLOPs (2 * MAC) per 224x224 window of ViT-B/16 at the given configuration, per stage:
    (algorithmic, executed). `algorithmic` is SURVEY's figure, capped by what this implementation runs (deep VPT: the 32
    prompt rows are dead -> 197 live rows; conv1 from the coarse grid); `executed` is what the tensor cores multiply,
    including the border rows of the shared-border decoder grid and -- with bf16 operands only (`split_precision`) -- the 2
    extra segments of the split-precision patch-embed / projection GEMMs."""
    deep = MODELS[model_key]["deep"]
    red = {"r8_deep": 8, "r16_shallow": 16, "r32_shallow": 32}[model_key]
    D, L, Hd, P, E = 768, 12, 3072, 196, 512
    T = 197 if deep else 229            # live rows
    Tk = 229                            # keys per query either way (deep: 197 live + 32 constant prompt keys)
    g = WINDOW // red
    Mp = (g + 1) * (g + 1)              # rows of the shared-border decoder grid
    st = {}
    seg = 3 if split_precision else 1
    st["patch_embed"] = (2.0 * P * D * D, 2.0 * P * D * seg * D)
    gemm_layer = 2.0 * T * D * (3 * D + D + Hd + Hd)
    attn_layer = 4.0 * T * Tk * D
    st["vit_gemms"] = (L * gemm_layer, L * gemm_layer)
    st["attention"] = (L * attn_layer, L * attn_layer)
    conv_alg = 2.0 * g * g * D * 9 * D
    if g >= 28:   # decoder grid >= 2x the patch grid: conv1 from the coarse grid (GEMM on 196 rows x 9 taps + a gather)
        c1 = 2.0 * P * D * 9 * D
        st["dec_conv1"] = (min(conv_alg, c1), c1)
    else:
        st["dec_conv1"] = (conv_alg, 2.0 * Mp * D * 9 * D)
    st["dec_conv2"] = (conv_alg, 2.0 * Mp * D * 9 * D)
    st["projection"] = (2.0 * g * g * D * E, 2.0 * Mp * seg * D * E)
    return st


def flops_totals(model_key: str, split_precision: bool = False) -> dict:
    st = flops_per_window(model_key, split_precision)
    alg = sum(a for a, _ in st.values())
    exe = sum(e for _, e in st.values())
    vit_alg = st["vit_gemms"][0] + st["attention"][0] + st["patch_embed"][0]
    vit_exe = st["vit_gemms"][1] + st["attention"][1] + st["patch_embed"][1]
    return dict(algorithmic=alg, executed=exe, vit_algorithmic=vit_alg, vit_executed=vit_exe,
                nominal_reference=58.33e9 if model_key == "r8_deep" else (45.38e9 if model_key == "r16_shallow" else 42.14e9))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.windows = []   # [start, end) indices into self.lines of the timed regions
        self._open = None

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
        (measured: a timed region that overlaps the start-up runs up to 2x slower). It is started before the first warm-up
        and the first timed region is entered only once it is polling."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.02)

    def region_begin(self):
        self._open = len(self.lines)

    def region_end(self):
        if self._open is not None:
            self.windows.append((self._open, len(self.lines)))
            self._open = None

    def _parse(self, rows):
        sm, mx, pw, reasons = [], [], [], set()
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w=statistics.median(pw) if pw else None, reasons=sorted(reasons), samples=len(sm))

    def summary(self, window=None):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        if window is not None:
            return self._parse(self.lines[window[0]:window[1]])
        rows = []
        for a, b in self.windows:
            rows += self.lines[a:b]
        return self._parse(rows)

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()


def build_model(device, model_key: str, operand_dtype: str = "fp16"):
    from clip_ebc_b200 import get_model
    from oracle import weights

    spec = MODELS[model_key]
    reduction, bins, anchors = weights.bins_and_anchors(spec["bins"])
    sd = weights.make_state_dict(0, input_size=224, num_vpt=32, deep_vpt=spec["deep"], variant="default")
    tf = weights.make_text_features(len(bins), seed=100)
    model = get_model("clip_vit_b_16", input_size=224, reduction=reduction, bins=bins, anchor_points=anchors,
                      prompt_type="word", num_vpt=32, vpt_drop=0.0, deep_vpt=spec["deep"], text_features=tf,
                      operand_dtype=operand_dtype)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval(), (sd, tf, anchors, reduction, spec["deep"])


# ------------------------------------------------------------------------------------------------ CPU / library legs
def cpu_port_windows_per_sec(sample_windows: int, repeats: int, parts):
    """The oracle (CPU restatement of the reference, fp32) on the host cores: windows/s on a bounded sample."""
    from oracle import clip_ebc_oracle as O
    from oracle import weights

    sd, tf, anchors, reduction, deep = parts
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    x = weights.make_image((sample_windows, 3, WINDOW, WINDOW), seed=7)
    for _ in range(2):
        O.clip_ebc_forward(x, sd, tf, anchors, reduction, 32, deep)  # warm-up
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.clip_ebc_forward(x, sd, tf, anchors, reduction, 32, deep)
        times.append(time.perf_counter() - t0)
    return sample_windows / statistics.median(times), cores, times


def gpu_library_baseline(device, parts, batch: int, seconds: float = 1.0):
    """The bar on the same B200: the oracle's functional PyTorch forward -- the
    reference's own operator sequence (229-token cats, F.linear / SDPA / F.conv2d / F.layer_norm: cuBLAS, cuDNN, flash
    SDPA kernels) -- run on the GPU on the same 64-window batch, in fp32 (TF32 off), TF32 and bf16 autocast. It is a
    baseline leg like `cpu_baseline`: the only thing bench.py does with oracle/ besides timing the CPU port."""
    from oracle import clip_ebc_oracle as O
    from oracle import weights

    sd, tf, anchors, reduction, deep = parts
    sdd = {k: v.to(device) for k, v in sd.items()}
    tfd = tf.to(device)
    xs = [weights.make_image((batch, 3, WINDOW, WINDOW), seed=50 + i).to(device) for i in range(4)]
    out = {}
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    for mode in ("fp32", "tf32", "bf16_autocast"):
        tf32 = mode != "fp32"
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32

        def step(i):
            if mode == "bf16_autocast":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return O.clip_ebc_forward(xs[i % 4], sdd, tfd, anchors, reduction, 32, deep)[1]
            return O.clip_ebc_forward(xs[i % 4], sdd, tfd, anchors, reduction, 32, deep)[1]

        for i in range(3):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(0); e1.record(); torch.cuda.synchronize()
        n = max(3, min(200, int(math.ceil(seconds * 1000.0 / max(e0.elapsed_time(e1), 1e-3)))))
        e0.record()
        for i in range(n):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out[mode] = {"windows_per_sec": batch / (ms / 1000.0), "ms_per_step": ms, "steps": n}
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    out["what"] = ("oracle/clip_ebc_oracle.py functional forward (the reference's operator sequence, nominal 229-token cats) on "
                   f"cuda via torch {torch.__version__}: cuBLAS / cuDNN / SDPA kernels, batch {batch} windows, device-resident inputs, "
                   "no CUDA graph, no torch.compile")
    return out


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port), all host threads."""
    if rank != 0:
        return
    from oracle import clip_ebc_oracle as O
    from oracle import weights

    reduction, bins, anchors = weights.bins_and_anchors("r8_t4_nwpu")
    sd = weights.make_state_dict(0)
    tf = weights.make_text_features(len(bins), seed=100)
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    sample = 8  # windows per step: a bounded sample of the 64-window batch (the rows of the batch are independent)
    x = weights.make_image((sample, 3, WINDOW, WINDOW), seed=7)
    warm = max(3, min(args.warmup, 5))
    for _ in range(warm):
        O.clip_ebc_forward(x, sd, tf, anchors, reduction, 32, True)
    steps = max(1, min(args.steps, 40))
    per = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.clip_ebc_forward(x, sd, tf, anchors, reduction, 32, True)
        per.append(time.perf_counter() - t0)
    dt = sum(per)
    wps = sample * steps / dt
    line = {
        "impl": "reference", "metric": "windows_per_sec", "value": wps, "unit": "windows/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1000 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same workload as the GPU arm (BASELINE.json configs[1]); the CPU arm times a bounded sample of it per step
        "config": {"workload": f"{WORKLOADS[HEADLINE]['config']}: {WORKLOADS[HEADLINE]['text']}",
                   "arm": f"reference algorithm on the host cores (CPU fp32 port, oracle/clip_ebc_oracle.py); each step = "
                          f"{sample} windows sampled from the 64-window batch",
                   "weights": "seeded random init with the reference's init distributions (oracle/weights.py)"},
        "cpu_baseline": {"value": wps, "unit": "windows/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {sample} windows after {warm} warm-up steps (oracle/clip_ebc_oracle.py, torch "
                                   f"{torch.__version__} fp32); median step {1000 * statistics.median(per):.0f} ms, min "
                                   f"{1000 * min(per):.0f}, max {1000 * max(per):.0f}"},
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


# ------------------------------------------------------------------------------------------------ the GPU arm
class Runner:
    def __init__(self, args):
        import torch.distributed as dist

        self.args = args
        self.dist = dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        assert torch.cuda.is_available(), "bench.py (impl ours) needs a CUDA device: there is no CPU fallback"
        assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        from clip_ebc_b200 import _lib

        self._lib = _lib
        self.lib = _lib.load()
        self.peaks = measured_peaks()
        self.sampler = ClockSampler(self.local_rank)
        self.models = {}
        self.parts = {}

    def model(self, key, operand_dtype=None):
        od = operand_dtype or self.args.operand_dtype
        if (key, od) not in self.models:
            self.models[(key, od)], self.parts[key] = build_model(self.dev, key, od)
        return self.models[(key, od)]

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        t = torch.tensor([v], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()

    # -- one timed region: `n` calls of step(i) (+ `tail`, e.g. the count all-gather), device-timed, max over ranks ------
    def timed(self, step, n, tail=None, stream=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        self.sampler.region_begin()
        l0 = self.lib.clipebc_launch_count()
        e0.record(stream)
        # bounded run-ahead: at most 2 steps in flight, so the host never sits on a full launch queue (with ~260 launches per
        # image the queue fills after 4 images; a blocked launch thread next to the nvidia-smi sampler made the timed region
        # vary by 2x between runs); the GPU still always has the next step queued
        inflight, outs = [], []
        for i in range(n):
            if len(inflight) == 2:
                inflight.pop(0).synchronize()
            o = step(i)
            if tail is not None:
                outs.append(o)
            ev = torch.cuda.Event()
            ev.record(stream)
            inflight.append(ev)
        res = tail(outs) if tail is not None else None
        e1.record(stream)
        self.barrier()
        self.sampler.region_end()
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        return ms, int(self.lib.clipebc_launch_count() - l0), res, self.sampler.windows[-1] if self.sampler.windows else None

    def reps_for(self, step, K, W):
        """Warm-up (W untimed steps), then how often the K driver steps must be repeated for a >= min-seconds region."""
        for i in range(W):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_est = max(2, min(K, 5))
        e0.record()
        for i in range(n_est):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        est = self.max_over_ranks(e0.elapsed_time(e1) / n_est)
        return max(1, int(math.ceil(self.args.min_seconds * 1000.0 / (K * max(est, 1e-3)))))

    def profile_pass(self, step, seconds_ms_per_step):
        """Per-kernel breakdown: CUDA events around every launch on its stream, in a pass that directly follows the timed
        region (same clock / power regime), ~0.4 s long."""
        n = max(3, min(400, int(math.ceil(400.0 / max(seconds_ms_per_step, 1e-3)))))
        self.lib.clipebc_profile_enable(1)
        for i in range(n):
            step(i)
        buf = ctypes.create_string_buffer(1 << 16)
        self._lib.check(self.lib.clipebc_profile_dump(buf, len(buf)), "profile_dump")
        self.lib.clipebc_profile_enable(0)
        return json.loads(buf.value.decode()), n

    def roofline(self, prof, prof_steps, windows_per_step, model_key, region_s, traffic_file=None):
        peaks = self.peaks
        gemm = {k: v for k, v in prof.items() if k.startswith("gemm")}
        gemm_ms = sum(v["ms"] for v in gemm.values())
        gemm_flops = sum(v["flops"] for v in gemm.values())
        gemm_launches = sum(v["launches"] for v in gemm.values())
        total_ms = sum(v["ms"] for v in prof.values())
        achieved = gemm_flops / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
        split = self.args.operand_dtype == "bf16"
        ft = flops_totals(model_key, split)
        st = flops_per_window(model_key, split)
        # algorithmic share of the GEMM family: everything but the attention kernel's FLOPs
        alg_gemm = sum(a for k, (a, _) in st.items() if k != "attention") * windows_per_step * prof_steps
        exe_gemm = sum(e for k, (_, e) in st.items() if k != "attention") * windows_per_step * prof_steps
        achieved_alg = achieved * (alg_gemm / exe_gemm) if exe_gemm > 0 else None
        sustained = region_s >= 1.0
        peak = peaks["bf16_sustained"] if sustained else peaks["bf16_burst"]
        traffic, traffic_note = None, "no ncu capture committed for this workload"
        if traffic_file and os.path.exists(traffic_file):
            per_tag = json.load(open(traffic_file))["per_tag"]
            num = den = 0.0
            for k, v in gemm.items():
                tag = k.split(":", 1)[1] if ":" in k else None
                if tag in per_tag:
                    nl = v["launches"] / prof_steps
                    num += nl * (per_tag[tag]["dram_read_bytes_per_launch"] + per_tag[tag]["dram_write_bytes_per_launch"])
                    den += nl
            if den > 0:
                traffic = num / den
                traffic_note = (f"bytes per launch, launch-weighted mean over the GEMMs of a step, from {os.path.relpath(traffic_file, ROOT)} "
                                "(ncu dram__bytes_read.sum + dram__bytes_write.sum, cold L2 per launch)")
        return {
            "bound": "tensor", "kernel": "gemm2_tcgen05_kernel (all epilogues)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "frac_burst": achieved / peaks["bf16_burst"], "frac_sustained": achieved / peaks["bf16_sustained"],
            "achieved_algorithmic": achieved_alg,
            "frac_algorithmic_burst": achieved_alg / peaks["bf16_burst"] if achieved_alg else None,
            "frac_algorithmic_sustained": achieved_alg / peaks["bf16_sustained"] if achieved_alg else None,
            "peak_burst": peaks["bf16_burst"], "peak_sustained": peaks["bf16_sustained"],
            "peak_source": f"{peaks['source']}: bf16_tflops_{'sustained' if sustained else 'burst'} (the instrumented pass follows a "
                           f"{region_s:.2f} s timed region back to back)",
            "traffic": traffic, "traffic_note": traffic_note,
            "algorithmic_bytes_per_launch": sum(v["bytes"] for v in gemm.values()) / max(1, gemm_launches),
            "avg_launch_ms": gemm_ms / max(1, gemm_launches), "gflop_per_launch": gemm_flops / max(1, gemm_launches) / 1e9,
            "launches_per_step": gemm_launches / prof_steps, "share_of_step": gemm_ms / total_ms if total_ms else None,
            "flops_executed_per_window": ft["executed"], "flops_algorithmic_per_window": ft["algorithmic"],
            "flops_reference_nominal_per_window": ft["nominal_reference"],
            "note": "achieved = executed 2*M*N*K of the GEMM launches (incl. split-precision extra segments and border rows) / their "
                    "summed duration; achieved_algorithmic scales it by algorithmic / executed FLOPs of those stages",
        }

    # -- workloads ----------------------------------------------------------------------------------------------------------
    def run_windows(self, name, spec, K, W, operand_dtype=None, with_e2e=True, with_profile=True):
        from oracle import weights

        model = self.model(spec["model"], operand_dtype)
        B = spec["batch"]
        g = WINDOW // model.reduction
        dev, world = self.dev, self.world
        # ring of distinct input batches larger than L2 (126 MB): 4 x 38.5 MB at 64 windows; weights + activations of one step
        # (~170 MB of 16-bit weights, ~4 MB of activations per window) exceed L2 as well, so no step sees a warm cache
        n_ring = max(2, int(math.ceil(160e6 / (B * 3 * WINDOW * WINDOW * 4))))
        ring = [weights.make_image((B, 3, WINDOW, WINDOW), seed=50 + i).to(dev) for i in range(n_ring)]
        step = lambda i: model(ring[i % n_ring])  # noqa: E731
        reps = self.reps_for(step, K, W)
        ms, launches, _, win = self.timed(step, K * reps)
        region_s = ms / 1000.0
        value = world * B * K * reps / region_s
        rec = {"workload": f"{spec['config']}: {spec['text']}", "value": value, "unit": "windows/s", "ms_per_step": ms / (K * reps),
               "windows_per_step_per_rank": B, "timed_passes": K * reps, "timed_region_s": region_s, "gpu_launches": launches,
               "images_per_sec": None, "clocks": self.sampler.summary(win),
               "l2": f"ring of {n_ring} distinct input batches ({n_ring * B * 3 * WINDOW * WINDOW * 4 / 1e6:.0f} MB > 126 MB L2); "
                     "per-step weights + activations > L2"}
        ft = flops_totals(spec["model"], (operand_dtype or self.args.operand_dtype) == "bf16")
        for regime in ("burst", "sustained"):
            pk = self.peaks[f"bf16_{regime}"]
            rec[f"tensor_frac_whole_step_executed_{regime}"] = (value / world) * ft["executed"] / 1e12 / pk
            rec[f"tensor_frac_whole_step_algorithmic_{regime}"] = (value / world) * ft["algorithmic"] / 1e12 / pk
        if with_profile:
            prof, n_prof = self.profile_pass(step, ms / (K * reps))
            traffic = os.path.join(ROOT, "profiles", "r02_gemm_dram.json") if name == HEADLINE else None
            if traffic and not os.path.exists(traffic):
                traffic = os.path.join(ROOT, "profiles", "r01k_gemm_dram.json")
            rec["roofline"] = self.roofline(prof, n_prof, B, spec["model"], region_s, traffic)
            rec["kernels"] = kernel_table(prof, n_prof)
            vit = sum(v["ms"] for k, v in prof.items()
                      if k in ("patchify", "gemm:patch_embed", "assemble_tokens", "layernorm", "gemm:qkv", "attention",
                               "gemm:out_proj", "gemm:c_fc", "gemm:c_proj")) / n_prof
            if vit > 0:
                rec["vit_forward"] = {
                    "ms_per_step_instrumented": vit,
                    "tflops_executed": B * ft["vit_executed"] / (vit / 1e3) / 1e12,
                    "frac_burst_executed": B * ft["vit_executed"] / (vit / 1e3) / 1e12 / self.peaks["bf16_burst"],
                    "frac_sustained_executed": B * ft["vit_executed"] / (vit / 1e3) / 1e12 / self.peaks["bf16_sustained"],
                    "note": "stem + 12 blocks + all LayerNorms (event-bracketed launches: no PDL overlap, an upper bound on the time); "
                            "north_star target: >= 0.60 of the dense bf16 peak"}
        if with_e2e:
            # end to end through the public API with HOST buffers (pinned): two-deep pipeline, the H2D copy of step i+1 (copy
            # stream) overlaps the kernels of step i (compute stream), D2H of every result inside the region
            host_in = [weights.make_image((B, 3, WINDOW, WINDOW), seed=50 + i).pin_memory() for i in range(min(4, n_ring))]
            dev_in = [torch.empty((B, 3, WINDOW, WINDOW), device=dev) for _ in range(2)]
            host_out = [torch.empty((B, 1, g, g)).pin_memory() for _ in range(2)]
            copy_s, comp_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            copied = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]

            def e2e_loop(n):
                for i in range(n):
                    b = i % 2
                    with torch.cuda.stream(copy_s):
                        if i >= 2:
                            copy_s.wait_event(consumed[b])
                        dev_in[b].copy_(host_in[i % len(host_in)], non_blocking=True)
                        copied[b].record(copy_s)
                    with torch.cuda.stream(comp_s):
                        comp_s.wait_event(copied[b])
                        out = model(dev_in[b])
                        consumed[b].record(comp_s)
                        host_out[b].copy_(out, non_blocking=True)
                    if i >= 1:
                        consumed[1 - b].synchronize()  # bounded run-ahead of the host: at most 2 steps in flight

            e2e_loop(W)
            self.barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.barrier()
            self.sampler.region_begin()
            s0.record(copy_s)
            comp_s.wait_event(s0)
            e2e_loop(K * reps)
            copy_s.wait_stream(comp_s)
            s1.record(copy_s)
            self.barrier()
            self.sampler.region_end()
            t2 = self.max_over_ranks(s0.elapsed_time(s1))
            rec["e2e"] = {"value": world * B * K * reps / (t2 / 1000.0), "unit": "windows/s",
                          "h2d_bytes_per_step": B * 3 * WINDOW * WINDOW * 4, "d2h_bytes_per_step": B * g * g * 4,
                          "ms_per_step": t2 / (K * reps), "timed_region_s": t2 / 1000.0,
                          "note": "pinned host buffers, 2-deep copy/compute pipeline"}
        del ring
        torch.cuda.empty_cache()
        return rec

    def run_sliding(self, name, spec, K, W, with_e2e=True, with_profile=True):
        from clip_ebc_b200 import sliding_window_predict
        from clip_ebc_b200.dist import gather_counts, shard_indices
        from oracle import weights

        model = self.model(spec["model"])
        H, Wd, stride = spec["H"], spec["W"], spec["stride"]
        dev, world, rank = self.dev, self.world, self.rank
        n_rows, n_cols = int(math.ceil((H - WINDOW) / stride) + 1), int(math.ceil((Wd - WINDOW) / stride) + 1)
        n_win = n_rows * n_cols
        r = model.reduction
        # the job: n_images synthetic images, image i on rank i % world (shard_indices). Image content cycles over N_DISTINCT
        # seeds (each image touches > 1 GB of activations >> 126 MB L2, so repeats see no warm cache).
        N_DISTINCT = 2 if H * Wd > 4e6 else 4

        def seed_of(i):
            return 600 + i % (N_DISTINCT * world)

        mine_distinct = {s: weights.make_image((1, 3, H, Wd), seed=s).to(dev)
                         for s in sorted({seed_of(i) for i in range(rank, N_DISTINCT * world * 2, world)})}

        def predict(i):  # global image index -> count [1] on the device
            return sliding_window_predict(model, mine_distinct[seed_of(i)], WINDOW, stride, return_device=True, return_count=True)[1]

        local_step = lambda j: predict(rank + j * world)  # noqa: E731  (the j-th image of this rank)
        reps = self.reps_for(local_step, K, W)
        n_local = K * reps
        n_images = n_local * world
        assert shard_indices(n_images, rank, world) == [rank + j * world for j in range(n_local)]
        # the collective is part of the warm-up too: its first call creates the NCCL communicator and loads torch's fill / index
        # kernels (lazy module loading) -- 70-200 ms that belong to no step
        gather_counts(local_step(0).reshape(1), world, rank, world)

        def tail(outs):
            # the path's only collective: ONE all-gather of the per-image counts of the whole job (4 * n_images bytes)
            return gather_counts(torch.cat([o.reshape(1) for o in outs]), n_images, rank, world)

        ms, launches, counts, win = self.timed(local_step, n_local, tail=tail)
        region_s = ms / 1000.0
        assert counts.shape == (n_images,) and bool(torch.isfinite(counts).all())
        # N-rank correctness on hardware: rank 0 predicts the distinct images of EVERY rank on its own GPU
        # (the 1-GPU path) and compares with the gathered counts bit for bit
        bit_exact = None
        if rank == 0:
            ok = True
            for i in range(min(n_images, N_DISTINCT * world)):
                img = mine_distinct.get(seed_of(i))
                if img is None:
                    img = weights.make_image((1, 3, H, Wd), seed=seed_of(i)).to(dev)
                c1 = sliding_window_predict(model, img, WINDOW, stride, return_device=True, return_count=True)[1]
                same = counts[i::N_DISTINCT * world]
                ok = ok and bool((same.view(torch.int32) == c1.view(torch.int32)).all())
            bit_exact = ok
        value = world * n_win * n_local / region_s
        rec = {"workload": f"{spec['config']}: {spec['text']}", "value": value, "unit": "windows/s", "ms_per_step": ms / n_local,
               "windows_per_step_per_rank": n_win, "timed_passes": n_local, "timed_region_s": region_s, "gpu_launches": launches,
               "images_per_sec": n_images / region_s, "n_images": n_images,
               "collective": {"op": "all_gather_into_tensor (NCCL)" if world > 1 else "none (1 rank)", "inside_timed_region": True,
                              "bytes": 4 * n_images, "calls": 1},
               "counts_bit_exact_vs_1gpu": bit_exact, "clocks": self.sampler.summary(win),
               "l2": "each image touches > 1 GB of activations (>> 126 MB L2)"}
        ft = flops_totals(spec["model"], self.args.operand_dtype == "bf16")
        for regime in ("burst", "sustained"):
            pk = self.peaks[f"bf16_{regime}"]
            rec[f"tensor_frac_whole_step_executed_{regime}"] = (value / world) * ft["executed"] / 1e12 / pk
            rec[f"tensor_frac_whole_step_algorithmic_{regime}"] = (value / world) * ft["algorithmic"] / 1e12 / pk
        if with_profile:
            prof, n_prof = self.profile_pass(local_step, ms / n_local)
            rec["roofline"] = self.roofline(prof, n_prof, n_win, spec["model"], region_s)
            rec["kernels"] = kernel_table(prof, n_prof)
        if with_e2e:
            # (1) the evaluation loop of the reference (test_nwpu.py:89-106 / eval.py:25-35) through its mirror in this package,
            # clip_ebc_b200.predict_counts: pinned HOST images in, per-image counts on the HOST out; the H2D copy of every
            # image and the one D2H of the counts are inside the timed region, the host only enqueues (no sync per image)
            from clip_ebc_b200 import predict_counts as loop_predict_counts

            host_img = [weights.make_image((1, 3, H, Wd), seed=600 + i + 10 * rank).pin_memory() for i in range(2)]
            n_e2e = max(2, n_local // 2)
            loop_predict_counts(model, [host_img[i % 2] for i in range(3)], dev, True, WINDOW, stride)
            self.barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.barrier()
            self.sampler.region_begin()
            s0.record()
            got = loop_predict_counts(model, (host_img[i % 2] for i in range(n_e2e)), dev, True, WINDOW, stride)
            s1.record()
            self.barrier()
            self.sampler.region_end()
            assert len(got) == n_e2e
            t2 = self.max_over_ranks(s0.elapsed_time(s1))
            rec["e2e"] = {"value": world * n_win * n_e2e / (t2 / 1000.0), "unit": "windows/s",
                          "h2d_bytes_per_step": 3 * H * Wd * 4, "d2h_bytes_per_step": 4,
                          "ms_per_step": t2 / n_e2e, "timed_region_s": t2 / 1000.0, "images_per_sec": world * n_e2e / (t2 / 1000.0),
                          "note": "clip_ebc_b200.predict_counts(model, pinned_host_images, device, sliding_window=True, 224, stride) -> "
                                  "host counts: the loop of test_nwpu.py; images are copied H2D one by one inside the region"}
            # (2) the reference's single call, synchronous per image: CPU image in, CPU density map out (utils/eval_utils.py:26-96)
            for i in range(2):
                sliding_window_predict(model, host_img[i % 2], WINDOW, stride)
            n_call = max(2, n_local // 4)
            self.barrier()
            self.sampler.region_begin()
            s0.record()
            for i in range(n_call):
                sliding_window_predict(model, host_img[i % 2], WINDOW, stride)
            s1.record()
            self.barrier()
            self.sampler.region_end()
            t3 = self.max_over_ranks(s0.elapsed_time(s1))
            rec["e2e_per_call"] = {"value": world * n_win * n_call / (t3 / 1000.0), "unit": "windows/s",
                                   "h2d_bytes_per_step": 3 * H * Wd * 4, "d2h_bytes_per_step": (H // r) * (Wd // r) * 4,
                                   "ms_per_step": t3 / n_call, "images_per_sec": world * n_call / (t3 / 1000.0),
                                   "note": "sliding_window_predict(model, cpu_image, 224, stride) -> CPU density map, synchronous per "
                                           "image exactly as the reference call"}
        del mine_distinct
        torch.cuda.empty_cache()
        return rec


def one_image_latency(R, spec):
    """N > 1 only: the latency of ONE image on all ranks (clip_ebc_b200.dist.sliding_window_predict_sharded: windows sharded,
    per-window maps all-gathered, every rank folds) next to the same image on rank 0 alone; strong scaling, bit-exact check."""
    from clip_ebc_b200 import sliding_window_predict
    from clip_ebc_b200.dist import sliding_window_predict_sharded
    from oracle import weights

    model = R.model(spec["model"])
    H, Wd, stride = spec["H"], spec["W"], spec["stride"]
    img = weights.make_image((1, 3, H, Wd), seed=4242).to(R.dev)  # the same image on every rank
    sharded = lambda j: sliding_window_predict_sharded(model, img, WINDOW, stride, R.rank, R.world, return_count=True)[1]  # noqa: E731
    single = lambda j: sliding_window_predict(model, img, WINDOW, stride, return_device=True, return_count=True)[1]  # noqa: E731
    out = {"workload": f"{spec['config']}: one image on {R.world} GPUs", "scaling": "strong"}
    d_s = sliding_window_predict_sharded(model, img, WINDOW, stride, R.rank, R.world)
    d_1 = sliding_window_predict(model, img, WINDOW, stride, return_device=True)
    same = torch.tensor([int(torch.equal(d_s.view(torch.int32), d_1.view(torch.int32)))], device=R.dev)
    if R.world > 1:
        R.dist.all_reduce(same, op=R.dist.ReduceOp.MIN)
    out["density_bit_exact_vs_1gpu_on_every_rank"] = bool(same.item())
    for name, step in (("sharded", sharded), ("one_gpu", single)):
        for _ in range(2):
            step(0)
        n = 3
        while True:
            ms, _, _, _ = R.timed(step, n)
            if ms >= 1000.0 or n >= 4096:
                break
            n = int(math.ceil(n * max(1.3, 1100.0 / max(ms, 1.0))))
        out[f"{name}_ms_per_image"] = ms / n
    out["speedup"] = out["one_gpu_ms_per_image"] / out["sharded_ms_per_image"]
    out["collective"] = {"op": "all_gather_into_tensor (NCCL) of the per-window maps", "bytes": 4 * (WINDOW // model.reduction) ** 2 *
                         (int(math.ceil((H - WINDOW) / stride) + 1) * int(math.ceil((Wd - WINDOW) / stride) + 1))}
    return out


def kernel_table(prof, prof_steps):
    return {k: {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["launches"] / prof_steps,
                "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["ms"] > 0 and v["flops"] > 0 else None,
                "gbs": (v["bytes"] / (v["ms"] / 1e3) / 1e9) if v["ms"] > 0 and v["bytes"] > 0 else None}
            for k, v in prof.items()}


def main():
    args = parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    R = Runner(args)
    K, W = args.steps, args.warmup
    if args.impl == "torch_gpu":
        if R.rank == 0:
            _, parts = build_model(R.dev, "r8_deep")
            lb = gpu_library_baseline(R.dev, parts, WORKLOADS[HEADLINE]["batch"], seconds=max(1.0, args.min_seconds))
            best = max(v["windows_per_sec"] for k, v in lb.items() if isinstance(v, dict))
            emit({"impl": "torch_gpu", "metric": "windows_per_sec", "value": best, "unit": "windows/s", "n_gpus": 1, "steps": K,
                  "warmup": 3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 autocast / tf32 / fp32",
                  "data": "synthetic", "config": {"workload": f"{WORKLOADS[HEADLINE]['config']}: {WORKLOADS[HEADLINE]['text']}"},
                  "gpu_library_baseline": lb, "gpu_launches": 0})
        if R.world > 1:
            R.dist.destroy_process_group()
        return

    names = list(WORKLOADS) if args.workload == "all" else [args.workload]
    head_name = HEADLINE if HEADLINE in names else names[0]
    R.sampler.start()
    R.model(WORKLOADS[head_name]["model"])
    R.sampler.wait_first_sample()
    results = {}
    for name in names:
        spec = WORKLOADS[name]
        if spec["kind"] == "windows":
            results[name] = R.run_windows(name, spec, K, W)
        else:
            results[name] = R.run_sliding(name, spec, K, W)
    # the precision BASELINE.json configs[1] names: the headline workload once more with bf16 operands (same kind::f16 rate)
    bf16 = None
    if head_name == HEADLINE and args.operand_dtype != "bf16":
        r16 = R.run_windows(HEADLINE, WORKLOADS[HEADLINE], K, W, operand_dtype="bf16", with_e2e=False, with_profile=False)
        bf16 = {k: r16[k] for k in ("value", "unit", "ms_per_step", "timed_region_s", "timed_passes")}
    clocks = R.sampler.summary()
    latency = one_image_latency(R, WORKLOADS["qnrf112"]) if R.world > 1 and args.workload == "all" else None
    R.sampler.stop()

    head = results[head_name]
    cpu_baseline = lib_baseline = None
    if R.rank == 0 and R.world == 1:
        parts = R.parts[WORKLOADS[HEADLINE]["model"]] if WORKLOADS[HEADLINE]["model"] in R.parts else build_model(R.dev, "r8_deep")[1]
        if not args.no_gpu_library_baseline:
            lib_baseline = gpu_library_baseline(R.dev, parts, WORKLOADS[HEADLINE]["batch"])
        if not args.no_cpu_baseline:
            wps, cores, times = cpu_port_windows_per_sec(8, 5, parts)
            cpu_baseline = {"value": wps, "unit": "windows/s", "cores": cores, "kind": "port",
                            "sample": f"8 of the 64 windows of one step, median of 5 runs after 2 warm-ups "
                                      f"({statistics.median(times):.2f} s per run, min {min(times):.2f}, max {max(times):.2f}), "
                                      "oracle/clip_ebc_oracle.py fp32"}

    if R.rank == 0:
        od = args.operand_dtype
        line = {
            "metric": "windows_per_sec", "value": head["value"], "unit": "windows/s", "n_gpus": R.world, "steps": K, "warmup": W,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": od,
            "dtype_note": "16-bit tensor-core operands on tcgen05 kind::f16: fp16 by default (11-bit mantissa; CLIP's released weights "
                          "are fp16), bf16 selectable at the same rate (`bf16_operands` is the headline workload measured that way); "
                          "fp32 accumulate / residual stream / LayerNorm statistics / softmax / head",
            "data": "synthetic",
            "config": {"workload": head["workload"], "l2": head["l2"],
                       "global_batch_windows": R.world * head["windows_per_step_per_rank"],
                       "parallelism": f"dp{R.world} (independent windows / whole images per rank, one process per GPU)",
                       "weights": "seeded random init with the reference's init distributions (oracle/weights.py)",
                       "timed_passes": head["timed_passes"], "min_seconds": args.min_seconds},
            "timed_region_s": head["timed_region_s"], "timed_passes": head["timed_passes"],
            "clocks": clocks, "e2e": head.get("e2e"), "gpu_launches": head["gpu_launches"],
            "roofline": head.get("roofline"), "cpu_baseline": cpu_baseline, "gpu_library_baseline": lib_baseline,
            "images_per_sec": results["sliding"]["images_per_sec"] if "sliding" in results else head.get("images_per_sec"),
            "images_per_sec_e2e": results["sliding"].get("e2e", {}).get("images_per_sec") if "sliding" in results else None,
            "counts_bit_exact_vs_1gpu": {k: v["counts_bit_exact_vs_1gpu"] for k, v in results.items() if "counts_bit_exact_vs_1gpu" in v},
            "bf16_operands": bf16,
            "one_image_latency": latency,
            "vit_forward": head.get("vit_forward"),
            "kernels": head.get("kernels"),
            "workloads": {k: {kk: vv for kk, vv in v.items() if kk != "kernels" or k != head_name} for k, v in results.items()},
        }
        for kk in ("tensor_frac_whole_step_executed_burst", "tensor_frac_whole_step_executed_sustained",
                   "tensor_frac_whole_step_algorithmic_burst", "tensor_frac_whole_step_algorithmic_sustained"):
            line[kk] = head.get(kk)
        emit(line)
    if R.world > 1:
        R.dist.destroy_process_group()


if __name__ == "__main__":
    main()
