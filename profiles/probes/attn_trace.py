"""clock64 time line of CTA 0 of the attention kernel (needs a library built with -DCLIPEBC_ATTN_TRACE:
   make -C clip_ebc_b200/csrc clean && make -C clip_ebc_b200/csrc EXTRA=-DCLIPEBC_ATTN_TRACE).

   python profiles/probes/attn_trace.py [n_win]
Prints, per role (warp) and tile, the cycle stamps of the phase boundaries relative to the first event, and the mean phase
durations of the steady state.
"""
import ctypes
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from clip_ebc_b200 import _lib, ops  # noqa: E402

n_win = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t_live, heads, n_const = 197, 12, 32
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(0)
qkv = (torch.randn((n_win * t_live, 3 * 64 * heads), generator=g) * 1.5).to(dev).to(torch.bfloat16)
ckv = (torch.randn((n_const, 3 * 64 * heads), generator=g) * 1.5).to(dev).to(torch.bfloat16)
lib = _lib.load()
fn = lib.clipebc_debug_attn_trace
fn.restype = ctypes.c_int
N = 20 * 16 * 16
buf = (ctypes.c_longlong * N)()
for _ in range(3):
    ops.attention(qkv, n_win, t_live, const_kv=ckv, heads=heads)
torch.cuda.synchronize()
fn(buf, N)  # drop the warm-up records
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
ops.attention(qkv, n_win, t_live, const_kv=ckv, heads=heads)
ev1.record()
torch.cuda.synchronize()
n = fn(buf, N)
print(f"kernel {ev0.elapsed_time(ev1) * 1e3:.1f} us")
ROLE_EVENTS = {0: (20, 21, 22), 1: (0, 4, 1, 2), 3: (0, 4, 1, 2)}
rec = []
for w in range(20):
    for k in range(16):
        for ev in ROLE_EVENTS.get(w, (10, 11, 12, 13, 14, 15)):
            t = buf[(w * 16 + k) * 16 + (ev & 15)]
            if t:
                rec.append((w, ev, k, t))
t0 = min(r[3] for r in rec)
t_end = max(r[3] for r in rec)
print(f"CTA 0 active for {t_end - t0} cycles")
by = defaultdict(dict)
for w, ev, k, t in rec:
    by[(w, k)][ev] = t - t0
names = {5: "PV0", 6: "PV1", 7: "PV2", 8: "PV3", 0: "top", 4: "qk_full", 1: "S issue", 2: "PV issue", 10: "wait S", 11: "S ready", 12: "softmax done", 13: "O ready",
         14: "O read", 15: "stored", 20: "tma top", 21: "qk slot", 22: "v slot"}
for w in sorted({w for (w, _) in by}):
    print(f"--- warp {w}")
    for k in sorted(k for (ww, k) in by if ww == w)[:7]:
        evs = by[(w, k)]
        print(f"  tile {k}: " + "  ".join(f"{names.get(e, e)}={evs[e]}" for e in sorted(evs, key=lambda e: evs[e])))
# steady-state phase durations of the softmax warps (lane quadrant 0 and 2 of both chains)
for w in (4, 6, 8, 10):
    ks = sorted(k for (ww, k) in by if ww == w)[1:-1]
    ks = [k for k in ks if all(e in by[(w, k)] for e in (10, 11, 12, 13, 14, 15))]
    if len(ks) < 2:
        continue
    def mean(f):
        return sum(f(by[(w, k)]) for k in ks) / len(ks)
    print(f"warp {w}: wait S {mean(lambda e: e[11] - e[10]):.0f}  softmax {mean(lambda e: e[12] - e[11]):.0f}  "
          f"wait O (P.V) {mean(lambda e: e[13] - e[12]):.0f}  O read {mean(lambda e: e[14] - e[13]):.0f}  store {mean(lambda e: e[15] - e[14]):.0f}"
          f"  period {(by[(w, ks[-1])][10] - by[(w, ks[0])][10]) / (len(ks) - 1):.0f}")
