"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the REAL reference (imported from /root/reference via
oracle/ref_loader.py) on seeded synthetic weights and inputs. Runs only in the build container; the fixtures it writes
are committed so that the GPU box (which has no /root/reference) can pin both the oracle and the CUDA path to the
reference's own outputs.

    python -m oracle.make_golden            # rewrites tests/golden/

Every case is described by oracle/golden_cases.py; inputs and weights are regenerated from seeds (numpy PCG64), only
the reference's outputs are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader, weights  # noqa: E402
from oracle.golden_cases import CASES, ORACLE_ONLY_CASES, backbone_of, case_inputs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def run_case(case: dict) -> dict:
    sd, tf, bins, anchors, reduction, x = case_inputs(case)
    model = ref_loader.build_reference_model(sd, tf, bins, anchors, reduction, num_vpt=case["num_vpt"],
                                             deep_vpt=case["deep_vpt"], input_size=224, backbone=backbone_of(case))
    _, ev = ref_loader.load_reference()
    out = {}
    resnet = backbone_of(case).startswith("resnet")
    if case["kind"] == "forward" and resnet:
        taps = {}
        hooks = [
            model.image_encoder.avgpool.register_forward_hook(lambda m, i, o: taps.__setitem__("stem", o.detach().clone())),
            model.image_encoder.layer4.register_forward_hook(lambda m, i, o: taps.__setitem__("layer4", o.detach().clone())),
            model.image_decoder.register_forward_hook(lambda m, i, o: taps.__setitem__("decoder", o.detach().clone())),
        ]
        model.training = True
        with torch.no_grad():
            logits, exp = model(x)
        model.training = False
        for h in hooks:
            h.remove()
        out["logits"] = logits.numpy()
        out["exp"] = exp.numpy()
        out["tap_stem"] = taps["stem"][0, :8, :4, :4].numpy()
        out["tap_layer4"] = taps["layer4"][0, :16, :2, :2].numpy()
        out["tap_decoder"] = taps["decoder"][0, :16, :2, :2].numpy()
    elif case["kind"] == "forward":
        taps = {}
        hooks = [
            model.image_encoder.ln_pre.register_forward_hook(lambda m, i, o: taps.__setitem__("ln_pre", o.detach().clone())),
            model.image_encoder.ln_post.register_forward_hook(lambda m, i, o: taps.__setitem__("ln_post", o.detach().clone())),
            model.image_decoder.register_forward_hook(lambda m, i, o: taps.__setitem__("decoder", o.detach().clone())),
        ]
        model.training = True  # top-level flag only: sub-modules stay in eval mode; forward returns (logits, exp)
        with torch.no_grad():
            logits, exp = model(x)
        model.training = False
        for h in hooks:
            h.remove()
        out["logits"] = logits.numpy()
        out["exp"] = exp.numpy()
        # small slices of intermediate activations of window 0 (debugging aid for the oracle and the kernels)
        out["tap_ln_pre"] = taps["ln_pre"][0, :4, :].numpy()          # [4 tokens, 768]
        out["tap_ln_post"] = taps["ln_post"][0, 1:5, :].numpy()       # patch tokens 0..3 after ln_post
        out["tap_decoder"] = taps["decoder"][0, :, :2, :2].numpy()    # [768, 2, 2]
    else:
        with torch.no_grad():
            dens = ev.sliding_window_predict(model, x, case["window"], case["stride"])
        out["density"] = dens.numpy()
        out["count"] = np.array(dens.sum(dim=(1, 2, 3)).item(), dtype=np.float64)
    return out


def main() -> None:
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    only = sys.argv[1:]  # optional: names of the cases to (re)generate
    for case in CASES + ORACLE_ONLY_CASES:
        if only and case["name"] not in only:
            continue
        res = run_case(case)
        path = os.path.join(OUT, case["name"] + ".npz")
        np.savez_compressed(path, **res)
        print(case["name"], {k: v.shape for k, v in res.items()}, f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
