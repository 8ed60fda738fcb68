"""TEST INFRASTRUCTURE -- generates tests/golden/eval_*.npz by running the REAL reference functions of the steps either
side of the hot path (utils/eval_utils.py: resize_density_map, calculate_errors; datasets/transforms.py:
Resize2Multiple, ZeroPad2Multiple; torchvision Normalize as used by datasets/crowd.py:64) on seeded inputs.
Runs only in the build container (needs /root/reference); the fixtures are committed.

    python -m oracle.make_golden_eval
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.golden_cases import (RESIZE_DENSITY_CASES, SUB_X, SUB_Y, TRANSFORM_CASES, make_density, make_points,  # noqa: E402
                                 make_u8_image)

OUT = os.path.join(ROOT, "tests", "golden")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main() -> None:
    assert ref_loader.available(), "needs /root/reference"
    ev = _load("ref_eval_utils_only", f"{ref_loader.REF}/utils/eval_utils.py")
    tr = _load("ref_transforms", f"{ref_loader.REF}/datasets/transforms.py")
    from torchvision.transforms import Normalize  # the class datasets/crowd.py:64 instantiates

    norm = Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    os.makedirs(OUT, exist_ok=True)

    for c in RESIZE_DENSITY_CASES:
        x = make_density(c["shape"], c["seed"], c["zero_image"])
        y = ev.resize_density_map(x, c["size"])
        np.savez_compressed(os.path.join(OUT, c["name"] + ".npz"), out=y.numpy(),
                            out_sum=y.sum(dim=(1, 2, 3)).double().numpy(), in_sum=x.sum(dim=(1, 2, 3)).double().numpy())
        print(c["name"], tuple(y.shape))

    for c in TRANSFORM_CASES:
        u8 = make_u8_image(c["shape"], c["seed"])
        pts = make_points(c["n_points"], c["shape"][1], c["shape"][2], c["seed"] + 1000)
        img = torch.from_numpy(u8).float() / 255.0  # datasets/crowd.py:218
        res = {}
        for tag, t in (("resize", tr.Resize2Multiple(c["window"], c["stride"])),
                       ("pad", tr.ZeroPad2Multiple(c["window"], c["stride"]))):
            o, lab = t(img.clone(), pts.clone())
            o = norm(o)
            res[f"{tag}_shape"] = np.array(o.shape)
            res[f"{tag}_sub"] = o[:, ::SUB_Y, ::SUB_X].numpy()
            res[f"{tag}_chan_sum"] = o.double().sum(dim=(1, 2)).numpy()
            res[f"{tag}_labels"] = lab.numpy()
        np.savez_compressed(os.path.join(OUT, c["name"] + ".npz"), **res)
        print(c["name"], {k: v.shape for k, v in res.items()})

    # calculate_errors known answers
    rng = np.random.Generator(np.random.PCG64(77))
    pred = rng.uniform(0, 500, 37)
    gt = np.rint(pred + rng.normal(0, 30, 37))
    err = ev.calculate_errors(pred, gt)
    np.savez_compressed(os.path.join(OUT, "eval_calculate_errors.npz"), pred=pred, gt=gt, mae=np.float64(err["mae"]),
                        rmse=np.float64(err["rmse"]))
    print("eval_calculate_errors", err)


if __name__ == "__main__":
    main()
