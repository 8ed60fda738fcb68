"""Parity metrics of BASELINE.json's north_star, shared by the CPU (oracle) and GPU (CUDA path) tests."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star gates
DENSITY_MAX_REL = 2e-2    # max |d - d_ref| / max |d_ref|
COUNT_REL = 5e-3          # |sum d - sum d_ref| / |sum d_ref|
ARGMAX_AGREE = 0.995      # fraction of cells whose bin argmax equals the reference's


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def density_max_rel(d, ref):
    d, ref = np.asarray(d, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(d - ref).max() / np.abs(ref).max())


def count_rel(d, ref):
    d, ref = np.asarray(d, np.float64), np.asarray(ref, np.float64)
    return float(abs(d.sum() - ref.sum()) / abs(ref.sum()))


def argmax_agreement(logits, ref_logits):
    return float((np.asarray(logits).argmax(1) == np.asarray(ref_logits).argmax(1)).mean())


def margin_conditioned_agreement(logits, ref_logits, min_margin):
    """Agreement restricted to cells whose reference top1-top2 logit margin exceeds min_margin (reported, not gated)."""
    ref = np.asarray(ref_logits)
    s = np.sort(ref, axis=1)
    m = (s[:, -1] - s[:, -2]) > min_margin
    if not m.any():
        return 1.0
    return float((np.asarray(logits).argmax(1) == ref.argmax(1))[m].mean())
