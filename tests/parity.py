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


CELL_REL_FLOOR = 0.05     # per-cell check: cells whose reference value exceeds this fraction of the map maximum
CELL_MAX_REL = 2e-2       # ... must each be within 2e-2 of their OWN reference value (not of the global maximum)


def cell_max_rel(d, ref, floor=CELL_REL_FLOOR):
    """max over cells with |ref| > floor * max|ref| of |d - ref| / |ref| -- a per-cell relative error, so that a large
    global maximum cannot hide errors on small-valued cells (density_max_rel normalises by the global maximum)."""
    d, ref = np.asarray(d, np.float64), np.asarray(ref, np.float64)
    m = np.abs(ref) > floor * np.abs(ref).max()
    return float((np.abs(d - ref)[m] / np.abs(ref)[m]).max()) if m.any() else 0.0


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
