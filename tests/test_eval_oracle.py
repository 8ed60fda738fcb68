"""CPU: the oracle of the steps either side of the hot path (oracle/eval_oracle.py) against the fixtures produced by
the reference's own resize_density_map / calculate_errors / Resize2Multiple / ZeroPad2Multiple / Normalize, and the
host-side mirrors in clip_ebc_b200 (size arithmetic, assertions, result-file format; no GPU work)."""
import numpy as np
import pytest
import torch

from oracle import eval_oracle as E
from oracle.golden_cases import (RESIZE_DENSITY_CASES, SUB_X, SUB_Y, TRANSFORM_CASES, make_density, make_points,
                                 make_u8_image)

from . import parity

ATOL = 2e-6  # same arithmetic as the reference, possibly other CPU kernels: fp32 rounding of O(1) values


@pytest.mark.parametrize("case", RESIZE_DENSITY_CASES, ids=[c["name"] for c in RESIZE_DENSITY_CASES])
def test_resize_density_map_oracle(case):
    gold = parity.load_golden(case["name"])
    x = make_density(case["shape"], case["seed"], case["zero_image"])
    y = E.resize_density_map(x, case["size"])
    assert y.shape == gold["out"].shape
    assert np.abs(y.numpy() - gold["out"]).max() < ATOL
    # Reference behaviour, kept: the map is multiplied by sum(resized) / sum(x) (not its inverse), so the output sums to
    # sum(resized)^2 / sum(x); an all-zero input gives 0/0 -> nan_to_num -> an all-zero map.
    if case["zero_image"] is None:
        want = float(gold["out_sum"][0])
        assert abs(float(y.double().sum()) - want) < 1e-5 * abs(want)
    else:
        assert float(y.abs().max()) == 0.0 and float(np.abs(gold["out"]).max()) == 0.0


@pytest.mark.parametrize("case", TRANSFORM_CASES, ids=[c["name"] for c in TRANSFORM_CASES])
def test_transforms_oracle(case):
    gold = parity.load_golden(case["name"])
    u8 = make_u8_image(case["shape"], case["seed"])
    h, w = case["shape"][1:]
    pts = make_points(case["n_points"], h, w, case["seed"] + 1000)
    for tag in ("resize", "pad"):
        out = E.preprocess(u8, tag, case["window"], case["stride"])
        assert tuple(out.shape) == tuple(gold[f"{tag}_shape"])
        assert np.abs(out[:, ::SUB_Y, ::SUB_X].numpy() - gold[f"{tag}_sub"]).max() < 2e-5
        assert np.abs(out.double().sum(dim=(1, 2)).numpy() - gold[f"{tag}_chan_sum"]).max() < 1e-5 * out[0].numel()
    nh, nw = E.resize2multiple_size(h, w, case["window"], case["stride"])
    lab = E.resize_labels(pts, h, w, nh, nw)
    assert np.array_equal(lab.numpy(), gold["resize_labels"])
    assert np.array_equal(pts.numpy(), gold["pad_labels"])  # padding right/bottom leaves the points alone


def test_calculate_errors_known_answer():
    from clip_ebc_b200.eval_utils import calculate_errors

    gold = parity.load_golden("eval_calculate_errors")
    for fn in (E.calculate_errors, calculate_errors):
        err = fn(gold["pred"], gold["gt"])
        assert err["mae"] == float(gold["mae"]) and err["rmse"] == float(gold["rmse"])  # same numpy expression: bit-exact
    with pytest.raises(AssertionError):
        calculate_errors([1.0], np.array([1.0]))
    with pytest.raises(AssertionError):
        calculate_errors(np.array([1.0, 2.0]), np.array([1.0]))


SIZES = [(500, 731), (300, 350), (200, 260), (448, 672), (224, 224), (225, 225), (1080, 1920), (337, 1000), (280, 392)]


@pytest.mark.parametrize("window,stride", [(224, 112), (224, 224), ((224, 448), (112, 200))])
def test_new_size_arithmetic_matches_oracle(window, stride):
    from clip_ebc_b200.transforms import Resize2Multiple, ZeroPad2Multiple

    r, z = Resize2Multiple(window, stride), ZeroPad2Multiple(window, stride)
    wh, ww = (window, window) if isinstance(window, int) else window
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    for h, w in SIZES:
        assert r.new_size(h, w) == E.resize2multiple_size(h, w, window, stride)
        assert z.new_size(h, w) == E.zeropad2multiple_size(h, w, window, stride)
        for nh, nw in (r.new_size(h, w), z.new_size(h, w)):
            assert nh >= wh and nw >= ww and (nh - wh) % sh == 0 and (nw - ww) % sw == 0
        assert z.new_size(h, w)[0] >= h or h < wh


def test_transform_constructor_assertions_match_reference_messages():
    from clip_ebc_b200.transforms import Resize2Multiple, ZeroPad2Multiple

    for cls in (Resize2Multiple, ZeroPad2Multiple):
        with pytest.raises(AssertionError, match="stride should be no larger than window_size"):
            cls(224, 300)
        with pytest.raises(AssertionError, match="window_size should be positive"):
            cls((0, 224), 1)
        with pytest.raises(AssertionError, match="stride should be a tuple"):
            cls(224, (1, 2, 3))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            cls(224, 112)(torch.zeros(3, 300, 300), torch.zeros(0, 2))  # CPU tensors are refused, never silently handled


def test_nwpu_result_file_format(tmp_path):
    from clip_ebc_b200.eval_loop import nwpu_result_text, write_nwpu_results

    ids, preds = ["3610", "3611", "3612"], [12.5, 0.0, 1234.56789]
    text = nwpu_result_text(ids, preds)
    assert text == E.nwpu_result_text(ids, preds) == "3610 12.5\n3611 0.0\n3612 1234.56789"  # no trailing newline
    path = tmp_path / "out.txt"
    write_nwpu_results(str(path), ids, preds)
    assert path.read_text() == text
    with pytest.raises(AssertionError):
        nwpu_result_text(ids, preds[:2])


def test_evaluate_refuses_foreign_models_and_cpu():
    from clip_ebc_b200.eval_loop import evaluate

    with pytest.raises(TypeError):
        evaluate(torch.nn.Linear(1, 1), [], torch.device("cpu"))


def test_resize_density_map_shape_errors_like_reference():
    from clip_ebc_b200.eval_utils import resize_density_map

    with pytest.raises(RuntimeError):
        resize_density_map(torch.zeros(2, 1, 8, 8), (16, 16))  # the reference's x * scale_factor fails to broadcast too
