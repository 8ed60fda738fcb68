"""CPU: the oracle (oracle/clip_ebc_oracle.py) against every golden fixture produced by the real reference."""
import numpy as np
import pytest
import torch

from oracle import clip_ebc_oracle as O
from oracle.golden_cases import CASES, ORACLE_ONLY_CASES, case_inputs

from . import parity

# The fixtures were produced by the reference on the build container's CPU; another host CPU may pick different
# oneDNN / MKL kernels, so allow fp32 rounding noise (values are O(1)..O(10)).
ATOL = 2e-4


@pytest.mark.parametrize("case", CASES + ORACLE_ONLY_CASES, ids=[c["name"] for c in CASES + ORACLE_ONLY_CASES])
def test_oracle_matches_reference_fixture(case):
    torch.manual_seed(0)
    sd, tf, bins, anchors, reduction, x = case_inputs(case)
    gold = parity.load_golden(case["name"])
    if case["kind"] == "forward" and O.is_resnet(sd):
        taps = {}
        logits, exp = O.clip_ebc_forward(x, sd, tf, anchors, reduction, taps=taps)
        scale = max(1.0, float(np.abs(gold["tap_decoder"]).max()))  # untrained 2048-channel maps reach O(100)
        assert np.abs(logits.numpy() - gold["logits"]).max() < ATOL * scale
        assert np.abs(exp.numpy() - gold["exp"]).max() < ATOL
        assert np.abs(taps["stem"][0, :8, :4, :4].numpy() - gold["tap_stem"]).max() < ATOL
        assert np.abs(taps["layer4"][0, :16, :2, :2].numpy() - gold["tap_layer4"]).max() < ATOL * scale
        assert np.abs(taps["decoder"][0, :16, :2, :2].numpy() - gold["tap_decoder"]).max() < ATOL * scale
        assert parity.argmax_agreement(logits.numpy(), gold["logits"]) >= 0.999
    elif case["kind"] == "forward":
        taps = {}
        logits, exp = O.clip_ebc_forward(x, sd, tf, anchors, reduction, case["num_vpt"], case["deep_vpt"], 224, taps)
        assert np.abs(logits.numpy() - gold["logits"]).max() < ATOL
        assert np.abs(exp.numpy() - gold["exp"]).max() < ATOL
        assert np.abs(taps["ln_pre"][0, :4].numpy() - gold["tap_ln_pre"]).max() < ATOL
        assert np.abs(taps["ln_post"][0, :, :, :].flatten(1).t()[:4].numpy() - gold["tap_ln_post"]).max() < ATOL
        assert np.abs(taps["decoder"][0, :, :2, :2].numpy() - gold["tap_decoder"]).max() < ATOL
        assert parity.argmax_agreement(logits.numpy(), gold["logits"]) >= 0.999
    else:
        dens = O.sliding_window_predict(x, sd, tf, anchors, reduction, case["window"], case["stride"], case["num_vpt"],
                                        case["deep_vpt"], 224)
        assert dens.shape == gold["density"].shape
        assert np.abs(dens.numpy() - gold["density"]).max() < ATOL
        assert parity.count_rel(dens.numpy(), gold["density"]) < 1e-6


def test_window_origins_known_answers():
    """SURVEY.md Appendix B (from utils/eval_utils.py:54-66)."""
    r, c = O.window_origins(448, 448, (224, 224), (224, 224))
    assert (r, c) == ([0, 224], [0, 224])
    r, c = O.window_origins(1536, 2048, (224, 224), (112, 112))
    assert len(r) == 13 and len(c) == 18 and r[-3:] == [1120, 1232, 1312] and c[-3:] == [1680, 1792, 1824]
    r, c = O.window_origins(3072, 4096, (224, 224), (224, 224))
    assert len(r) == 14 and len(c) == 19 and r[-2:] == [2688, 2848] and c[-2:] == [3808, 3872]
    r, c = O.window_origins(3072, 4096, (224, 224), (112, 112))
    assert len(r) == 27 and len(c) == 36 and r[-2:] == [2800, 2848] and c[-2:] == [3808, 3872]
    r, c = O.window_origins(448, 672, (224, 224), (112, 112))
    assert r == [0, 112, 224] and c == [0, 112, 224, 336, 448]
