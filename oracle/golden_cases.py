"""TEST INFRASTRUCTURE -- the parity cases shared by oracle/make_golden.py (reference run), tests/test_oracle.py (oracle
vs fixtures, CPU) and tests/test_parity_gpu.py (CUDA path vs fixtures + oracle, B200).

kind "forward": model(x) on explicit windows  -> logits [B,N,g,g], exp [B,1,g,g]     (models/clip/model.py:191-217)
kind "sliding": sliding_window_predict(image) -> density [1,1,H//r,W//r], count      (utils/eval_utils.py:26-96)
"""
from __future__ import annotations

from . import weights

CASES = [
    # BASELINE.json configs[0]: 448x448, window 224, stride 224, r8, truncation 4, deep VPT 32, default-init weights
    dict(name="c1_sliding_448_s224_r8_deep", kind="sliding", bins="r8_t4_nwpu", deep_vpt=True, num_vpt=32,
         variant="default", wseed=0, xseed=1, shape=(1, 3, 448, 448), window=224, stride=224),
    # the 4 windows of the same image through model(x): logits for the argmax-agreement gate
    dict(name="c1_forward_r8_deep", kind="forward", bins="r8_t4_nwpu", deep_vpt=True, num_vpt=32,
         variant="default", wseed=0, xseed=1, shape=(4, 3, 224, 224)),
    # overlapping windows (stride 112, 3x5 = 15 windows, shared patch grid), randomised biases / BN statistics
    dict(name="sliding_448x672_s112_r8_deep_stress", kind="sliding", bins="r8_t4_nwpu", deep_vpt=True, num_vpt=32,
         variant="stress", wseed=3, xseed=2, shape=(1, 3, 448, 672), window=224, stride=112),
    # clamped last windows that are NOT on the 16-pixel patch grid (origins 0 / 76 and 0 / 200 / 276): per-window unfold
    dict(name="sliding_300x500_s200_r8_offgrid", kind="sliding", bins="r8_t4_nwpu", deep_vpt=True, num_vpt=32,
         variant="stress", wseed=3, xseed=4, shape=(1, 3, 300, 500), window=224, stride=200),
    # BASELINE.json configs[3]: reduction 16 / 32 bin sets with shallow VPT
    dict(name="forward_r16_shallow_stress", kind="forward", bins="r16_t8_qnrf", deep_vpt=False, num_vpt=32,
         variant="stress", wseed=5, xseed=6, shape=(8, 3, 224, 224)),
    dict(name="forward_r32_shallow_stress", kind="forward", bins="r32_t19_qnrf", deep_vpt=False, num_vpt=32,
         variant="stress", wseed=5, xseed=7, shape=(16, 3, 224, 224)),
    # r32 with overlapping stride 112: window cell origins floor (112 // 32 = 3), reference behaviour kept
    dict(name="sliding_448x672_s112_r32_shallow", kind="sliding", bins="r32_t19_qnrf", deep_vpt=False, num_vpt=32,
         variant="default", wseed=8, xseed=9, shape=(1, 3, 448, 672), window=224, stride=112),
    # non-224 windows: bicubic positional-embedding interpolation (_clip/image_encoder.py:183-198)
    dict(name="forward_r8_deep_160x192", kind="forward", bins="r8_t4_nwpu", deep_vpt=True, num_vpt=32,
         variant="stress", wseed=10, xseed=11, shape=(2, 3, 160, 192)),
]


def case_inputs(case: dict):
    """-> (state_dict, text_features, bins, anchors, reduction, x) regenerated from the case's seeds."""
    reduction, bins, anchors = weights.bins_and_anchors(case["bins"])
    sd = weights.make_state_dict(case["wseed"], input_size=224, num_vpt=case["num_vpt"], deep_vpt=case["deep_vpt"],
                                 variant=case["variant"])
    tf = weights.make_text_features(len(bins), seed=100 + case["wseed"])
    x = weights.make_image(case["shape"], seed=case["xseed"])
    return sd, tf, bins, anchors, reduction, x
