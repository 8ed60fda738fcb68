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
    # the ViT-B/32 backbone behind the same boundary (patch 32: 7x7 patches per 224 window, x4 bilinear
    # resample to the reduction-8 grid, models/clip/model.py:20,195-196)
    dict(name="b32_forward_r8_deep", kind="forward", bins="r8_t4_nwpu", deep_vpt=True, num_vpt=32,
         variant="stress", wseed=12, xseed=13, shape=(4, 3, 224, 224), patch=32),
    # overlapping stride 112 is off the 32-pixel patch grid (per-window unfold), 448 x 672 -> 15 windows
    dict(name="b32_sliding_448x672_s112_r8_deep", kind="sliding", bins="r8_t4_nwpu", deep_vpt=True, num_vpt=32,
         variant="default", wseed=14, xseed=15, shape=(1, 3, 448, 672), window=224, stride=112, patch=32),
    # stride 224 keeps the origins on the patch grid (shared patch grid), reduction 16, shallow VPT
    dict(name="b32_sliding_448x448_s224_r16_shallow", kind="sliding", bins="r16_t8_qnrf", deep_vpt=False, num_vpt=32,
         variant="stress", wseed=16, xseed=17, shape=(1, 3, 448, 448), window=224, stride=224, patch=32),
]


# Windows with more than 256 tokens (the tcgen05 attention tile): 320 x 320 windows have 400 patches (433 tokens with
# cls + prompts, positional embedding resized 14 -> 20), 448 x 448 windows have 784 (817 tokens)
CASES += [
    dict(name="forward_r8_deep_320x320", kind="forward", bins="r8_t4_nwpu", deep_vpt=True, num_vpt=32,
         variant="stress", wseed=18, xseed=19, shape=(2, 3, 320, 320)),
    dict(name="sliding_448x896_w448_r8_deep", kind="sliding", bins="r8_t4_nwpu", deep_vpt=True, num_vpt=32,
         variant="default", wseed=20, xseed=21, shape=(1, 3, 448, 896), window=448, stride=448),
    dict(name="sliding_448x672_w448_s224_r16_shallow", kind="sliding", bins="r16_t8_qnrf", deep_vpt=False, num_vpt=32,
         variant="stress", wseed=22, xseed=23, shape=(1, 3, 448, 672), window=448, stride=224),
]

# trained-CLIP-like activation outliers (oracle/weights.py variant "outlier": MLP hidden activations ~1e4, one residual
# channel carrying several hundred): the regime where fp16 (range) and bf16 (mantissa) operands differ in kind
CASES += [
    dict(name="forward_r8_deep_outlier", kind="forward", bins="r8_t4_nwpu", deep_vpt=True, num_vpt=32,
         variant="outlier", wseed=30, xseed=31, shape=(4, 3, 224, 224)),
]


# ViT-L/14 (width 1024, 24 layers, 16 heads, patch 14, embed 768, decoder 1024 channels, x1.75 resample to the
# reduction-8 grid; 257 live tokens + 32 prompts = 289 keys per 224 x 224 window: the streamed-K/V attention kernel)
CASES += [
    dict(name="l14_forward_r8_deep", kind="forward", bins="r8_t4_nwpu", deep_vpt=True, num_vpt=32,
         variant="stress", wseed=24, xseed=25, shape=(1, 3, 224, 224), patch=14),
    dict(name="l14_sliding_224x448_s224_r8_shallow", kind="sliding", bins="r8_t4_nwpu", deep_vpt=False, num_vpt=32,
         variant="default", wseed=26, xseed=27, shape=(1, 3, 224, 448), window=224, stride=224, patch=14),
]

# the CLIP-ResNet encoders behind the same boundary (_clip/image_encoder.py:10-115; Bottleneck decoder,
# models/clip/model.py:228-239). resnet50 at reduction 8 (layer4 stride 1, x2 resample), reduction 32 (layer4 stride 2, no
# resample); resnet101 at reduction 16 (two decoder blocks, the second with a downsample conv); overlapping sliding windows.
CASES += [
    dict(name="rn50_forward_r8", kind="forward", bins="r8_t4_nwpu", backbone="resnet50", variant="stress", wseed=40, xseed=41,
         shape=(2, 3, 224, 224), num_vpt=0, deep_vpt=False),
    dict(name="rn50_forward_r32", kind="forward", bins="r32_t19_qnrf", backbone="resnet50", variant="stress", wseed=42, xseed=43,
         shape=(3, 3, 224, 224), num_vpt=0, deep_vpt=False),
    dict(name="rn101_forward_r16", kind="forward", bins="r16_t8_qnrf", backbone="resnet101", variant="stress", wseed=44, xseed=45,
         shape=(2, 3, 224, 224), num_vpt=0, deep_vpt=False),
    dict(name="rn50_sliding_448x672_s112_r8", kind="sliding", bins="r8_t4_nwpu", backbone="resnet50", variant="stress", wseed=46,
         xseed=47, shape=(1, 3, 448, 672), window=224, stride=112, num_vpt=0, deep_vpt=False),
    dict(name="rn50_sliding_300x500_s200_r8_default", kind="sliding", bins="r8_t4_nwpu", backbone="resnet50", variant="default",
         wseed=48, xseed=49, shape=(1, 3, 300, 500), window=224, stride=200, num_vpt=0, deep_vpt=False),
    # the wider CLIP-ResNets: stem widths 80 / 96 / 128 (channel counts that are not multiples of 64 -> zero-padded operands),
    # embed_dim 640 (padded to 768 for the head epilogue) / 768 / 1024
    dict(name="rn50x4_forward_r8", kind="forward", bins="r8_t4_nwpu", backbone="resnet50x4", variant="stress", wseed=59, xseed=51,
         shape=(2, 3, 224, 224), num_vpt=0, deep_vpt=False),
    dict(name="rn50x16_forward_r16", kind="forward", bins="r16_t8_qnrf", backbone="resnet50x16", variant="stress", wseed=52, xseed=53,
         shape=(1, 3, 224, 288), num_vpt=0, deep_vpt=False),
    dict(name="rn50x64_forward_r8", kind="forward", bins="r8_t4_nwpu", backbone="resnet50x64", variant="stress", wseed=54, xseed=55,
         shape=(1, 3, 224, 224), num_vpt=0, deep_vpt=False),
]

# cases pinned for the oracle only (no CUDA implementation yet): none
ORACLE_ONLY_CASES = []


def backbone_of(case: dict) -> str:
    if "backbone" in case:
        return case["backbone"]
    return {32: "vit_b_32", 14: "vit_l_14"}.get(case.get("patch", 16), "vit_b_16")


def case_inputs(case: dict):
    """-> (state_dict, text_features, bins, anchors, reduction, x) regenerated from the case's seeds."""
    reduction, bins, anchors = weights.bins_and_anchors(case["bins"])
    if case.get("backbone", "").startswith("resnet"):
        sd = weights.make_resnet_state_dict(case["wseed"], case["backbone"], case["variant"])
        tf = weights.make_text_features(len(bins), seed=100 + case["wseed"], embed=weights.RESNETS[case["backbone"]]["embed"])
        return sd, tf, bins, anchors, reduction, weights.make_image(case["shape"], seed=case["xseed"])
    sd = weights.make_state_dict(case["wseed"], input_size=224, num_vpt=case["num_vpt"], deep_vpt=case["deep_vpt"],
                                 variant=case["variant"], patch=case.get("patch", 16))
    tf = weights.make_text_features(len(bins), seed=100 + case["wseed"], embed=768 if case.get("patch", 16) == 14 else 512)
    x = weights.make_image(case["shape"], seed=case["xseed"])
    return sd, tf, bins, anchors, reduction, x


# ---------------------------------------------------------------------------------------------------------------
# the steps either side of the hot path (oracle/eval_oracle.py, clip_ebc_b200/{transforms,
# eval_utils,eval_loop}.py). Fixtures tests/golden/eval_*.npz come from the reference's own functions.
# ---------------------------------------------------------------------------------------------------------------
RESIZE_DENSITY_CASES = [
    # (name, seed, input shape [B,1,h,w], target size (H,W), zero_image: index of an all-zero map or None)
    # the reference function only broadcasts for [1,1,h,w] inputs (x * scale_factor with scale_factor [B,C]), which is
    # what its callers pass (notebooks/model.ipynb): one density map per call
    dict(name="eval_resize_density_up8", seed=21, shape=(1, 1, 24, 32), size=(192, 256), zero_image=None),
    dict(name="eval_resize_density_zero_map", seed=22, shape=(1, 1, 28, 28), size=(224, 224), zero_image=0),
    dict(name="eval_resize_density_odd", seed=23, shape=(1, 1, 37, 53), size=(300, 421), zero_image=None),
    dict(name="eval_resize_density_down", seed=24, shape=(1, 1, 56, 84), size=(30, 41), zero_image=None),
]

TRANSFORM_CASES = [
    # uint8 image [3,H,W] from a seed; window / stride; subsample steps for the stored outputs
    dict(name="eval_transform_500x731_s112", seed=31, shape=(3, 500, 731), window=224, stride=112, n_points=40),
    dict(name="eval_transform_300x350_s224", seed=32, shape=(3, 300, 350), window=224, stride=224, n_points=7),
    dict(name="eval_transform_200x260_s112", seed=33, shape=(3, 200, 260), window=224, stride=112, n_points=0),  # upscale to the window
    dict(name="eval_transform_448x672_s112", seed=34, shape=(3, 448, 672), window=224, stride=112, n_points=3),  # already a multiple
]
SUB_Y, SUB_X = 7, 5  # stored outputs are out[:, ::SUB_Y, ::SUB_X] plus per-channel float64 sums


def make_u8_image(shape, seed):
    """Smooth-ish uint8 test image (low-frequency pattern + noise) so that resampling has something to resample."""
    import numpy as np

    rng = np.random.Generator(np.random.PCG64(seed))
    c, h, w = shape
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    img = np.empty(shape, dtype=np.float64)
    for ch in range(c):
        fy, fx, ph = rng.uniform(0.01, 0.08), rng.uniform(0.01, 0.08), rng.uniform(0, 6.28)
        img[ch] = 127.5 + 90.0 * np.sin(fy * yy + fx * xx + ph) + rng.normal(0, 20.0, size=(h, w))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def make_points(n, h, w, seed):
    import numpy as np
    import torch

    rng = np.random.Generator(np.random.PCG64(seed))
    if n == 0:
        return torch.zeros((0, 2), dtype=torch.float32)
    pts = np.stack([rng.uniform(0, w - 1, n), rng.uniform(0, h - 1, n)], axis=1)
    return torch.from_numpy(pts.astype(np.float32))


def make_density(shape, seed, zero_image=None):
    import numpy as np
    import torch

    rng = np.random.Generator(np.random.PCG64(seed))
    x = np.abs(rng.normal(0.0, 0.5, size=shape)).astype(np.float32)
    if zero_image is not None:
        x[zero_image] = 0.0
    return torch.from_numpy(x)
