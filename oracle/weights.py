"""TEST INFRASTRUCTURE -- not product code. Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import it.

Seeded synthetic weights for CLIP-EBC (ViT-B/16 + VPT) under the reference's own ``state_dict`` key names, so the same
dictionary loads into (a) the real reference module (oracle/make_golden.py, in the build container), (b) the CPU oracle
(oracle/clip_ebc_oracle.py) and (c) the CUDA product (clip_ebc_b200.get_model(...).load_state_dict).

numpy's PCG64 stream is used instead of torch's RNG so that the weights are bit-identical on every machine (the GPU box
has no /root/reference and must regenerate exactly the tensors the golden fixtures were made with).

Distributions follow the reference constructors (``variant="default"``):
  image_encoder.*            /root/reference/models/clip/_clip/image_encoder.py:141-151, blocks.py:22-33 (torch defaults)
  vpt_*                      /root/reference/models/clip/model.py:70-76   U(+-sqrt(6/(3*16+768)))
  image_decoder / projection /root/reference/models/utils.py:366-379      kaiming-normal fan_out, BN (1, 0), stats (0, 1)
  logit_scale                /root/reference/models/clip/model.py:117     ln(1/0.07)
``variant="stress"`` additionally randomises every bias, LayerNorm/BatchNorm affine + running statistics and uses a
sharper logit_scale, so that bias / BN-fold / epsilon paths cannot cancel out in parity tests.
``variant="outlier"`` is "stress" plus trained-CLIP-like activation outliers: MLP hidden activations of ~1e4 in two
blocks and one residual channel that carries a few hundred from there on (see OUTLIER_* below).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

WIDTH, LAYERS, HEADS, PATCH, EMBED, HIDDEN = 768, 12, 12, 16, 512, 3072

# bins / anchor points of the reference configs (configs/reduction_{8,16,32}.json), "fine" granularity, "average" anchors
BIN_CONFIGS = {
    # reduction 8, truncation 4, nwpu  (configs/reduction_8.json["4"]["nwpu"])
    "r8_t4_nwpu": dict(reduction=8, bins=[(0, 0), (1, 1), (2, 2), (3, 3), (4, float("inf"))],
                       anchor_points=[0.0, 1.0, 2.0, 3.0, 4.21931]),
    # reduction 16, truncation key "8", qnrf (configs/reduction_16.json["8"]["qnrf"])
    "r16_t8_qnrf": dict(reduction=16, bins=[(i, i) for i in range(8)] + [(8, float("inf"))],
                        anchor_points=[0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 9.23349]),
    # reduction 32, truncation 19, qnrf (configs/reduction_32.json["19"]["qnrf"])
    "r32_t19_qnrf": dict(reduction=32, bins=[(i, i) for i in range(19)] + [(19, float("inf"))],
                         anchor_points=[float(i) for i in range(19)] + [23.01897]),
}


OUTLIER_LAYERS, OUTLIER_UNITS, OUTLIER_FC_SCALE, OUTLIER_PROJ_W = (2, 7), 6, 6000.0, 0.02


def _u(rng, shape, bound):
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def _n(rng, shape, std):
    return (rng.standard_normal(size=shape, dtype=np.float32) * np.float32(std)).astype(np.float32)


def make_state_dict(seed: int = 0, input_size: int = 224, num_vpt: int = 32, deep_vpt: bool = True,
                    variant: str = "default", patch: int = 16) -> Dict[str, torch.Tensor]:
    """Reference-keyed fp32 state_dict (without ``text_encoder.*``: the text tower is not on the hot path).
    patch = 16 (ViT-B/16) or 32 (ViT-B/32): the two differ in conv1 / positional-embedding shapes only."""
    assert variant in ("default", "stress", "outlier")
    PATCH = patch
    # ViT-B/16 and ViT-B/32: width 768, 12 layers, embed 512; ViT-L/14 (patch 14): width 1024, 24 layers, embed 768
    # (models/clip/model.py:16-24); hidden = 4 * width, heads = width // 64
    WIDTH, LAYERS, EMBED = (1024, 24, 768) if patch == 14 else (768, 12, 512)
    HIDDEN = 4 * WIDTH
    rng = np.random.default_rng(seed)
    stress = variant in ("stress", "outlier")
    sd: Dict[str, np.ndarray] = {}
    g0 = input_size // PATCH

    for l in range(LAYERS if deep_vpt else 1):
        sd[f"vpt_{l}"] = _u(rng, (num_vpt, WIDTH), math.sqrt(6.0 / (3 * PATCH + WIDTH)))
    sd["logit_scale"] = np.array(math.log(30.0) if stress else math.log(1 / 0.07), dtype=np.float32)

    ie = "image_encoder."
    sd[ie + "class_embedding"] = _n(rng, (WIDTH,), WIDTH ** -0.5)
    sd[ie + "positional_embedding"] = _n(rng, (g0 * g0 + 1, WIDTH), WIDTH ** -0.5)
    sd[ie + "conv1.weight"] = _u(rng, (WIDTH, 3, PATCH, PATCH), 1.0 / math.sqrt(3 * PATCH * PATCH))

    def ln(prefix):
        if stress:
            sd[prefix + ".weight"] = (1.0 + _n(rng, (WIDTH,), 0.1)).astype(np.float32)
            sd[prefix + ".bias"] = _n(rng, (WIDTH,), 0.1)
        else:
            sd[prefix + ".weight"] = np.ones((WIDTH,), np.float32)
            sd[prefix + ".bias"] = np.zeros((WIDTH,), np.float32)

    ln(ie + "ln_pre")
    for l in range(LAYERS):
        p = f"{ie}transformer.resblocks.{l}."
        sd[p + "attn.in_proj_weight"] = _u(rng, (3 * WIDTH, WIDTH), math.sqrt(6.0 / (WIDTH + 3 * WIDTH)))  # xavier
        sd[p + "attn.in_proj_bias"] = _n(rng, (3 * WIDTH,), 0.02) if stress else np.zeros((3 * WIDTH,), np.float32)
        sd[p + "attn.out_proj.weight"] = _u(rng, (WIDTH, WIDTH), 1.0 / math.sqrt(WIDTH))
        sd[p + "attn.out_proj.bias"] = _n(rng, (WIDTH,), 0.02) if stress else np.zeros((WIDTH,), np.float32)
        ln(p + "ln_1")
        sd[p + "mlp.c_fc.weight"] = _u(rng, (HIDDEN, WIDTH), 1.0 / math.sqrt(WIDTH))
        sd[p + "mlp.c_fc.bias"] = _u(rng, (HIDDEN,), 1.0 / math.sqrt(WIDTH))
        sd[p + "mlp.c_proj.weight"] = _u(rng, (WIDTH, HIDDEN), 1.0 / math.sqrt(HIDDEN))
        sd[p + "mlp.c_proj.bias"] = _u(rng, (WIDTH,), 1.0 / math.sqrt(HIDDEN))
        ln(p + "ln_2")
    ln(ie + "ln_post")

    for k in (1, 2):
        sd[f"image_decoder.0.conv{k}.weight"] = _n(rng, (WIDTH, WIDTH, 3, 3), math.sqrt(2.0 / (WIDTH * 9)))
        bn = f"image_decoder.0.bn{k}."
        if stress:
            sd[bn + "weight"] = rng.uniform(0.5, 1.5, size=(WIDTH,)).astype(np.float32)
            sd[bn + "bias"] = _n(rng, (WIDTH,), 0.1)
            sd[bn + "running_mean"] = _n(rng, (WIDTH,), 0.1)
            sd[bn + "running_var"] = rng.uniform(0.5, 1.5, size=(WIDTH,)).astype(np.float32)
        else:
            sd[bn + "weight"] = np.ones((WIDTH,), np.float32)
            sd[bn + "bias"] = np.zeros((WIDTH,), np.float32)
            sd[bn + "running_mean"] = np.zeros((WIDTH,), np.float32)
            sd[bn + "running_var"] = np.ones((WIDTH,), np.float32)
        sd[bn + "num_batches_tracked"] = np.array(0, dtype=np.int64)
    sd["projection.weight"] = _n(rng, (EMBED, WIDTH, 1, 1), math.sqrt(2.0 / EMBED))
    sd["projection.bias"] = _n(rng, (EMBED,), 0.1) if stress else np.zeros((EMBED,), np.float32)
    if variant == "outlier":
        # trained-CLIP-like activation outliers (applied after all draws, so the stream above is the "stress" one): in
        # OUTLIER_LAYERS a handful of MLP hidden units get c_fc rows scaled up until QuickGELU(c_fc) reaches ~1e4 (the top
        # of what a 16-bit operand with an fp16 exponent can hold is 65504), their c_proj columns are scaled down and routed
        # into ONE residual channel, which therefore carries a "massive activation" of a few hundred through every later
        # LayerNorm -- the regime in which fp16 and bf16 operands differ in kind (range vs mantissa)
        orng = np.random.default_rng(seed + 7919)
        for l in OUTLIER_LAYERS:
            if l >= LAYERS:
                continue
            p = f"{ie}transformer.resblocks.{l}."
            units = orng.choice(HIDDEN, size=OUTLIER_UNITS, replace=False)
            chan = int(orng.integers(0, WIDTH))
            sd[p + "mlp.c_fc.weight"][units, :] *= np.float32(OUTLIER_FC_SCALE)
            sd[p + "mlp.c_proj.weight"][:, units] *= np.float32(0.02)
            sd[p + "mlp.c_proj.weight"][chan, units] = np.float32(OUTLIER_PROJ_W)
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


# CLIP-ResNet image encoders (OpenAI checkpoints via prepare.py) and their decoder_cfg (models/clip/model.py:228-239)
RESNETS = {"resnet50": dict(layers=(3, 4, 6, 3), width=64, embed=1024, decoder=(2048,)),
           "resnet101": dict(layers=(3, 4, 23, 3), width=64, embed=512, decoder=(2048, 1024)),
           "resnet50x4": dict(layers=(4, 6, 10, 6), width=80, embed=640, decoder=(1280,)),
           "resnet50x16": dict(layers=(6, 8, 18, 8), width=96, embed=768, decoder=(1536,)),
           "resnet50x64": dict(layers=(3, 15, 36, 10), width=128, embed=1024, decoder=(2048,))}


def make_resnet_state_dict(seed: int = 0, backbone: str = "resnet50", variant: str = "stress") -> Dict[str, torch.Tensor]:
    """Reference-keyed fp32 state_dict of CLIP_EBC with a CLIP ModifiedResNet encoder (width 64) and its Bottleneck decoder
    (models/clip/model.py:50-93,228-239; _clip/image_encoder.py:33-58; _clip/blocks.py:56-86; models/utils.py:334-371).
    Conv weights: nn.Conv2d default (kaiming-uniform, a = sqrt 5 -> U(+-1/sqrt(fan_in))) in the encoder, kaiming-normal fan_out in
    the decoder / projection (`_init_weights`). "stress" randomises every BatchNorm's affine parameters and running statistics
    (with gains > 1 on the residual branches, so that they are not drowned by the identity path of 16 untrained blocks)."""
    assert variant in ("default", "stress")
    spec = RESNETS[backbone]
    rng = np.random.default_rng(seed)
    stress = variant == "stress"
    sd: Dict[str, np.ndarray] = {}
    sd["logit_scale"] = np.array(math.log(30.0) if stress else math.log(1 / 0.07), dtype=np.float32)

    def conv(name, o, i, k, decoder=False):
        if decoder:
            sd[name] = _n(rng, (o, i, k, k), math.sqrt(2.0 / (o * k * k)))
        else:
            sd[name] = _u(rng, (o, i, k, k), 1.0 / math.sqrt(i * k * k))

    def bn(prefix, c, gain=1.0):
        if stress:
            sd[prefix + ".weight"] = (rng.uniform(0.5, 1.5, size=(c,)) * gain).astype(np.float32)
            sd[prefix + ".bias"] = _n(rng, (c,), 0.1)
            sd[prefix + ".running_mean"] = _n(rng, (c,), 0.1)
            sd[prefix + ".running_var"] = rng.uniform(0.5, 1.5, size=(c,)).astype(np.float32)
        else:
            sd[prefix + ".weight"] = np.ones((c,), np.float32)
            sd[prefix + ".bias"] = np.zeros((c,), np.float32)
            sd[prefix + ".running_mean"] = np.zeros((c,), np.float32)
            sd[prefix + ".running_var"] = np.ones((c,), np.float32)
        sd[prefix + ".num_batches_tracked"] = np.array(0, dtype=np.int64)

    e = "image_encoder."
    # gain of the residual branches' BatchNorms: large enough for the branches to matter next to the identity path, small enough
    # for 16 (resnet50) / 33 (resnet101) untrained blocks not to leave the 16-bit range
    nb = sum(spec["layers"])
    g = (2.0 if nb <= 16 else 1.5 if nb <= 26 else 1.3 if nb <= 40 else 1.2) if stress else 1.0
    w = spec["width"]
    conv(e + "conv1.weight", w // 2, 3, 3); bn(e + "bn1", w // 2, g)
    conv(e + "conv2.weight", w // 2, w // 2, 3); bn(e + "bn2", w // 2, g)
    conv(e + "conv3.weight", w, w // 2, 3); bn(e + "bn3", w, g)
    inplanes = w
    for li, (planes, blocks) in enumerate(zip((w, 2 * w, 4 * w, 8 * w), spec["layers"]), start=1):
        for b in range(blocks):
            p = f"{e}layer{li}.{b}."
            stride = 2 if (b == 0 and li > 1) else 1   # tensor shapes do not depend on layer4's stride
            conv(p + "conv1.weight", planes, inplanes, 1); bn(p + "bn1", planes, g)
            conv(p + "conv2.weight", planes, planes, 3); bn(p + "bn2", planes, g)
            conv(p + "conv3.weight", planes * 4, planes, 1); bn(p + "bn3", planes * 4, g)
            if stride > 1 or inplanes != planes * 4:
                conv(p + "downsample.0.weight", planes * 4, inplanes, 1); bn(p + "downsample.1", planes * 4)
            inplanes = planes * 4
    c_in = inplanes
    for j, c_out in enumerate(spec["decoder"]):
        p = f"image_decoder.{j}."
        conv(p + "conv1.weight", c_out, c_in, 1, True); bn(p + "bn1", c_out)
        conv(p + "conv2.weight", c_out, c_out, 3, True); bn(p + "bn2", c_out)
        conv(p + "conv3.weight", c_out, c_out, 1, True); bn(p + "bn3", c_out)
        if c_in != c_out:
            conv(p + "downsample.0.weight", c_out, c_in, 1, True); bn(p + "downsample.1", c_out)
        c_in = c_out
    sd["projection.weight"] = _n(rng, (spec["embed"], c_in, 1, 1), math.sqrt(2.0 / spec["embed"]))
    sd["projection.bias"] = _n(rng, (spec["embed"],), 0.1) if stress else np.zeros((spec["embed"],), np.float32)
    return {k: torch.from_numpy(np.ascontiguousarray(v)) if v.ndim else torch.tensor(v) for k, v in sd.items()}


def make_text_features(n_bins: int, seed: int = 100, embed: int = EMBED) -> torch.Tensor:
    """Stand-in for ``text_encoder(prompts)`` ([N, 512]); a constant input of the hot path (model.py:127-129)."""
    rng = np.random.default_rng(seed)
    return torch.from_numpy(_n(rng, (n_bins, embed), 1.0))


def make_image(shape: Tuple[int, ...], seed: int = 1) -> torch.Tensor:
    """ImageNet-normalised pixels are ~unit variance (datasets/crowd.py:64): standard normal synthetic images."""
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.standard_normal(size=shape, dtype=np.float32))


def bins_and_anchors(name: str) -> Tuple[int, List[Tuple[float, float]], List[float]]:
    c = BIN_CONFIGS[name]
    return c["reduction"], [(float(a), float(b)) for a, b in c["bins"]], list(c["anchor_points"])
