"""TEST INFRASTRUCTURE -- not product code. Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
(`cpu_baseline`, `--impl reference`) may import this module; the product path (clip_ebc_b200/) never does.

CPU fp32 restatement of the CLIP-EBC inference hot path, written from the reference sources as the reference
*executes* it (nominal 229-token sequence, cats and all), each function citing the file:line it follows. It is the
checker for the CUDA path and the "port" CPU baseline that bench.py times.

Parity pinning: the reference has no tests or golden vectors for this path, so the oracle is pinned
against outputs of the reference itself, generated in the build container by oracle/make_golden.py (which imports the
real /root/reference classes) and committed under tests/golden/. tests/test_oracle.py checks the oracle against every
fixture on CPU; tests/test_oracle_vs_reference.py additionally runs the live reference when /root/reference exists.
All third-party arithmetic is PyTorch (reference pins torch==2.2.1, requirements.txt:11; this image has 2.11.0).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

def dims(sd: Dict[str, Tensor]) -> Tuple[int, int, int]:
    """(width, layers, heads) of the vision transformer, read off the state_dict: ViT-B (768, 12, 12) or ViT-L/14
    (1024, 24, 16); heads = width // 64 (_clip/model.py:50)."""
    width = int(sd["image_encoder.conv1.weight"].shape[0])
    layers = 0
    while f"image_encoder.transformer.resblocks.{layers}.ln_1.weight" in sd:
        layers += 1
    return width, layers, width // 64


def patch_size(sd: Dict[str, Tensor]) -> int:
    """ViT patch size = encoder reduction (_clip/image_encoder.py:141, models/clip/model.py:78): 16 or 32, read off the
    patch-embedding kernel."""
    return int(sd["image_encoder.conv1.weight"].shape[-1])


def _ln(x: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    # _clip/blocks.py:8-14 -- nn.LayerNorm(768), eps 1e-5, computed in fp32
    return F.layer_norm(x.float(), (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def interpolate_pos_embed(pos: Tensor, g0: int, h: int, w: int) -> Tensor:
    # _clip/image_encoder.py:183-198 -- returned as-is when the grid matches, else bicubic resize of the patch part
    if h == g0 and w == g0:
        return pos
    grid = pos[1:].reshape(g0, g0, -1).permute(2, 0, 1).unsqueeze(0)
    grid = F.interpolate(grid, size=(h, w), mode="bicubic").squeeze(0)
    return torch.cat([pos[:1], grid.permute(1, 2, 0).reshape(h * w, -1)], dim=0)


def residual_attention_block(z: Tensor, sd: Dict[str, Tensor], l: int) -> Tensor:
    """_clip/blocks.py:22-42. z: [L, B, width] (sequence first, as nn.MultiheadAttention(batch_first=False) sees it)."""
    p = f"image_encoder.transformer.resblocks.{l}."
    L, B, D = z.shape
    HEADS = D // 64
    a = _ln(z, sd, p + "ln_1")
    qkv = F.linear(a, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])  # rows of W = [Wq; Wk; Wv]
    q, k, v = qkv.chunk(3, dim=-1)

    def heads(t):  # [L, B, width] -> [B, heads, L, 64]; head h = channels 64h..64h+63
        return t.reshape(L, B, HEADS, D // HEADS).permute(1, 2, 0, 3)

    o = F.scaled_dot_product_attention(heads(q), heads(k), heads(v))  # scale 1/sqrt(64), no mask, no dropout
    o = o.permute(2, 0, 1, 3).reshape(L, B, D)
    z = z + F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])  # blocks.py:40
    hdn = F.linear(_ln(z, sd, p + "ln_2"), sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])
    hdn = hdn * torch.sigmoid(1.702 * hdn)  # QuickGELU, blocks.py:17-19
    return z + F.linear(hdn, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])  # blocks.py:41


def forward_vpt(x: Tensor, sd: Dict[str, Tensor], num_vpt: int, deep_vpt: bool, input_size: int = 224,
                taps: Optional[dict] = None) -> Tensor:
    """models/clip/model.py:142-189 (`_forward_vpt`). x: [B, 3, h, w] -> [B, 768, h/16, w/16]."""
    B, _, H, W = x.shape
    PATCH = patch_size(sd)
    WIDTH, LAYERS, _ = dims(sd)
    hp, wp = H // PATCH, W // PATCH
    f = F.conv2d(x, sd["image_encoder.conv1.weight"], stride=PATCH)  # :147, no bias
    f = f.reshape(B, WIDTH, -1).permute(0, 2, 1)  # :148-149
    cls = sd["image_encoder.class_embedding"] + torch.zeros(B, 1, WIDTH, device=x.device)
    f = torch.cat([cls, f], dim=1)  # :150-153
    f = f + interpolate_pos_embed(sd["image_encoder.positional_embedding"], input_size // PATCH, hp, wp)  # :155-156
    f = _ln(f, sd, "image_encoder.ln_pre").permute(1, 0, 2)  # :157-158 -> [1 + L, B, 768]
    if taps is not None:
        taps["ln_pre"] = f.permute(1, 0, 2).clone()
    vpt = sd["vpt_0"].unsqueeze(0).expand(B, -1, -1).permute(1, 0, 2)  # _prepare_vpt :131-140 (dropout is identity)
    for l in range(LAYERS):
        z = torch.cat([f[:1], vpt, f[1:]], dim=0)  # :164-168
        z = residual_attention_block(z, sd, l)  # :171
        if l < LAYERS - 1:  # :174-178
            vpt = sd[f"vpt_{l + 1}"].unsqueeze(0).expand(B, -1, -1).permute(1, 0, 2) if deep_vpt else z[1:num_vpt + 1]
        f = torch.cat([z[:1], z[num_vpt + 1:]], dim=0)  # :180-183
        if taps is not None and l in (0, LAYERS - 1):
            taps[f"block{l}"] = f.permute(1, 0, 2).clone()
    f = _ln(f.permute(1, 0, 2), sd, "image_encoder.ln_post")  # :185-186
    f = f[:, 1:, :].permute(0, 2, 1).reshape(B, WIDTH, hp, wp)  # :187-188
    return f


def basic_block(x: Tensor, sd: Dict[str, Tensor], prefix: str = "image_decoder.0.") -> Tensor:
    """models/utils.py:290-303 with BatchNorm2d in eval mode (running statistics, eps 1e-5)."""

    def bn(t, k):
        b = f"{prefix}bn{k}."
        return F.batch_norm(t, sd[b + "running_mean"], sd[b + "running_var"], sd[b + "weight"], sd[b + "bias"], False,
                            0.0, 1e-5)

    out = F.relu(bn(F.conv2d(x, sd[prefix + "conv1.weight"], padding=1), 1))
    out = bn(F.conv2d(out, sd[prefix + "conv2.weight"], padding=1), 2)
    return F.relu(out + x)


def is_resnet(sd: Dict[str, Tensor]) -> bool:
    """CLIP ModifiedResNet image encoder (resnet50 / resnet101 ...) rather than a VisionTransformer?"""
    return "image_encoder.layer1.0.conv1.weight" in sd


def _bn(t: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    # nn.BatchNorm2d in eval mode: running statistics, eps 1e-5
    return F.batch_norm(t, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], False, 0.0, 1e-5)


def clip_bottleneck(x: Tensor, sd: Dict[str, Tensor], p: str, stride: int) -> Tensor:
    """_clip/blocks.py:56-101 -- all convs have stride 1; an avgpool follows conv2 (and precedes the downsample conv) when
    stride > 1."""
    out = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"]), sd, p + "bn1"))
    out = F.relu(_bn(F.conv2d(out, sd[p + "conv2.weight"], padding=1), sd, p + "bn2"))
    if stride > 1:
        out = F.avg_pool2d(out, stride)
    out = _bn(F.conv2d(out, sd[p + "conv3.weight"]), sd, p + "bn3")
    identity = x
    if p + "downsample.0.weight" in sd:
        identity = F.avg_pool2d(x, stride) if stride > 1 else x  # nn.AvgPool2d(1) is the identity
        identity = _bn(F.conv2d(identity, sd[p + "downsample.0.weight"]), sd, p + "downsample.1")
    return F.relu(out + identity)


def resnet_encoder_reduction(reduction: int) -> int:
    # _clip/image_encoder.py:50,68: layer4 keeps stride 1 when reduction <= 16
    return 16 if reduction <= 16 else 32


def forward_resnet(x: Tensor, sd: Dict[str, Tensor], reduction: int, taps: Optional[dict] = None) -> Tensor:
    """_clip/image_encoder.py:77-115 (`ModifiedResNet._stem` / `forward`, features_only, out_indices=(-1,))."""
    e = "image_encoder."
    x = F.relu(_bn(F.conv2d(x, sd[e + "conv1.weight"], stride=2, padding=1), sd, e + "bn1"))
    x = F.relu(_bn(F.conv2d(x, sd[e + "conv2.weight"], padding=1), sd, e + "bn2"))
    x = F.relu(_bn(F.conv2d(x, sd[e + "conv3.weight"], padding=1), sd, e + "bn3"))
    x = F.avg_pool2d(x, 2)
    if taps is not None:
        taps["stem"] = x.clone()
    for layer in (1, 2, 3, 4):
        first_stride = 1 if layer == 1 or (layer == 4 and reduction <= 16) else 2
        i = 0
        while f"{e}layer{layer}.{i}.conv1.weight" in sd:
            x = clip_bottleneck(x, sd, f"{e}layer{layer}.{i}.", first_stride if i == 0 else 1)
            i += 1
        if taps is not None:
            taps[f"layer{layer}"] = x.clone()
    return x


def decoder_bottleneck(x: Tensor, sd: Dict[str, Tensor], p: str) -> Tensor:
    """models/utils.py:334-390 (`Bottleneck`, expansion 1, stride 1) -- the decoder block of the ResNet backbones."""
    out = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"]), sd, p + "bn1"))
    out = F.relu(_bn(F.conv2d(out, sd[p + "conv2.weight"], padding=1), sd, p + "bn2"))
    out = _bn(F.conv2d(out, sd[p + "conv3.weight"]), sd, p + "bn3")
    identity = x
    if p + "downsample.0.weight" in sd:
        identity = _bn(F.conv2d(x, sd[p + "downsample.0.weight"]), sd, p + "downsample.1")
    return F.relu(out + identity)


def clip_ebc_forward(x: Tensor, sd: Dict[str, Tensor], text_features: Tensor, anchor_points: Sequence[float],
                     reduction: int, num_vpt: int = 32, deep_vpt: bool = True, input_size: int = 224,
                     taps: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """models/clip/model.py:191-217 (`CLIP_EBC.forward`, ViT branch). Returns (logits [B,N,g,g], exp [B,1,g,g])."""
    with torch.no_grad():
        if is_resnet(sd):
            PATCH = resnet_encoder_reduction(reduction)
            f = forward_resnet(x.float(), sd, reduction, taps)  # :193
        else:
            PATCH = patch_size(sd)
            f = forward_vpt(x.float(), sd, num_vpt, deep_vpt, input_size, taps)
            if taps is not None:
                taps["ln_post"] = f.clone()
        if reduction != PATCH:
            f = F.interpolate(f, scale_factor=PATCH / reduction, mode="bilinear")  # :195-196
        if is_resnet(sd):
            j = 0
            while f"image_decoder.{j}.conv1.weight" in sd:  # make_resnet_layers(Bottleneck, decoder_cfg), model.py:83-87,228-239
                f = decoder_bottleneck(f, sd, f"image_decoder.{j}.")
                j += 1
        else:
            f = basic_block(f, sd)  # :197
        if taps is not None:
            taps["decoder"] = f.clone()
        f = F.conv2d(f, sd["projection.weight"], sd["projection.bias"])  # :198
        img = F.normalize(f.permute(0, 2, 3, 1), p=2, dim=-1)  # :200,203
        txt = F.normalize(text_features.float(), p=2, dim=-1)  # :204
        logits = sd["logit_scale"].exp() * img @ txt.t()  # :207-208
        logits = logits.permute(0, 3, 1, 2)  # :209
        probs = logits.softmax(dim=1)  # :211
        anchors = torch.tensor(list(anchor_points), dtype=torch.float32, device=x.device).view(1, -1, 1, 1)
        exp = (probs * anchors).sum(dim=1, keepdim=True)  # :212
        return logits, exp


def window_origins(H: int, W: int, window: Tuple[int, int], stride: Tuple[int, int]) -> Tuple[List[int], List[int]]:
    """utils/eval_utils.py:54-66 -- clamped window origins (rows, cols)."""
    wh, ww = window
    sh, sw = stride
    nr = int(np.ceil((H - wh) / sh) + 1)
    nc = int(np.ceil((W - ww) / sw) + 1)
    rows = [(i * sh if i * sh + wh <= H else H - wh) for i in range(nr)]
    cols = [(j * sw if j * sw + ww <= W else W - ww) for j in range(nc)]
    return rows, cols


def fold_average(preds: np.ndarray, H: int, W: int, window: Tuple[int, int], stride: Tuple[int, int],
                 reduction: int) -> np.ndarray:
    """utils/eval_utils.py:78-95 -- fp32 accumulate in window order, divide by the coverage count."""
    rows, cols = window_origins(H, W, window, stride)
    wh, ww = window
    pm = np.zeros((preds.shape[1], H // reduction, W // reduction), dtype=np.float32)
    cm = np.zeros_like(pm)
    idx = 0
    for xs in rows:
        for ys in cols:
            xe, ye = xs + wh, ys + ww
            pm[:, xs // reduction: xe // reduction, ys // reduction: ye // reduction] += preds[idx]
            cm[:, xs // reduction: xe // reduction, ys // reduction: ye // reduction] += 1.0
            idx += 1
    return pm / cm


def sliding_window_predict(image: Tensor, sd: Dict[str, Tensor], text_features: Tensor,
                           anchor_points: Sequence[float], reduction: int, window_size: Union[int, Tuple[int, int]],
                           stride: Union[int, Tuple[int, int]], num_vpt: int = 32, deep_vpt: bool = True,
                           input_size: int = 224, max_batch: Optional[int] = None) -> Tensor:
    """utils/eval_utils.py:26-96. image [1,3,H,W] -> [1,1,H//r,W//r] (CPU fp32).

    `max_batch` only chunks the model call (results are per-window independent); the reference runs one batch."""
    assert image.dim() == 4, f"Image must be a 4D tensor (1, c, h, w), got {image.shape}"
    window = (int(window_size),) * 2 if isinstance(window_size, (int, float)) else tuple(window_size)
    stride = (int(stride),) * 2 if isinstance(stride, (int, float)) else tuple(stride)
    assert stride[0] <= window[0] and stride[1] <= window[1]
    H, W = image.shape[-2:]
    rows, cols = window_origins(H, W, window, stride)
    wins = torch.cat([image[:, :, xs:xs + window[0], ys:ys + window[1]] for xs in rows for ys in cols], dim=0)
    outs = []
    step = max_batch or wins.shape[0]
    for i in range(0, wins.shape[0], step):
        outs.append(clip_ebc_forward(wins[i:i + step], sd, text_features, anchor_points, reduction, num_vpt, deep_vpt,
                                     input_size)[1])
    preds = torch.cat(outs, dim=0).numpy()
    return torch.tensor(fold_average(preds, H, W, window, stride, reduction)).unsqueeze(0)
