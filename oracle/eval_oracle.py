"""TEST INFRASTRUCTURE -- CPU restatement (fp32 PyTorch / numpy) of the steps either side of the hot path
(SURVEY.md section 8f): what the reference does to an image before `sliding_window_predict` and to the density map
after it. Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product
(clip_ebc_b200/) never does. All `path:line` citations are into /root/reference.

Pinned by tests/golden/eval_*.npz, produced by running the reference's own functions (oracle/make_golden.py).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # datasets/crowd.py:64
IMAGENET_STD = (0.229, 0.224, 0.225)


def calculate_errors(pred_counts: np.ndarray, gt_counts: np.ndarray) -> Dict[str, float]:
    """utils/eval_utils.py:8-16."""
    assert isinstance(pred_counts, np.ndarray) and isinstance(gt_counts, np.ndarray)
    assert len(pred_counts) == len(gt_counts)
    return {"mae": np.mean(np.abs(pred_counts - gt_counts)), "rmse": np.sqrt(np.mean((pred_counts - gt_counts) ** 2))}


def resize_density_map(x: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """utils/eval_utils.py:19-23: bilinear resize times sum(resized) / sum(x), 0/0 or x/0 -> 0 (reference arithmetic, kept as is)."""
    x_sum = torch.sum(x, dim=(-1, -2))
    x = F.interpolate(x, size=size, mode="bilinear")
    scale = torch.nan_to_num(torch.sum(x, dim=(-1, -2)) / x_sum, nan=0.0, posinf=0.0, neginf=0.0)
    return x * scale


def _pair(v) -> Tuple[int, int]:
    return (int(v), int(v)) if isinstance(v, (int, float)) else tuple(int(t) for t in v)


def resize2multiple_size(h: int, w: int, window, stride) -> Tuple[int, int]:
    """datasets/transforms.py:95-100 (Python round() = banker's rounding, as in the reference)."""
    (wh, ww), (sh, sw) = _pair(window), _pair(stride)
    return (int(max(round((h - wh) / sh), 0) * sh + wh), int(max(round((w - ww) / sw), 0) * sw + ww))


def zeropad2multiple_size(h: int, w: int, window, stride) -> Tuple[int, int]:
    """datasets/transforms.py:129-133."""
    (wh, ww), (sh, sw) = _pair(window), _pair(stride)
    return (int(max(np.ceil((h - wh) / sh), 0) * sh + wh), int(max(np.ceil((w - ww) / sw), 0) * sw + ww))


def resize_image(image: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """datasets/transforms.py:27-35: TF.resize(BICUBIC, antialias=True) on a float [C,H,W] tensor is
    F.interpolate(mode="bicubic", align_corners=False, antialias=True) with no clamping (torchvision
    transforms/_functional_tensor.py resize; checked bit-identical in the build container)."""
    if image.shape[-2:] == (height, width):
        return image
    return F.interpolate(image[None], size=(height, width), mode="bicubic", align_corners=False, antialias=True)[0]


def resize_labels(label: torch.Tensor, h: int, w: int, height: int, width: int) -> torch.Tensor:
    """datasets/transforms.py:36-41: point labels (x, y) follow the resize and are clamped into the image."""
    label = label.clone()
    if len(label) > 0 and (h != height or w != width):
        label[:, 0] = (label[:, 0] * width / w).clamp(min=0, max=width - 1)
        label[:, 1] = (label[:, 1] * height / h).clamp(min=0, max=height - 1)
    return label


def zero_pad(image: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """datasets/transforms.py:138-140: pad right and bottom with 0 (before normalisation)."""
    h, w = image.shape[-2:]
    return F.pad(image, (0, width - w, 0, height - h), value=0.0)


def normalize(image: torch.Tensor) -> torch.Tensor:
    """datasets/crowd.py:64,226: torchvision Normalize(mean, std) = (x - mean) / std per channel."""
    mean = torch.tensor(IMAGENET_MEAN, dtype=image.dtype).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=image.dtype).view(3, 1, 1)
    return (image - mean) / std


def preprocess(image_u8: np.ndarray, mode: str, window, stride) -> torch.Tensor:
    """uint8 [3,H,W] -> the tensor the reference feeds to sliding_window_predict (datasets/crowd.py:213-228):
    /255, Resize2Multiple | ZeroPad2Multiple | nothing, then Normalize."""
    image = torch.from_numpy(image_u8).float() / 255.0
    h, w = image.shape[-2:]
    if mode == "resize":
        image = resize_image(image, *resize2multiple_size(h, w, window, stride))
    elif mode == "pad":
        image = zero_pad(image, *zeropad2multiple_size(h, w, window, stride))
    else:
        assert mode == "none"
    return normalize(image)


def evaluate(predict, images: Iterable[torch.Tensor], target_points: Sequence[Sequence]) -> Tuple[Dict[str, float], List[float]]:
    """eval.py:25-40: per image `pred_density.sum(dim=(1,2,3))`, then MAE / RMSE against len(points).
    `predict(image[1,3,H,W]) -> density[1,1,h,w]` is the (oracle) sliding-window predictor."""
    pred_counts, target_counts = [], []
    for image, pts in zip(images, target_points):
        dens = predict(image)
        pred_counts.append(dens.sum(dim=(1, 2, 3)).cpu().numpy().tolist())
        target_counts.append([len(pts)])
    pred = np.array([v for sub in pred_counts for v in sub])
    tgt = np.array([v for sub in target_counts for v in sub])
    return calculate_errors(pred, tgt), pred.tolist()


def nwpu_result_text(image_ids: Sequence[str], preds: Sequence[float]) -> str:
    """test_nwpu.py:111-116: '<id> <count>' per line, no newline after the last line."""
    return "\n".join(f"{i} {p}" for i, p in zip(image_ids, preds))
