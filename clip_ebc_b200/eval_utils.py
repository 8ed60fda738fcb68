"""Host-side mirror of /root/reference/utils/eval_utils.py: ``sliding_window_predict`` (:26-96), ``resize_density_map``
(:19-23) and ``calculate_errors`` (:8-16).

``sliding_window_predict``:

Same signature, same assertions (AssertionError with the reference's messages) and the same return value: a **CPU**
fp32 tensor of shape (1, 1, H // r, W // r). The body, however, is one C-ABI call: window enumeration, patch-grid
sharing between overlapping windows, the ViT/decoder/head kernels and the atomic-free fold all run on the device; the
only transfers are the image H2D (if the caller passed a CPU tensor) and the density map D2H.
"""
from __future__ import annotations

from typing import Dict, Tuple, Union

import numpy as np
import torch
from torch import Tensor, nn

from .model import CLIP_EBC


def _pair(v, what: str) -> Tuple[int, int]:
    v = (int(v), int(v)) if isinstance(v, (int, float)) else v
    v = tuple(v)
    assert isinstance(v, tuple) and len(v) == 2 and v[0] > 0 and v[1] > 0, \
        f"{what} must be a positive integer tuple (h, w), got {v}"
    return int(v[0]), int(v[1])


def sliding_window_predict(
    model: nn.Module,
    image: Tensor,
    window_size: Union[int, Tuple[int, int]],
    stride: Union[int, Tuple[int, int]],
    return_device: bool = False,
    return_count: bool = False,
):
    """Density map of one image by overlapping windows; overlapping regions are averaged.

    Args as in the reference. Extensions (default off, so the call is a drop-in):
      return_device: keep the result on the GPU instead of the reference's CPU tensor (throughput runs).
      return_count:  also return ``density.sum()`` computed on the device (what eval.py:35 / test_nwpu.py:100 do next).
    """
    assert len(image.shape) == 4, f"Image must be a 4D tensor (1, c, h, w), got {image.shape}"
    window_size = _pair(window_size, "Window size")
    stride = _pair(stride, "Stride")
    assert stride[0] <= window_size[0] and stride[1] <= window_size[1], \
        f"Stride must be smaller than window size, got {stride} and {window_size}"
    if not isinstance(model, CLIP_EBC):
        raise TypeError("clip_ebc_b200.sliding_window_predict drives clip_ebc_b200.CLIP_EBC models only "
                        f"(got {type(model).__name__}); there is no generic PyTorch fallback path")
    assert image.shape[0] == 1, f"The batch size must be 1 due to varying image sizes, got {image.shape[0]}"

    model.eval()  # eval_utils.py:73
    dev = model._device()
    with torch.no_grad():
        img = image if image.device == dev else image.to(dev, non_blocking=True)
        out = model.sliding_window_density(img, window_size, stride, with_count=return_count)
    dens, cnt = out if return_count else (out, None)
    if not return_device:
        dens = dens.cpu()  # the reference returns a CPU tensor (eval_utils.py:76,96)
        cnt = cnt.cpu() if cnt is not None else None
    return (dens, cnt) if return_count else dens


def sliding_window_predict_batch(model: nn.Module, images, window_size: Union[int, Tuple[int, int]],
                                 stride: Union[int, Tuple[int, int]], return_device: bool = False):
    """`sliding_window_predict` for a list of images of different sizes in ONE pass (extension, not in the reference:
    eval.py processes one image per iteration). The windows of all images are batched through the ViT / decoder / head
    together, every image is folded on its own; results are identical to per-image calls.

    Returns (densities, counts): a list of [1,1,H_i//r,W_i//r] maps and a [n] tensor of their sums -- CPU tensors like the
    reference unless `return_device`.
    """
    window_size = _pair(window_size, "Window size")
    stride = _pair(stride, "Stride")
    assert stride[0] <= window_size[0] and stride[1] <= window_size[1], \
        f"Stride must be smaller than window size, got {stride} and {window_size}"
    if not isinstance(model, CLIP_EBC):
        raise TypeError("clip_ebc_b200.sliding_window_predict_batch drives clip_ebc_b200.CLIP_EBC models only "
                        f"(got {type(model).__name__})")
    for image in images:
        assert len(image.shape) == 4, f"Image must be a 4D tensor (1, c, h, w), got {image.shape}"
        assert image.shape[0] == 1, f"The batch size must be 1 due to varying image sizes, got {image.shape[0]}"
    model.eval()
    dev = model._device()
    with torch.no_grad():
        imgs = [im if im.device == dev else im.to(dev, non_blocking=True) for im in images]
        dens, cnt = model.sliding_window_density_batch(imgs, window_size, stride)
    if not return_device:
        dens, cnt = [d.cpu() for d in dens], cnt.cpu()
    return dens, cnt


def calculate_errors(pred_counts: np.ndarray, gt_counts: np.ndarray) -> Dict[str, float]:
    """MAE / RMSE of per-image counts (reference utils/eval_utils.py:8-16; host arithmetic on a few hundred numbers)."""
    assert isinstance(pred_counts, np.ndarray), f"Expected numpy.ndarray, got {type(pred_counts)}"
    assert isinstance(gt_counts, np.ndarray), f"Expected numpy.ndarray, got {type(gt_counts)}"
    assert len(pred_counts) == len(gt_counts), \
        f"Length of predictions and ground truths should be equal, but got {len(pred_counts)} and {len(gt_counts)}"
    errors = {
        "mae": np.mean(np.abs(pred_counts - gt_counts)),
        "rmse": np.sqrt(np.mean((pred_counts - gt_counts) ** 2)),
    }
    return errors


def resize_density_map(x: Tensor, size: Tuple[int, int]) -> Tensor:
    """Bilinear resize of a density map times nan_to_num(sum(resized) / sum(x)) (reference utils/eval_utils.py:19-23; the
    reference multiplies by this ratio rather than its inverse -- kept bit for bit in meaning).

    The reference expression ``x * scale_factor`` (scale_factor of shape [B, C]) only broadcasts for one single-channel
    map, which is what its callers pass (notebooks/model.ipynb); other shapes raise here as they do there. A CPU input
    (what ``sliding_window_predict`` returns by default) is moved to the GPU, resized there (csrc/preproc.cu) and
    returned on its original device.
    """
    if x.dim() != 4 or x.shape[0] != 1 or x.shape[1] != 1:
        raise RuntimeError(f"The size of tensor a ({tuple(x.shape)}) must be (1, 1, h, w): the reference's "
                           "x * scale_factor does not broadcast for batched or multi-channel maps")
    if not torch.cuda.is_available():
        raise RuntimeError("clip_ebc_b200.resize_density_map needs a CUDA device (there is no CPU fallback)")
    from . import ops

    dev = x.device
    xd = x if x.is_cuda else x.cuda()
    out = ops.resize_density_map(xd.float().contiguous()[0, 0], (int(size[0]), int(size[1])))
    out = out[None, None]
    return out if dev.type == "cuda" else out.to(dev)
