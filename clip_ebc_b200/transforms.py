"""Pre-step of the hot path: host-side mirror of the two evaluation transforms of the reference and of the
normalisation its datasets apply.

  Resize2Multiple   /root/reference/datasets/transforms.py:69-105   (TF.resize BICUBIC antialias=True, :27-35)
  ZeroPad2Multiple  /root/reference/datasets/transforms.py:108-140  (TF.pad right/bottom, fill 0)
  Normalize         /root/reference/datasets/crowd.py:64,226        (torchvision Normalize, ImageNet statistics)

Same constructor arguments, assertions and `(image, label) -> (image, label)` call convention as the reference. The
image work runs in the sm_100a kernels of csrc/preproc.cu; images must be CUDA tensors ([C,H,W] float32 in [0,1] as
after ToTensor, or uint8, which is divided by 255 on the fly) -- there is no CPU path. Labels (point lists, a few
hundred floats) stay wherever they are and follow the resize arithmetic of the reference.

`preprocess()` is the fused form the evaluation loop uses: uint8 -> [0,1] -> resize | pad -> normalise in one or two
kernels without materialising the intermediate float image.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import ops

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _check_pairs(window_size, stride):
    window_size = (int(window_size), int(window_size)) if isinstance(window_size, (int, float)) else window_size
    window_size = tuple(window_size)
    stride = (int(stride), int(stride)) if isinstance(stride, (int, float)) else stride
    stride = tuple(stride)
    assert len(window_size) == 2, f"window_size should be a tuple (h, w), got {window_size}."
    assert len(stride) == 2, f"stride should be a tuple (h, w), got {stride}."
    assert all(s > 0 for s in window_size), f"window_size should be positive, got {window_size}."
    assert all(s > 0 for s in stride), f"stride should be positive, got {stride}."
    assert stride[0] <= window_size[0] and stride[1] <= window_size[1], \
        f"stride should be no larger than window_size, got {stride} and {window_size}."
    return window_size, stride


def _require_cuda_image(image: Tensor) -> None:
    if not isinstance(image, Tensor) or not image.is_cuda:
        raise RuntimeError("clip_ebc_b200.transforms run on the GPU: pass a CUDA tensor (there is no CPU fallback)")
    if image.dim() != 3:
        raise RuntimeError(f"expected a [C, H, W] image, got {tuple(image.shape)}")


def _resize_labels(label: Tensor, image_height: int, image_width: int, height: int, width: int) -> Tensor:
    # datasets/transforms.py:36-41
    if len(label) > 0 and (image_height != height or image_width != width):
        label[:, 0] = label[:, 0] * width / image_width
        label[:, 1] = label[:, 1] * height / image_height
        label[:, 0] = label[:, 0].clamp(min=0, max=width - 1)
        label[:, 1] = label[:, 1].clamp(min=0, max=height - 1)
    return label


class Resize2Multiple(object):
    """
    Resize the image so that it satisfies:
        img_h = window_h + stride_h * n_h
        img_w = window_w + stride_w * n_w
    """

    def __init__(self, window_size: Tuple[int, int], stride: Tuple[int, int]) -> None:
        self.window_size, self.stride = _check_pairs(window_size, stride)

    def new_size(self, image_height: int, image_width: int) -> Tuple[int, int]:
        window_height, window_width = self.window_size
        stride_height, stride_width = self.stride
        new_height = int(max(round((image_height - window_height) / stride_height), 0) * stride_height + window_height)
        new_width = int(max(round((image_width - window_width) / stride_width), 0) * stride_width + window_width)
        return new_height, new_width

    def __call__(self, image: Tensor, label: Tensor) -> Tuple[Tensor, Tensor]:
        _require_cuda_image(image)
        image_height, image_width = image.shape[-2:]
        new_height, new_width = self.new_size(image_height, image_width)
        if new_height == image_height and new_width == image_width:
            return image, label
        return ops.resize_bicubic_aa(image, new_height, new_width), _resize_labels(label, image_height, image_width,
                                                                                 new_height, new_width)


class ZeroPad2Multiple(object):
    def __init__(self, window_size: Tuple[int, int], stride: Tuple[int, int]) -> None:
        self.window_size, self.stride = _check_pairs(window_size, stride)

    def new_size(self, image_height: int, image_width: int) -> Tuple[int, int]:
        window_height, window_width = self.window_size
        stride_height, stride_width = self.stride
        new_height = int(max(np.ceil((image_height - window_height) / stride_height), 0) * stride_height + window_height)
        new_width = int(max(np.ceil((image_width - window_width) / stride_width), 0) * stride_width + window_width)
        return new_height, new_width

    def __call__(self, image: Tensor, label: Tensor) -> Tuple[Tensor, Tensor]:
        _require_cuda_image(image)
        image_height, image_width = image.shape[-2:]
        new_height, new_width = self.new_size(image_height, image_width)
        if new_height == image_height and new_width == image_width:
            return image, label
        assert new_height >= image_height and new_width >= image_width, \
            f"new size should be no less than the original size, got {new_height} and {new_width}."
        # only the right and bottom sides are padded so that the label coordinates are not affected
        return ops.pad_normalize(image, new_height, new_width), label


def normalize(image: Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> Tensor:
    """torchvision Normalize(mean, std) of a [C,H,W] image (uint8 input is first divided by 255)."""
    _require_cuda_image(image)
    return ops.pad_normalize(image, image.shape[-2], image.shape[-1], mean, std)


def preprocess(image: Tensor, transforms: Optional[object] = None, mean=IMAGENET_MEAN, std=IMAGENET_STD,
               label: Optional[Tensor] = None):
    """datasets/crowd.py:213-228 in fused form: uint8 (or [0,1] float) [C,H,W] -> `/ 255.` -> transforms -> Normalize.

    `transforms` is None, a Resize2Multiple or a ZeroPad2Multiple. Returns the normalised image, or (image, label) when a
    label tensor is given.
    """
    _require_cuda_image(image)
    h, w = int(image.shape[-2]), int(image.shape[-1])
    if transforms is None:
        out = ops.pad_normalize(image, h, w, mean, std)
    elif isinstance(transforms, Resize2Multiple):
        nh, nw = transforms.new_size(h, w)
        if (nh, nw) == (h, w):
            out = ops.pad_normalize(image, h, w, mean, std)
        else:
            out = ops.resize_bicubic_aa(image, nh, nw, mean, std)
            if label is not None:
                label = _resize_labels(label, h, w, nh, nw)
    elif isinstance(transforms, ZeroPad2Multiple):
        nh, nw = transforms.new_size(h, w)
        assert nh >= h and nw >= w, f"new size should be no less than the original size, got {nh} and {nw}."
        out = ops.pad_normalize(image, nh, nw, mean, std)
    else:
        raise TypeError(f"unsupported transform {type(transforms).__name__}: expected Resize2Multiple, ZeroPad2Multiple or None")
    return out if label is None else (out, label)
