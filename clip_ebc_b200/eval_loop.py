"""Evaluation-loop plumbing around the hot path: host-side mirror of
``evaluate()`` (/root/reference/eval.py:11-40) and of the NWPU test loop / result file (test_nwpu.py:89-116).

Same signature and return value as the reference. What differs is where the work happens: every image goes through one
C-ABI call (unfold, ViT, decoder, head, fold and the per-image count all on the device) and the per-image counts stay
on the GPU until the loop ends, so there is no host synchronisation per image (the reference does `.cpu()` on every
count, eval.py:35): the host only enqueues, and the H2D copy of image i+1 runs on a side stream while the kernels of image
i run (`_prefetch_to_device`).
"""
from __future__ import annotations

import collections
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from .eval_utils import calculate_errors, sliding_window_predict, sliding_window_predict_batch
from .model import CLIP_EBC

# Cross-image window batching: images are collected until their windows fill at least one internal pass (148 windows)
# and then go through ONE C-ABI call (clipebc_sliding_window_predict_batch). Images that fill a pass on their own
# (NWPU / QNRF scale) are unaffected; small ones (a dozen windows each) no longer leave most SMs idle.
BATCH_WINDOWS = 148


def _n_windows(h: int, w: int, window, stride) -> int:
    wh, ww = (window, window) if isinstance(window, (int, float)) else window
    sh, sw = (stride, stride) if isinstance(stride, (int, float)) else stride
    return (int(np.ceil((h - wh) / sh) + 1)) * (int(np.ceil((w - ww) / sw) + 1))


class _WindowBatcher:
    """Collects same-model images and flushes them through sliding_window_predict_batch; counts come back in order."""

    def __init__(self, model, window_size, stride, target_windows=BATCH_WINDOWS):
        self.model, self.window_size, self.stride, self.target = model, window_size, stride, target_windows
        self.pending, self.pending_windows, self.out = [], 0, []

    def add(self, image: Tensor) -> None:
        for b in range(image.shape[0]):
            im = image[b:b + 1]
            self.pending_windows += _n_windows(im.shape[-2], im.shape[-1], self.window_size, self.stride)
            if self.pending_windows >= self.target:
                self.pending.append(im)
                self.flush()
            else:
                # held until later images fill the pass: the caller's tensor may be a slot of the prefetch ring that is
                # overwritten a few images from now, so a small image that waits is copied (large ones never wait)
                self.pending.append(im.clone())

    def flush(self) -> None:
        if self.pending:
            _, cnt = sliding_window_predict_batch(self.model, self.pending, self.window_size, self.stride, return_device=True)
            self.out.append(cnt)
            self.pending, self.pending_windows = [], 0

    def counts(self) -> Tensor:
        self.flush()
        return torch.cat(self.out) if self.out else torch.empty(0)


def _prefetch_to_device(items: Iterable, device: torch.device, image_of: Callable = lambda item: item,
                        depth: int = 2) -> Iterator[Tuple[Tensor, object]]:
    """Yields (image on `device`, item) for every item, with the host->device copies of the next `depth` images already
    issued on a side stream: the copy of image i+1 (38 MB for 2048x1536, 151 MB for 4096x3072) runs while the caller's stream
    computes image i, instead of in front of it on the same stream. The copies land in a ring of depth + 1 persistent device
    buffers (grown to the largest image seen), so the steady state allocates nothing; a slot is overwritten only after the
    work the caller enqueued on its previous image has completed. Images already on the device pass through.

    The yielded tensor is valid until the generator is advanced again (the caller consumes one image per iteration)."""
    copy_stream = torch.cuda.Stream(device)
    compute = torch.cuda.current_stream(device)
    n_slots = max(1, depth) + 1
    slots: List[Optional[Tensor]] = [None] * n_slots
    reusable: List[Optional[torch.cuda.Event]] = [None] * n_slots
    queue: collections.deque = collections.deque()
    it = iter(items)
    state = {"next_slot": 0}

    def issue() -> None:
        try:
            item = next(it)
        except StopIteration:
            return
        image = image_of(item)
        if image.device == device:
            queue.append((image, None, item, None))
            return
        slot = state["next_slot"]
        state["next_slot"] = (slot + 1) % n_slots
        n = image.numel()
        with torch.cuda.stream(copy_stream):
            if reusable[slot] is not None:
                copy_stream.wait_event(reusable[slot])
            buf = slots[slot]
            if buf is None or buf.numel() < n or buf.dtype != image.dtype:
                buf = torch.empty(n, dtype=image.dtype, device=device)
                buf.record_stream(compute)  # allocated under the copy stream, read (and outlived) by the compute stream
                slots[slot] = buf
            dev_image = buf[:n].view(image.shape)
            dev_image.copy_(image, non_blocking=True)
            done = torch.cuda.Event()
            done.record(copy_stream)
        queue.append((dev_image, done, item, slot))  # `item` keeps the (pinned) host tensor alive until the copy was consumed

    for _ in range(max(1, depth)):
        issue()
    while queue:
        dev_image, done, item, slot = queue.popleft()
        if done is not None:
            compute.wait_event(done)
        issue()
        yield dev_image, item
        if slot is not None:  # everything the caller does with this image has been enqueued by now
            ev = torch.cuda.Event()
            ev.record(compute)
            reusable[slot] = ev


def _device_counts(model: CLIP_EBC, image: Tensor, sliding_window: bool, window_size, stride) -> Tensor:
    """counts [B] (device) of a batch of same-sized images [B,3,H,W] already on the model's device."""
    if sliding_window:
        outs = []
        for b in range(image.shape[0]):  # the reference asserts nothing here, but its fold only handles one image
            _, cnt = sliding_window_predict(model, image[b:b + 1], window_size, stride, return_device=True, return_count=True)
            outs.append(cnt.reshape(1))
        return torch.cat(outs)
    return model(image).sum(dim=(1, 2, 3))  # eval.py:33,35 on the device


def evaluate(
    model: nn.Module,
    data_loader: Iterable,
    device: torch.device,
    sliding_window: bool = False,
    window_size: Optional[int] = None,
    stride: Optional[int] = None,
) -> Dict[str, float]:
    """MAE / RMSE of the predicted counts over `data_loader` (batches of `(image, target_points, density)`)."""
    if not isinstance(model, CLIP_EBC):
        raise TypeError(f"clip_ebc_b200.evaluate drives clip_ebc_b200.CLIP_EBC models only (got {type(model).__name__})")
    model.eval()
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("clip_ebc_b200.evaluate needs a CUDA device (there is no CPU fallback)")
    pred_counts, target_counts = [], []
    if sliding_window:
        assert window_size is not None, f"Window size must be provided when sliding_window is True, but got {window_size}"
        assert stride is not None, f"Stride must be provided when sliding_window is True, but got {stride}"
    batcher = _WindowBatcher(model, window_size, stride) if sliding_window else None

    for image, (_, target_points, _) in _prefetch_to_device(data_loader, device, image_of=lambda batch: batch[0]):
        target_counts.append([len(p) for p in target_points])
        with torch.set_grad_enabled(False):
            if batcher is not None:
                batcher.add(image)
            else:
                pred_counts.append(_device_counts(model, image, sliding_window, window_size, stride))

    if batcher is not None:
        pred = batcher.counts().cpu().numpy().astype(np.float64)     # the only device->host transfer of the loop
    elif pred_counts:
        pred = torch.cat(pred_counts).cpu().numpy().astype(np.float64)
    else:
        pred = np.array([])
    target = np.array([item for sublist in target_counts for item in sublist])
    assert len(pred) == len(target), f"Length of predictions and ground truths should be equal, but got {len(pred)} and {len(target)}"
    return calculate_errors(pred, target)


def predict_counts(model: nn.Module, images: Iterable[Tensor], device: torch.device, sliding_window: bool = False,
                   window_size: Optional[int] = None, stride: Optional[int] = None) -> List[float]:
    """The loop of test_nwpu.py:89-106: one predicted count per image ([3,H,W] or [1,3,H,W] tensors), in order."""
    if not isinstance(model, CLIP_EBC):
        raise TypeError(f"clip_ebc_b200.predict_counts drives clip_ebc_b200.CLIP_EBC models only (got {type(model).__name__})")
    model.eval()
    device = torch.device(device)
    outs = []
    batcher = _WindowBatcher(model, window_size, stride) if sliding_window else None
    for image, _ in _prefetch_to_device(images, device):
        image = image.unsqueeze(0) if image.dim() == 3 else image
        with torch.set_grad_enabled(False):
            if batcher is not None:
                batcher.add(image)
            else:
                outs.append(_device_counts(model, image, sliding_window, window_size, stride))
    if batcher is not None:
        return batcher.counts().cpu().tolist()
    return torch.cat(outs).cpu().tolist() if outs else []


def nwpu_result_text(image_ids: Sequence[str], preds: Sequence[float]) -> str:
    """The submission format of test_nwpu.py:111-116: '<id> <count>' per line, no newline at the end of the file."""
    assert len(image_ids) == len(preds), f"{len(image_ids)} ids but {len(preds)} predictions"
    return "\n".join(f"{image_id} {pred}" for image_id, pred in zip(image_ids, preds))


def write_nwpu_results(path: str, image_ids: Sequence[str], preds: Sequence[float]) -> None:
    with open(path, "w") as f:
        f.write(nwpu_result_text(image_ids, preds))
