"""Thin tensor-level wrappers over the single-kernel C-ABI entry points (tests, profiling, bench).

torch is used for device memory and streams only; all arithmetic happens in ``libclipebc_b200.so``.
Every wrapper requires CUDA tensors and raises otherwise -- there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _lib

EPI_F32, EPI_BIAS_F32, EPI_BIAS_BF16, EPI_BIAS_GELU_BF16, EPI_BIAS_RESID_F32, EPI_BIAS_RELU_MASK_BF16, \
    EPI_BIAS_RESID_RELU_SPLIT = range(7)
EPI_BIAS_RESID16_RELU_MASK_BF16 = 9  # `resid` is a 16-bit tensor in the output format


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("clip_ebc_b200 kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("clip_ebc_b200 kernels need contiguous tensors")
    return t.data_ptr()


def window_origins(H: int, W: int, window: Tuple[int, int], stride: Tuple[int, int]):
    """Host-only integer part of sliding_window_predict (reference utils/eval_utils.py:54-66)."""
    import ctypes as C

    lib = _lib.load()
    nr, nc = C.c_int(0), C.c_int(0)
    _lib.check(lib.clipebc_window_origins(H, W, window[0], window[1], stride[0], stride[1], C.byref(nr), C.byref(nc),
                                          None, None), "window_origins")
    ro, co = (C.c_int * nr.value)(), (C.c_int * nc.value)()
    _lib.check(lib.clipebc_window_origins(H, W, window[0], window[1], stride[0], stride[1], C.byref(nr), C.byref(nc),
                                          ro, co), "window_origins")
    return list(ro), list(co)


def _dt16(fp16: bool) -> torch.dtype:
    return torch.float16 if fp16 else torch.bfloat16


def to_16(x: torch.Tensor, fp16: bool = False) -> torch.Tensor:
    out = torch.empty(x.shape, dtype=_dt16(fp16), device=x.device)
    _lib.check(_lib.load().clipebc_f32_to_16(_ptr(x), _ptr(out), x.numel(), int(fp16), _stream()), "f32_to_16")
    return out


def gemm(a: torch.Tensor, w: torch.Tensor, epi: int, bias: Optional[torch.Tensor] = None,
         resid: Optional[torch.Tensor] = None, M: Optional[int] = None, K: Optional[int] = None,
         seg_row_shift: Optional[Sequence[int]] = None, seg_col_start: Optional[Sequence[int]] = None,
         mask_hw: Tuple[int, int] = (0, 0), mask_lead: bool = True, block_n: int = 0, out: Optional[torch.Tensor] = None,
         out_fp16: Optional[bool] = None) -> torch.Tensor:
    """D = A[M,K] @ W[N,K]^T with a fused epilogue. a, w: both bf16 or both fp16, 2-D contiguous."""
    assert a.dtype == w.dtype and a.dtype in (torch.bfloat16, torch.float16) and a.dim() == 2 and w.dim() == 2
    ab_fp16 = a.dtype == torch.float16
    out_fp16 = ab_fp16 if out_fp16 is None else out_fp16
    N = w.shape[0]
    K = w.shape[1] if K is None else K
    M = a.shape[0] if M is None else M
    n_seg = 1 if seg_row_shift is None and seg_col_start is None else len(seg_row_shift or seg_col_start)
    rs = _lib.int_array(seg_row_shift or [0] * n_seg)
    cs = _lib.int_array(seg_col_start or [0] * n_seg)
    if out is None:
        if epi in (EPI_F32, EPI_BIAS_F32, EPI_BIAS_RESID_F32):
            out = torch.empty((M, N), dtype=torch.float32, device=a.device)
        elif epi == EPI_BIAS_RESID_RELU_SPLIT:
            out = torch.empty((M, 2 * N), dtype=_dt16(out_fp16), device=a.device)
        else:
            out = torch.empty((M, N), dtype=_dt16(out_fp16), device=a.device)
    _lib.check(_lib.load().clipebc_gemm_bf16(
        epi, _ptr(a), a.shape[0], a.shape[1], a.stride(0), _ptr(w), w.stride(0), M, N, K, n_seg, rs, cs, _ptr(out),
        out.stride(0), _ptr(bias), _ptr(resid), resid.stride(0) if resid is not None else 0, mask_hw[0], mask_hw[1],
        int(mask_lead), block_n, int(ab_fp16), int(out_fp16), _stream()), "gemm")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out_dtype: torch.dtype = torch.bfloat16,
              n_rows_out: Optional[int] = None, rows_out_per_group: int = 1, rows_in_per_group: int = 1,
              in_row_offset: int = 0) -> torch.Tensor:
    """nn.LayerNorm over rows of 768 (ViT-B) or 1024 (ViT-L/14) channels."""
    width = int(x.shape[-1])
    assert x.dtype == torch.float32 and width in (768, 1024)
    n = x.numel() // width if n_rows_out is None else n_rows_out
    kind = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[out_dtype]
    out = torch.empty((n, width), dtype=out_dtype, device=x.device)
    _lib.check(_lib.load().clipebc_layernorm(_ptr(x), _ptr(gamma), _ptr(beta), width, _ptr(out), kind, n,
                                             rows_out_per_group, rows_in_per_group, in_row_offset, _stream()),
               "layernorm")
    return out


def attention(qkv: torch.Tensor, n_win: int, t_live: int, const_kv: Optional[torch.Tensor] = None,
              out_fp16: bool = False, heads: int = 12) -> torch.Tensor:
    width = 64 * heads
    assert qkv.dtype == torch.bfloat16 and qkv.shape == (n_win * t_live, 3 * width)
    n_const = 0 if const_kv is None else const_kv.shape[0]
    out = torch.empty((n_win * t_live, width), dtype=_dt16(out_fp16), device=qkv.device)
    _lib.check(_lib.load().clipebc_attention(_ptr(qkv), _ptr(const_kv), n_const, n_win, t_live, heads, _ptr(out),
                                             int(out_fp16), _stream()), "attention")
    return out


def patchify(image: torch.Tensor, y0: int = 0, x0: int = 0, gh: Optional[int] = None,
             gw: Optional[int] = None, fp16: bool = False, patch: int = 16, split: bool = True) -> torch.Tensor:
    """-> [n*gh*gw, 2*KP] (split) or [n*gh*gw, KP], KP = 3*patch^2 rounded up to 64: columns [0, KP) = hi, [KP, 2 KP) = lo of
    the hi/lo split of the pixels (pad columns zero)."""
    n, c, H, W = image.shape
    assert c == 3 and image.dtype == torch.float32
    gh = (H - y0) // patch if gh is None else gh
    gw = (W - x0) // patch if gw is None else gw
    kp = (3 * patch * patch + 63) // 64 * 64
    out = torch.zeros((n * gh * gw, (2 if split else 1) * kp), dtype=_dt16(fp16), device=image.device)
    _lib.check(_lib.load().clipebc_patchify(_ptr(image), n, H, W, y0, x0, gh, gw, int(patch), kp, int(split), _ptr(out),
                                            int(fp16), _stream()), "patchify")
    return out


def resample_to_padded(Y: torch.Tensor, n_win: int, hp: int, wp: int, gh: int, gw: int, fp16: bool = False):
    width = int(Y.shape[-1])
    ub = torch.empty((n_win * (gh + 1) * (gw + 1), width), dtype=_dt16(fp16), device=Y.device)
    uf = torch.empty((n_win * (gh + 1) * (gw + 1), width), dtype=torch.float32, device=Y.device)
    _lib.check(_lib.load().clipebc_resample_to_padded(_ptr(Y), n_win, hp, wp, gh, gw, width, _ptr(ub), _ptr(uf), int(fp16),
                                                      _stream()), "resample")
    return ub, uf


def fold_average(preds: torch.Tensor, row_cells: Sequence[int], col_cells: Sequence[int], Ho: int, Wo: int,
                 want_count: bool = False):
    """preds f32 [n_rows*n_cols, 1, gh, gw] (device) -> density [Ho, Wo] (device), optional count [1]."""
    gh, gw = preds.shape[-2:]
    assert preds.shape[0] == len(row_cells) * len(col_cells)
    dens = torch.empty((Ho, Wo), dtype=torch.float32, device=preds.device)
    cnt = torch.empty((1,), dtype=torch.float32, device=preds.device) if want_count else None
    _lib.check(_lib.load().clipebc_fold_average(_ptr(preds), _lib.int_array(row_cells), _lib.int_array(col_cells),
                                                len(row_cells), len(col_cells), gh, gw, Ho, Wo, _ptr(dens), _ptr(cnt),
                                                _stream()), "fold")
    return (dens, cnt) if want_count else dens


# ------------------------------------------------------------------------------------ pre / post steps (section 8f)
def _image_in(image: torch.Tensor):
    if image.dim() != 3 or image.shape[0] > 4:
        raise RuntimeError(f"expected a [C<=4, H, W] image, got {tuple(image.shape)}")
    if image.dtype not in (torch.uint8, torch.float32):
        raise RuntimeError(f"image must be uint8 or float32, got {image.dtype}")
    return _ptr(image), int(image.dtype == torch.uint8), int(image.shape[0]), int(image.shape[1]), int(image.shape[2])


def _stats(mean, std):
    if mean is None and std is None:
        return None, None
    return _lib.float_array(mean), _lib.float_array(std)


def resize_bicubic_aa(image: torch.Tensor, height: int, width: int, mean=None, std=None) -> torch.Tensor:
    """TF.resize(image, (height, width), BICUBIC, antialias=True) on a [C,H,W] uint8 (-> /255) or [0,1] float image,
    optionally followed by Normalize(mean, std) -- one pair of kernels (width pass, height pass)."""
    ptr, u8, c, h, w = _image_in(image)
    tmp = torch.empty((c, h, width), dtype=torch.float32, device=image.device)
    out = torch.empty((c, height, width), dtype=torch.float32, device=image.device)
    m, s = _stats(mean, std)
    _lib.check(_lib.load().clipebc_resize_bicubic_aa(ptr, u8, c, h, w, _ptr(tmp), _ptr(out), int(height), int(width), m, s,
                                                     _stream()), "resize_bicubic_aa")
    return out


def pad_normalize(image: torch.Tensor, height: int, width: int, mean=None, std=None) -> torch.Tensor:
    """Right/bottom zero padding of a [C,H,W] uint8 (-> /255) or [0,1] float image to (height, width), then Normalize."""
    ptr, u8, c, h, w = _image_in(image)
    out = torch.empty((c, int(height), int(width)), dtype=torch.float32, device=image.device)
    m, s = _stats(mean, std)
    _lib.check(_lib.load().clipebc_pad_normalize(ptr, u8, c, h, w, _ptr(out), int(height), int(width), m, s, _stream()),
               "pad_normalize")
    return out


def resize_density_map(x: torch.Tensor, size: Tuple[int, int], return_sums: bool = False):
    """One density map [h, w] -> [H, W]: bilinear resize times nan_to_num(sum(resized) / sum(x))."""
    assert x.dtype == torch.float32 and x.dim() == 2
    lib = _lib.load()
    ws = torch.empty(lib.clipebc_resize_density_workspace_floats(), dtype=torch.float32, device=x.device)
    out = torch.empty((int(size[0]), int(size[1])), dtype=torch.float32, device=x.device)
    sums = torch.empty(2, dtype=torch.float32, device=x.device) if return_sums else None
    _lib.check(lib.clipebc_resize_density_map(_ptr(x), int(x.shape[0]), int(x.shape[1]), int(size[0]), int(size[1]),
                                              _ptr(out), _ptr(ws), _ptr(sums), _stream()), "resize_density_map")
    return (out, sums) if return_sums else out
