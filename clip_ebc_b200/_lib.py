"""ctypes binding of the C-ABI in ``include/clipebc_b200.h`` (``libclipebc_b200.so``).

This is the only place the Python host side touches native code. There is no fallback: if the shared library is
missing the import of any compute entry point raises, and every non-zero return code becomes a ``RuntimeError``
carrying ``clipebc_last_error()`` (the reference raises ``AssertionError``/``RuntimeError`` from Python, see
INTEGRATION.md).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libclipebc_b200.so")

_lib: Optional[C.CDLL] = None


class ClipEbcConfig(C.Structure):
    """``clipebc_config`` of the header (ABI v9). Build it with :func:`make_config`, which fills ``struct_size``."""

    _fields_ = [
        ("struct_size", C.c_uint32),
        ("input_size", C.c_int),
        ("reduction", C.c_int),
        ("num_vpt", C.c_int),
        ("deep_vpt", C.c_int),
        ("num_bins", C.c_int),
        ("window_chunk", C.c_int),
        ("operand_fp16", C.c_int),
        ("patch", C.c_int),
        ("width", C.c_int),
        ("layers", C.c_int),
        ("embed_dim", C.c_int),
        ("decoder_conv1_fine", C.c_int),
        ("encoder", C.c_int),
    ]


def make_config(input_size: int, reduction: int, num_vpt: int, deep_vpt: int, num_bins: int, window_chunk: int = 0,
                operand_fp16: int = 1, patch: int = 16, width: int = 768, layers: int = 12, embed_dim: int = 512,
                decoder_conv1_fine: int = 0, encoder: int = 0) -> ClipEbcConfig:
    return ClipEbcConfig(C.sizeof(ClipEbcConfig), int(input_size), int(reduction), int(num_vpt), int(deep_vpt), int(num_bins),
                         int(window_chunk), int(operand_fp16), int(patch), int(width), int(layers), int(embed_dim),
                         int(decoder_conv1_fine), int(encoder))


_vp, _i, _i64, _fp = C.c_void_p, C.c_int, C.c_int64, C.c_void_p  # float* passed as raw addresses
_ip = C.POINTER(C.c_int)
_cfp = C.POINTER(C.c_float)  # HOST float arrays

# name -> (restype, argtypes): every symbol declared in include/clipebc_b200.h
SIGNATURES = {
    "clipebc_last_error": (C.c_char_p, []),
    "clipebc_abi_version": (_i, []),
    "clipebc_launch_count": (_i64, []),
    "clipebc_profile_enable": (_i, [_i]),
    "clipebc_profile_dump": (_i, [C.c_char_p, _i]),
    "clipebc_profile_enabled": (_i, []),
    "clipebc_config_epoch": (_i64, []),
    "clipebc_note_replayed_launches": (None, [_i64]),
    "clipebc_model_create": (_i, [C.POINTER(ClipEbcConfig), C.POINTER(_vp)]),
    "clipebc_model_destroy": (None, [_vp]),
    "clipebc_model_set_tensor": (_i, [_vp, C.c_char_p, _fp, C.POINTER(_i64), _i]),
    "clipebc_model_pack": (_i, [_vp, _vp]),
    "clipebc_forward_windows": (_i, [_vp, _fp, _i, _i, _i, _fp, _fp, _vp]),
    "clipebc_sliding_window_predict": (_i, [_vp, _fp, _i, _i, _i, _i, _i, _i, _fp, _fp, _vp]),
    "clipebc_sliding_window_predict_batch": (_i, [_vp, _i, C.POINTER(_vp), _ip, _ip, _i, _i, _i, _i, C.POINTER(_vp), _fp, _vp]),
    "clipebc_window_origins": (_i, [_i, _i, _i, _i, _i, _i, _ip, _ip, _ip, _ip]),
    "clipebc_f32_to_16": (_i, [_fp, _vp, _i64, _i, _vp]),
    "clipebc_gemm_bf16": (_i, [_i, _vp, _i64, _i64, _i64, _vp, _i64, _i, _i, _i, _i, _ip, _ip, _vp, _i, _fp, _fp, _i,
                               _i, _i, _i, _i, _i, _i, _vp]),
    "clipebc_layernorm": (_i, [_fp, _fp, _fp, _i, _vp, _i, _i64, _i, _i, _i, _vp]),
    "clipebc_attention": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "clipebc_patchify": (_i, [_fp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "clipebc_resample_to_padded": (_i, [_fp, _i, _i, _i, _i, _i, _i, _vp, _fp, _i, _vp]),
    "clipebc_fold_average": (_i, [_fp, _ip, _ip, _i, _i, _i, _i, _i, _i, _fp, _fp, _vp]),
    "clipebc_resize_bicubic_aa": (_i, [_vp, _i, _i, _i, _i, _fp, _fp, _i, _i, _cfp, _cfp, _vp]),
    "clipebc_pad_normalize": (_i, [_vp, _i, _i, _i, _i, _fp, _i, _i, _cfp, _cfp, _vp]),
    "clipebc_resize_density_workspace_floats": (_i, []),
    "clipebc_resize_density_map": (_i, [_fp, _i, _i, _i, _i, _fp, _fp, _fp, _vp]),
}


def load() -> C.CDLL:
    """Load the shared library (once) and attach the prototypes. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C clip_ebc_b200/csrc`). clip_ebc_b200 has no CPU / PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return (load().clipebc_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    """Map a C-ABI return code to the Python error convention."""
    if rc != 0:
        kind = {1: "invalid argument", 2: "CUDA error", 3: "bad state"}.get(rc, f"error {rc}")
        raise RuntimeError(f"clipebc_b200 {what}: {kind}: {last_error()}")


def int_array(values):
    arr = (C.c_int * len(values))(*[int(v) for v in values])
    return arr


def float_array(values):
    return (C.c_float * len(values))(*[float(v) for v in values])
