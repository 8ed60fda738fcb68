"""clip_ebc_b200 -- B200-native (sm_100a) implementation of the CLIP-EBC inference hot path.

Drop-in for the reference's ``get_model()/model(x)`` (models/__init__.py:10-29, models/clip/model.py:191-217) and
``sliding_window_predict`` (utils/eval_utils.py:26-96) for the CLIP ViT-B/16, ViT-B/32 and ViT-L/14 backbones with VPT. All arithmetic runs in
hand-written CUDA kernels behind the C-ABI of ``include/clipebc_b200.h``; there is no CPU or PyTorch fallback.
"""
from ._lib import LIB_PATH, load  # noqa: F401

__all__ = ["get_model", "sliding_window_predict", "sliding_window_predict_batch", "resize_density_map", "calculate_errors", "evaluate", "predict_counts", "write_nwpu_results", "CLIP_EBC",
           "Resize2Multiple", "ZeroPad2Multiple", "load", "LIB_PATH"]


def __getattr__(name):  # lazy: importing the package must not require torch.cuda
    if name in ("get_model", "CLIP_EBC"):
        from . import model

        return getattr(model, name)
    if name in ("sliding_window_predict", "sliding_window_predict_batch", "resize_density_map", "calculate_errors"):
        from . import eval_utils

        return getattr(eval_utils, name)
    if name in ("Resize2Multiple", "ZeroPad2Multiple"):
        from . import transforms

        return getattr(transforms, name)
    if name in ("evaluate", "predict_counts", "write_nwpu_results"):
        from . import eval_loop

        return getattr(eval_loop, name)
    raise AttributeError(name)
