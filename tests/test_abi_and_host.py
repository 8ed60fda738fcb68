"""CPU tests: the C-ABI library loads and exports every symbol include/clipebc_b200.h declares, the host-only integer
logic (window enumeration) matches the oracle and its plain-C restatement bit for bit, and the Python host mirrors the
reference's error conventions. No compute entry point is called (no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(os.path.join(ROOT, "clip_ebc_b200", "libclipebc_b200.so")):
        import __graft_entry__

        __graft_entry__.build()
    from clip_ebc_b200 import _lib

    return _lib.load()


@pytest.fixture(scope="module")
def fold_c():
    path = os.path.join(ROOT, "oracle", "_build", "libfold_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make"], cwd=os.path.join(ROOT, "oracle"))
    so = C.CDLL(path)
    so.oracle_num_windows.restype = C.c_int
    so.oracle_fold.restype = C.c_int
    return so


def test_every_header_symbol_is_exported_and_bound(lib):
    from clip_ebc_b200 import _lib

    header = open(os.path.join(ROOT, "include", "clipebc_b200.h")).read()
    declared = set(re.findall(r"\b(clipebc_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert lib.clipebc_abi_version() == 9
    assert lib.clipebc_launch_count() == 0  # nothing ran on this CPU box


def test_library_is_self_contained():
    """The boundary is a plain C-ABI: no torch / python symbols, no libcuda.so link-time dependency."""
    out = subprocess.check_output(["ldd", os.path.join(ROOT, "clip_ebc_b200", "libclipebc_b200.so")], text=True)
    assert "libtorch" not in out and "libc10" not in out and "libpython" not in out
    assert "libcuda.so" not in out and "not found" not in out


CASES = [(448, 448, 224, 224), (1536, 2048, 224, 112), (3072, 4096, 224, 224), (3072, 4096, 224, 112),
         (448, 672, 224, 112), (300, 500, 224, 200), (224, 224, 224, 224), (225, 1000, 224, 7), (480, 700, 224, 100)]


@pytest.mark.parametrize("H,W,win,s", CASES)
def test_window_origins_bit_exact(lib, fold_c, H, W, win, s):
    from clip_ebc_b200 import ops
    from oracle import clip_ebc_oracle as O

    ro, co = ops.window_origins(H, W, (win, win), (s, s))
    ref_r, ref_c = O.window_origins(H, W, (win, win), (s, s))
    assert (ro, co) == (ref_r, ref_c)
    # plain-C restatement
    n = fold_c.oracle_num_windows(H, win, s)
    arr = (C.c_int * n)()
    fold_c.oracle_window_origins(H, win, s, arr)
    assert list(arr) == ro


def test_c_fold_oracle_matches_numpy_oracle(fold_c):
    from oracle import clip_ebc_oracle as O

    rng = np.random.default_rng(0)
    for (H, W, win, s, r) in [(448, 672, 224, 112, 8), (300, 500, 224, 200, 8), (448, 672, 224, 112, 32)]:
        ro, co = O.window_origins(H, W, (win, win), (s, s))
        g = win // r
        preds = rng.standard_normal((len(ro) * len(co), 1, g, g), dtype=np.float32)
        ref = O.fold_average(preds, H, W, (win, win), (s, s), r)[0]
        out = np.empty((H // r, W // r), np.float32)
        rc = fold_c.oracle_fold(preds.ctypes.data_as(C.c_void_p), H, W, win, win, s, s, r, out.ctypes.data_as(C.c_void_p))
        assert rc == 0 and np.array_equal(out, ref)


def test_window_origins_errors(lib):
    nr, nc = C.c_int(), C.c_int()
    assert lib.clipebc_window_origins(448, 448, 224, 224, 448, 224, C.byref(nr), C.byref(nc), None, None) == 1  # stride > window
    assert b"stride" in lib.clipebc_last_error()
    assert lib.clipebc_window_origins(100, 448, 224, 224, 224, 224, C.byref(nr), C.byref(nc), None, None) == 1  # image < window
    assert lib.clipebc_window_origins(448, 448, 0, 224, 224, 224, C.byref(nr), C.byref(nc), None, None) == 1


def test_model_create_validates_config(lib):
    from clip_ebc_b200 import _lib

    h = C.c_void_p()
    bad = _lib.make_config(224, 12, 32, 1, 5)  # reduction 12
    assert lib.clipebc_model_create(C.byref(bad), C.byref(h)) == 1
    bad2 = _lib.make_config(224, 8, 32, 1, 5, operand_fp16=7)  # operand format
    assert lib.clipebc_model_create(C.byref(bad2), C.byref(h)) == 1
    bad3 = _lib.make_config(224, 8, 32, 1, 5, patch=14, width=1000)  # width
    assert lib.clipebc_model_create(C.byref(bad3), C.byref(h)) == 1
    bad4 = _lib.make_config(224, 8, 32, 1, 5, patch=14, width=1024, layers=24, embed_dim=700)
    assert lib.clipebc_model_create(C.byref(bad4), C.byref(h)) == 1
    for cfg in (_lib.make_config(224, 8, 32, 1, 5), _lib.make_config(224, 8, 32, 1, 5, patch=32),
                _lib.make_config(224, 8, 32, 0, 5, patch=14, width=1024, layers=24, embed_dim=768),
                _lib.make_config(224, 8, 0, 1, 5)):  # num_vpt = 0 is a valid reference configuration
        assert lib.clipebc_model_create(C.byref(cfg), C.byref(h)) == 0, lib.clipebc_last_error()
        # packing without tensors is a state error, reported by name (on a host without a GPU the device check comes first)
        rc = lib.clipebc_model_pack(h, None)
        assert rc in (2, 3) and (rc == 2 or b"missing tensor" in lib.clipebc_last_error())
        # forward before pack is refused as well
        assert lib.clipebc_forward_windows(h, C.c_void_p(16), 1, 224, 224, C.c_void_p(16), None, None) == 3
        lib.clipebc_model_destroy(h)


def test_model_create_rejects_a_struct_of_another_abi(lib):
    """struct_size is the first field: a binding built against an older, shorter clipebc_config (INTEGRATION.md of ABI v8
    listed 7 ints, the struct had 8) is rejected instead of being read past its end."""
    from clip_ebc_b200 import _lib

    class OldConfig(C.Structure):  # the ABI-v8 layout: no struct_size, 8 ints
        _fields_ = [(n, C.c_int) for n in ("input_size", "reduction", "num_vpt", "deep_vpt", "num_bins", "window_chunk",
                                           "operand_fp16", "patch")]

    h = C.c_void_p()
    raw = lib.clipebc_model_create
    old_argtypes = raw.argtypes
    raw.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    try:
        old = OldConfig(224, 8, 32, 1, 5, 0, 1, 16)
        assert raw(C.byref(old), C.byref(h)) == 1          # first word 224 != sizeof(clipebc_config)
        assert b"struct_size" in lib.clipebc_last_error()
        cfg = _lib.make_config(224, 8, 32, 1, 5)
        cfg.struct_size = 0
        assert raw(C.byref(cfg), C.byref(h)) == 1
        cfg.struct_size = C.sizeof(_lib.ClipEbcConfig) + 4
        assert raw(C.byref(cfg), C.byref(h)) == 1
    finally:
        raw.argtypes = old_argtypes
    assert C.sizeof(_lib.ClipEbcConfig) == 56


def test_window_origins_bit_exact_randomised(lib):
    """2000 seeded random geometries (including strides that do not divide anything and H == window): the C-ABI window
    enumeration equals the reference formula int(np.ceil((H - h) / s) + 1) with clamped last origins, bit for bit."""
    from clip_ebc_b200 import ops
    from oracle import clip_ebc_oracle as O

    rng = np.random.default_rng(123)
    for _ in range(2000):
        win_h, win_w = int(rng.integers(16, 513)), int(rng.integers(16, 513))
        sh, sw = int(rng.integers(1, win_h + 1)), int(rng.integers(1, win_w + 1))
        H, W = win_h + int(rng.integers(0, 4000)), win_w + int(rng.integers(0, 4000))
        ro, co = ops.window_origins(H, W, (win_h, win_w), (sh, sw))
        ref_r, ref_c = O.window_origins(H, W, (win_h, win_w), (sh, sw))
        assert (ro, co) == (ref_r, ref_c), (H, W, win_h, win_w, sh, sw)
        assert ro[-1] + win_h <= H and co[-1] + win_w <= W and ro[0] == 0 and co[0] == 0


def test_python_host_mirrors_reference_interface():
    from clip_ebc_b200 import get_model, sliding_window_predict
    from oracle import weights

    r, bins, anchors = weights.bins_and_anchors("r8_t4_nwpu")
    with pytest.raises(AssertionError):
        get_model("clip_vit_b_99", input_size=224, reduction=8, bins=bins, anchor_points=anchors)
    with pytest.raises(AssertionError, match="num_vpt"):
        get_model("clip_vit_b_16", input_size=224, reduction=8, bins=bins, anchor_points=anchors)
    with pytest.raises(AssertionError):  # listed by the reference's vit_backbones but absent from its clip_names (models/__init__.py:21)
        get_model("clip_vit_l_14_336px", input_size=336, reduction=8, bins=bins, anchor_points=anchors, num_vpt=32, vpt_drop=0.0,
                  deep_vpt=True)
    # the CLIP-ResNets: the reference needs none of the ViT arguments (models/clip/model.py:50-52); same state_dict keys
    for name, embed, n_keys in (("resnet50", 1024, 351), ("resnet101", 512, 681), ("resnet50x4", 640, 537)):
        rn = get_model("clip_" + name, input_size=224, reduction=8, bins=bins, anchor_points=anchors,
                       text_features=weights.make_text_features(5, embed=embed))
        sd_rn = weights.make_resnet_state_dict(0, name)
        assert set(sd_rn) == set(rn.state_dict()) and len(sd_rn) == n_keys
        rn.load_state_dict(sd_rn, strict=True)
        assert rn.encoder_reduction == 16 and rn.reduction == 8 and rn.clip_embed_dim == embed
    assert get_model("clip_resnet50", input_size=224, reduction=32, bins=bins, anchor_points=anchors,
                     text_features=weights.make_text_features(5, embed=1024)).encoder_reduction == 32
    model = get_model("CLIP_ViT_B_16", input_size=224, reduction=8, bins=bins, anchor_points=anchors, prompt_type="word",
                      num_vpt=32, vpt_drop=0.0, deep_vpt=True, text_features=weights.make_text_features(5))
    assert model.reduction == 8 and model.bins == bins and tuple(model.anchor_points.shape) == (1, 5, 1, 1)
    sd = weights.make_state_dict(0)
    assert set(sd) == set(model.state_dict())  # reference key names, strict round-trip
    model.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError, match="unexpected|Unexpected|Missing|missing"):
        model.load_state_dict({**sd, "bogus.weight": torch.zeros(1)}, strict=True)
    # ViT-B/32 behind the same entry point: patch 32, 7 x 7 positional grid, same key names
    b32 = get_model("clip_vit_b_32", input_size=224, reduction=8, bins=bins, anchor_points=anchors, prompt_type="word",
                    num_vpt=32, vpt_drop=0.0, deep_vpt=True, text_features=weights.make_text_features(5))
    sd32 = weights.make_state_dict(0, patch=32)
    assert set(sd32) == set(b32.state_dict())
    assert tuple(sd32["image_encoder.conv1.weight"].shape) == (768, 3, 32, 32)
    assert tuple(sd32["image_encoder.positional_embedding"].shape) == (50, 768)
    b32.load_state_dict(sd32, strict=True)
    assert b32.encoder_reduction == 32 and b32.reduction == 8
    # ViT-L/14: width 1024, 24 layers, patch 14 (16 x 16 positional grid), embed 768 -- same key names
    l14 = get_model("clip_vit_l_14", input_size=224, reduction=8, bins=bins, anchor_points=anchors, num_vpt=32, vpt_drop=0.0,
                    deep_vpt=True, text_features=weights.make_text_features(5, embed=768))
    sd14 = weights.make_state_dict(0, patch=14)
    assert set(sd14) == set(l14.state_dict())
    assert tuple(sd14["image_encoder.conv1.weight"].shape) == (1024, 3, 14, 14)
    assert tuple(sd14["image_encoder.positional_embedding"].shape) == (257, 1024)
    assert tuple(sd14["projection.weight"].shape) == (768, 1024, 1, 1)
    l14.load_state_dict(sd14, strict=True)
    assert l14.encoder_reduction == 14 and l14.image_encoder_depth == 24
    with pytest.raises(AssertionError, match="text_features"):  # [N, 512] features do not fit a ViT-L/14 head
        get_model("clip_vit_l_14", input_size=224, reduction=8, bins=bins, anchor_points=anchors, num_vpt=32, vpt_drop=0.0,
                  deep_vpt=True, text_features=weights.make_text_features(5))
    # the reference computes text_features in __init__ (models/clip/model.py:97-129): their absence is a construction error
    with pytest.raises(ValueError, match="text_features is required"):
        get_model("clip_vit_b_16", input_size=224, reduction=8, bins=bins, anchor_points=anchors, num_vpt=32, vpt_drop=0.0,
                  deep_vpt=True)
    # text_encoder.* entries of a reference checkpoint are accepted and round-trip
    model.load_state_dict({**sd, "text_encoder.ln_final.weight": torch.ones(512)}, strict=True)
    assert "text_encoder.ln_final.weight" in model.state_dict()
    x = torch.zeros(1, 3, 448, 448)
    with pytest.raises(AssertionError, match="4D"):
        sliding_window_predict(model, x[0], 224, 224)
    with pytest.raises(AssertionError, match="Stride must be smaller"):
        sliding_window_predict(model, x, 224, 448)
    with pytest.raises(RuntimeError, match="CUDA device only"):  # no CPU fallback
        sliding_window_predict(model, x, 224, 224)
    shallow = get_model("clip_vit_b_16", input_size=224, reduction=16, bins=bins, anchor_points=anchors, num_vpt=32,
                        vpt_drop=0.0, deep_vpt=False, text_features=weights.make_text_features(5))
    assert [k for k in shallow.state_dict() if k.startswith("vpt_")] == ["vpt_0"]
    # num_vpt = 0 (accepted by the reference): the zero-sized prompt tensors are not part of what the native side loads
    bare = get_model("clip_vit_b_16", input_size=224, reduction=8, bins=bins, anchor_points=anchors, num_vpt=0, vpt_drop=0.0,
                     deep_vpt=True, text_features=weights.make_text_features(5))
    assert tuple(bare.vpt_0.shape) == (0, 768)
    assert not any(k.startswith("vpt_") for k in bare._hot_path_tensors())
