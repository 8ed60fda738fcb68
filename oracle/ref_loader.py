"""TEST INFRASTRUCTURE -- runs only in the build container, where /root/reference is mounted read-only.

Loads the *real* reference classes (CLIP_EBC, sliding_window_predict) file by file. The reference cannot be imported
as a package here: `timm`, `ftfy`, `tensorboardX` are missing and `models/clip/_clip/__init__.py:31-41` downloads
checkpoints at import. Nothing from the reference is copied into this repository; the modules are executed where
they lie. Used by oracle/make_golden.py and tests/test_oracle_vs_reference.py to pin the oracle.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

REF = os.environ.get("CLIPEBC_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "models", "clip"))


_cache = {}


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _stub_pkg(name, path):
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    sys.modules[name] = mod
    return mod


def load_reference():
    """Returns (clip_model_module, eval_utils_module) of the reference."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    if "ftfy" not in sys.modules:
        ftfy = types.ModuleType("ftfy")
        ftfy.fix_text = lambda s: s  # prompts are ASCII
        sys.modules["ftfy"] = ftfy
    _stub_pkg("models", f"{REF}/models")
    _load("models.utils", f"{REF}/models/utils.py")
    _stub_pkg("models.clip", f"{REF}/models/clip")
    _load("models.clip.utils", f"{REF}/models/clip/utils.py")
    clip_pkg = _stub_pkg("models.clip._clip", f"{REF}/models/clip/_clip")
    _load("models.clip._clip.blocks", f"{REF}/models/clip/_clip/blocks.py")
    ie = _load("models.clip._clip.image_encoder", f"{REF}/models/clip/_clip/image_encoder.py")
    te = _load("models.clip._clip.text_encoder", f"{REF}/models/clip/_clip/text_encoder.py")
    tok = _load("models.clip._clip.simple_tokenizer", f"{REF}/models/clip/_clip/simple_tokenizer.py").SimpleTokenizer()

    def vit_b_16_img(features_only=False, input_size=None, **kw):
        m = ie.VisionTransformer(input_resolution=224, patch_size=16, output_dim=512, width=768, layers=12, heads=12,
                                 features_only=features_only)
        if input_size is not None:
            m.adjust_pos_embed(input_size, input_size)
        return m

    def vit_b_32_img(features_only=False, input_size=None, **kw):
        m = ie.VisionTransformer(input_resolution=224, patch_size=32, output_dim=512, width=768, layers=12, heads=12,
                                 features_only=features_only)
        if input_size is not None:
            m.adjust_pos_embed(input_size, input_size)
        return m

    def vit_l_14_img(features_only=False, input_size=None, **kw):
        m = ie.VisionTransformer(input_resolution=224, patch_size=14, output_dim=768, width=1024, layers=24, heads=16,
                                 features_only=features_only)
        if input_size is not None:
            m.adjust_pos_embed(input_size, input_size)
        return m

    def vit_l_14_txt():
        m = te.CLIPTextEncoder(embed_dim=768, context_length=77, vocab_size=49408, transformer_width=768,
                               transformer_heads=12, transformer_layers=12)
        torch.nn.init.normal_(m.positional_embedding, std=0.01)
        torch.nn.init.normal_(m.text_projection, std=768 ** -0.5)
        return m

    def vit_b_16_txt():
        m = te.CLIPTextEncoder(embed_dim=512, context_length=77, vocab_size=49408, transformer_width=512,
                               transformer_heads=8, transformer_layers=12)
        torch.nn.init.normal_(m.positional_embedding, std=0.01)  # torch.empty in the reference (text_encoder.py:28,31)
        torch.nn.init.normal_(m.text_projection, std=512 ** -0.5)
        return m

    def tokenize(texts, context_length=77):
        sot, eot = tok.encoder["<|startoftext|>"], tok.encoder["<|endoftext|>"]
        out = torch.zeros(len(texts), context_length, dtype=torch.int)
        for i, t in enumerate(texts):
            ids = [sot] + tok.encode(t) + [eot]
            out[i, :len(ids)] = torch.tensor(ids)
        return out

    clip_pkg.vit_b_16_img, clip_pkg.vit_b_16_txt, clip_pkg.tokenize = vit_b_16_img, vit_b_16_txt, tokenize
    clip_pkg.vit_b_32_img, clip_pkg.vit_b_32_txt = vit_b_32_img, vit_b_16_txt  # same text tower (embed_dim 512)
    clip_pkg.vit_l_14_img, clip_pkg.vit_l_14_txt = vit_l_14_img, vit_l_14_txt
    cm = _load("models.clip.model", f"{REF}/models/clip/model.py")
    ev = _load("ref_eval_utils", f"{REF}/utils/eval_utils.py")
    _cache["mods"] = (cm, ev)
    return cm, ev


def build_reference_model(sd, text_features, bins, anchor_points, reduction, num_vpt=32, deep_vpt=True,
                          input_size=224, backbone="vit_b_16"):
    """The reference's own CLIP_EBC (via its `_clip_ebc` factory), loaded with `sd`; text features overridden by the
    supplied constant (a plain attribute of the module, models/clip/model.py:129)."""
    cm, _ = load_reference()
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):  # the constructor prints the prompts (model.py:122)
        model = cm._clip_ebc(backbone=backbone, input_size=input_size, reduction=reduction, bins=bins,
                             anchor_points=anchor_points, prompt_type="word", num_vpt=num_vpt, vpt_drop=0.0,
                             deep_vpt=deep_vpt)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("text_encoder.") for k in missing), [k for k in missing if not k.startswith("text_encoder.")]
    model.text_features = text_features.clone()
    return model.eval()
