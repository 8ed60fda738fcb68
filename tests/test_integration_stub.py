"""INTEGRATION.md section 2 shows the ctypes binding a reference maintainer would write. This test extracts that code block from
the document and executes it VERBATIM against the built library, so the document cannot drift from the ABI again (the
round-1 stub described a 7-int config while the struct had 8).

CPU: the block runs (library loads, ABI version matches, prototypes resolve), its Config is accepted by
clipebc_model_create, and its to_native() reaches the library (which refuses to upload without a CUDA device).
B200: the stub's to_native / native_forward / native_sliding_window_predict, driven with a stand-in for the reference module,
give exactly the bits clip_ebc_b200's own host layer gives."""
import ctypes as C
import os
import re
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_namespace():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = text[text.index("## 2. Binding the C-ABI directly"):]
    block = re.search(r"```python\n(.*?)```", sec, flags=re.S).group(1)
    assert "class Config(C.Structure)" in block and "def to_native" in block
    ns = {}
    cwd = os.getcwd()
    os.chdir(ROOT)  # the stub loads "clip_ebc_b200/libclipebc_b200.so" relative to the repository root
    try:
        exec(compile(block, "INTEGRATION.md#2", "exec"), ns)
    finally:
        os.chdir(cwd)
    return ns


def _reference_standin(device):
    """What the stub needs of the reference's CLIP_EBC: state_dict(), text_features, anchor_points, reduction, num_vpt,
    deep_vpt, bins."""
    from oracle import weights

    reduction, bins, anchors = weights.bins_and_anchors("r8_t4_nwpu")
    sd = {k: v.to(device) for k, v in weights.make_state_dict(3, variant="stress").items()}
    sd["text_encoder.ln_final.weight"] = torch.ones(512, device=device)  # a reference checkpoint carries the text tower
    tf = weights.make_text_features(len(bins), seed=103)
    ref = types.SimpleNamespace(reduction=reduction, num_vpt=32, deep_vpt=True, bins=bins, text_features=tf.to(device),
                                anchor_points=torch.tensor(anchors).view(1, -1, 1, 1).to(device),
                                state_dict=lambda: sd)
    return ref, (sd, tf, bins, anchors, reduction)


def test_stub_executes_and_its_config_is_accepted():
    if not os.path.exists(os.path.join(ROOT, "clip_ebc_b200", "libclipebc_b200.so")):
        import __graft_entry__

        __graft_entry__.build()
    ns = _stub_namespace()
    from clip_ebc_b200 import _lib

    assert C.sizeof(ns["Config"]) == C.sizeof(_lib.ClipEbcConfig)
    assert [f[0] for f in ns["Config"]._fields_] == [f[0] for f in _lib.ClipEbcConfig._fields_]
    for backbone, dims in ns["BACKBONES"].items():
        cfg = ns["make_config"](backbone, 8, 32, True, 5)
        h = C.c_void_p()
        ns["check"](ns["lib"].clipebc_model_create(C.byref(cfg), C.byref(h)))
        ns["lib"].clipebc_model_destroy(h)
    if not torch.cuda.is_available():
        ref, _ = _reference_standin("cpu")
        with pytest.raises(RuntimeError):   # reaches clipebc_model_set_tensor, which needs the handle's CUDA device
            ns["to_native"](ref)


@pytest.mark.gpu
def test_stub_outputs_equal_the_maintained_host_layer():
    from clip_ebc_b200 import get_model, sliding_window_predict
    from oracle import weights

    ns = _stub_namespace()
    ref, (sd, tf, bins, anchors, reduction) = _reference_standin("cuda")
    h = ns["to_native"](ref)
    model = get_model("clip_vit_b_16", input_size=224, reduction=reduction, bins=bins, anchor_points=anchors,
                      prompt_type="word", num_vpt=32, vpt_drop=0.0, deep_vpt=True, text_features=tf)
    model.load_state_dict({k: v.cpu() for k, v in sd.items()}, strict=True)
    model = model.to("cuda").eval()
    model.use_cuda_graphs = False
    x = weights.make_image((3, 3, 224, 224), seed=400).cuda()
    assert torch.equal(ns["native_forward"](h, x, reduction), model(x))
    img = weights.make_image((1, 3, 448, 672), seed=401).cuda()
    assert torch.equal(ns["native_sliding_window_predict"](h, img, 224, 112, reduction),
                       sliding_window_predict(model, img, 224, 112))
    ns["lib"].clipebc_model_destroy(h)
