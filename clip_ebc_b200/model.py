"""Host-side mirror of the reference model boundary: ``get_model()`` and the ``CLIP_EBC`` module.

Reference interface being mirrored (same names, argument meaning and error behaviour):
  get_model(backbone, input_size, reduction, bins, anchor_points, **kwargs)   /root/reference/models/__init__.py:10-44
  _clip_ebc(...) / CLIP_EBC.__init__ / CLIP_EBC.forward                        /root/reference/models/clip/model.py:30-270

The module is a *parameter container* with the reference's ``state_dict`` keys (``vpt_{i}``, ``logit_scale``,
``image_encoder.*``, ``image_decoder.0.*``, ``projection.*``) so reference checkpoints load with ``strict=True``;
``forward`` hands raw device pointers to the C-ABI (``clipebc_forward_windows``). No arithmetic of the hot path is done
in PyTorch and there is no fallback: a CPU tensor or a missing ``libclipebc_b200.so`` raises.

Out of scope here: the CLIP text tower and tokenizer. They run once at construction in the
reference (models/clip/model.py:97-129) and produce the constant ``text_features`` [N, embed_dim]; this module takes that
matrix as a constructor argument (``text_features=``; construction FAILS without it, as the reference's would without its
text tower) and keeps any ``text_encoder.*`` checkpoint entries verbatim so a ``state_dict`` round-trips.

Train mode (``model.train()`` / ``model.training = True``) only switches the output signature to the reference's
``(logits, exp)`` (models/clip/model.py:214-217): this is an inference build -- BatchNorm uses its running statistics,
VPT dropout is not applied and no autograd graph is recorded (a warning says so once when gradients are enabled).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import warnings
from collections import OrderedDict
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib

clip_names = ["resnet50", "resnet50x4", "resnet50x16", "resnet50x64", "resnet101", "vit_b_16", "vit_b_32", "vit_l_14"]
resnet_backbones = ["resnet50", "resnet101", "resnet50x4", "resnet50x16", "resnet50x64"]
vit_backbones = ["vit_b_16", "vit_b_32", "vit_l_14", "vit_l_14_336px"]

# backbone -> (patch size = encoder reduction, width, layers, embed_dim); heads = width // 64 (_clip/model.py:50),
# models/clip/model.py:16-24
_VIT_DIMS = {"vit_b_16": (16, 768, 12, 512), "vit_b_32": (32, 768, 12, 512), "vit_l_14": (14, 1024, 24, 768)}
# CLIP-ResNet backbones: backbone -> (blocks per layer, stem width, embed_dim, decoder_cfg) -- the configs prepare.py extracts from
# the OpenAI checkpoints (_clip/__init__.py:73-96) and models/clip/model.py:228-239
_RESNET_DIMS = {"resnet50": ((3, 4, 6, 3), 64, 1024, (2048,)), "resnet101": ((3, 4, 23, 3), 64, 512, (2048, 1024)),
                "resnet50x4": ((4, 6, 10, 6), 80, 640, (1280,)), "resnet50x16": ((6, 8, 18, 8), 96, 768, (1536,)),
                "resnet50x64": ((3, 15, 36, 10), 128, 1024, (2048,))}


class _Block(nn.Module):
    """Parameter names of ResidualAttentionBlock (_clip/blocks.py:22-33). Never called."""

    def __init__(self, d: int, heads: int) -> None:
        super().__init__()
        self.attn = nn.MultiheadAttention(d, heads)
        self.ln_1 = nn.LayerNorm(d)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d, d * 4)), ("gelu", nn.Identity()),
                                              ("c_proj", nn.Linear(d * 4, d))]))
        self.ln_2 = nn.LayerNorm(d)


class _Transformer(nn.Module):
    def __init__(self, d: int, layers: int, heads: int) -> None:
        super().__init__()
        self.resblocks = nn.Sequential(*[_Block(d, heads) for _ in range(layers)])


class _ImageEncoder(nn.Module):
    """Parameter names of VisionTransformer(features_only=True) (_clip/image_encoder.py:118-160)."""

    def __init__(self, input_size: int, patch: int, width: int, layers: int, embed: int) -> None:
        super().__init__()
        scale = width ** -0.5
        g = input_size // patch
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch, stride=patch, bias=False)
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(g * g + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = _Transformer(width, layers, width // 64)
        self.ln_post = nn.LayerNorm(width)
        self.patch_size = (patch, patch)
        self.channels = width
        self.reduction = patch
        self.clip_embed_dim = embed


def _conv_bn(prefix_conv: str, prefix_bn: str, mod: nn.Module, cin: int, cout: int, k: int, stride: int = 1) -> None:
    setattr(mod, prefix_conv, nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, bias=False))
    setattr(mod, prefix_bn, nn.BatchNorm2d(cout))


class _ClipBottleneck(nn.Module):
    """Parameter names of the CLIP Bottleneck (_clip/blocks.py:56-86). Never called."""

    def __init__(self, inplanes: int, planes: int, stride: int) -> None:
        super().__init__()
        _conv_bn("conv1", "bn1", self, inplanes, planes, 1)
        _conv_bn("conv2", "bn2", self, planes, planes, 3)
        _conv_bn("conv3", "bn3", self, planes, planes * 4, 1)
        self.downsample = None
        if stride > 1 or inplanes != planes * 4:
            self.downsample = nn.Sequential(OrderedDict([("-1", nn.AvgPool2d(stride)),
                                                         ("0", nn.Conv2d(inplanes, planes * 4, 1, bias=False)),
                                                         ("1", nn.BatchNorm2d(planes * 4))]))


class _ModifiedResNet(nn.Module):
    """Parameter names of ModifiedResNet(features_only=True) (_clip/image_encoder.py:33-58)."""

    def __init__(self, layers: Tuple[int, int, int, int], width: int, embed: int, reduction: int) -> None:
        super().__init__()
        _conv_bn("conv1", "bn1", self, 3, width // 2, 3, stride=2)
        _conv_bn("conv2", "bn2", self, width // 2, width // 2, 3)
        _conv_bn("conv3", "bn3", self, width // 2, width, 3)
        inplanes = width
        for li, (planes, blocks) in enumerate(zip((width, 2 * width, 4 * width, 8 * width), layers), start=1):
            stride = 1 if li == 1 or (li == 4 and reduction <= 16) else 2
            seq = [_ClipBottleneck(inplanes, planes, stride)]
            inplanes = planes * 4
            seq += [_ClipBottleneck(inplanes, planes, 1) for _ in range(1, blocks)]
            setattr(self, f"layer{li}", nn.Sequential(*seq))
        self.channels = inplanes
        self.reduction = 16 if reduction <= 16 else 32
        self.clip_embed_dim = embed


class _DecoderBottleneck(nn.Module):
    """Parameter names of models/utils.py Bottleneck (expansion 1): the decoder block of the ResNet backbones."""

    def __init__(self, cin: int, cout: int) -> None:
        super().__init__()
        _conv_bn("conv1", "bn1", self, cin, cout, 1)
        _conv_bn("conv2", "bn2", self, cout, cout, 3)
        _conv_bn("conv3", "bn3", self, cout, cout, 1)
        self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, bias=False), nn.BatchNorm2d(cout)) if cin != cout else nn.Identity()


class _BasicBlock(nn.Module):
    """Parameter names of BasicBlock(768, 768) (models/utils.py:254-288)."""

    def __init__(self, c: int) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(c)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(c)


def _init_decoder(m: nn.Module) -> None:
    # models/utils.py:366-379
    for mod in m.modules():
        if isinstance(mod, nn.Conv2d):
            nn.init.kaiming_normal_(mod.weight, mode="fan_out", nonlinearity="relu")
            if mod.bias is not None:
                nn.init.constant_(mod.bias, 0.0)
        elif isinstance(mod, nn.BatchNorm2d):
            nn.init.constant_(mod.weight, 1.0)
            nn.init.constant_(mod.bias, 0.0)


class CLIP_EBC(nn.Module):
    """B200-native CLIP-EBC (ViT-B/16, ViT-B/32, ViT-L/14 + VPT, or CLIP-ResNet-50 / -101). Constructor arguments as in
    models/clip/model.py:31-45,
    plus: text_features (required, see the module docstring), window_chunk (windows per internal pass, 0 = default),
    operand_dtype ("fp16" | "bf16": the 16-bit tensor-core operand format) and decoder_conv1_fine (A/B switch of the
    decoder's conv1 form, see include/clipebc_b200.h)."""

    def __init__(
        self,
        backbone: str,
        bins: List[Tuple[float, float]],
        anchor_points: List[float],
        reduction: Optional[int] = None,
        freeze_text_encoder: bool = True,
        prompt_type: str = "number",
        input_size: Optional[int] = None,
        num_vpt: Optional[int] = None,
        deep_vpt: Optional[bool] = None,
        vpt_drop: Optional[float] = None,
        decoder_block: Any = None,
        decoder_cfg: Optional[List[Union[str, int]]] = None,
        text_features: Optional[Tensor] = None,
        window_chunk: int = 0,
        operand_dtype: str = "fp16",
        decoder_conv1_fine: bool = False,
    ) -> None:
        super().__init__()
        assert backbone in resnet_backbones + vit_backbones, \
            f"Backbone should be in {resnet_backbones + vit_backbones}, got {backbone}"
        if backbone not in _VIT_DIMS and backbone not in _RESNET_DIMS:
            raise NotImplementedError(
                f"clip_ebc_b200 implements the hot path for {sorted(_VIT_DIMS) + sorted(_RESNET_DIMS)} (got '{backbone}').")
        self.is_resnet = backbone in _RESNET_DIMS
        if self.is_resnet:
            rn_layers, rn_width, embed, rn_decoder = _RESNET_DIMS[backbone]
            patch, width, layers = 0, 0, 0
            assert reduction is not None, "Expected reduction to be an integer for the CLIP-ResNet backbones, got None."
        else:
            patch, width, layers, embed = _VIT_DIMS[backbone]
            assert input_size is not None, "Expected input_size to be an integer, got None."
            assert num_vpt is not None, "Expected num_vpt to be an integer, got None."
            assert deep_vpt is not None, "Expected deep_vpt to be a boolean, got None."
            assert vpt_drop is not None, "Expected vpt_drop to be a float, got None."
        self.patch, self.width, self.layers, self.embed_dim = patch, width, layers, embed
        assert prompt_type in ["number", "word"], f"Expected prompt_type to be 'number' or 'word', got {prompt_type}"
        if not freeze_text_encoder:
            raise NotImplementedError("freeze_text_encoder=False (training the text tower) is outside the inference hot path")
        default_cfg = list(rn_decoder) if self.is_resnet else [width]
        if decoder_cfg is not None and list(decoder_cfg) != default_cfg:
            raise NotImplementedError(f"only the reference default decoder_cfg={default_cfg} is implemented")
        assert bins is not None and anchor_points is not None and len(bins) == len(anchor_points)
        if text_features is None:
            # the reference computes this matrix here, in __init__, with its CLIP text tower (models/clip/model.py:97-129);
            # that tower is outside this build, so the matrix is an input -- and its absence is a construction error, not
            # something to discover at the first forward
            raise ValueError(
                f"text_features is required: pass text_features=[{len(bins)}, {embed}] (the CLIP text-tower output for the "
                "bin prompts; with the reference installed: reference_model.text_features). The text tower and tokenizer "
                "are outside the scope of clip_ebc_b200.")

        self.backbone = backbone
        if self.is_resnet:
            # models/clip/model.py:50-52: ModifiedResNet(features_only=True, out_indices=(-1,), reduction=reduction)
            self.image_encoder = _ModifiedResNet(rn_layers, rn_width, embed, int(reduction))
            self.input_size = 224 if input_size is None else int(input_size)
            self.num_vpt, self.deep_vpt = 0, False
            self.encoder_reduction = self.image_encoder.reduction
            self.reduction = int(reduction)
            blocks, cin = [], self.image_encoder.channels
            for cout in rn_decoder:  # make_resnet_layers(Bottleneck, decoder_cfg, expansion=1), model.py:83-87
                blocks.append(_DecoderBottleneck(cin, cout))
                cin = cout
            self.channels = cin
            self.image_decoder = nn.Sequential(*blocks)
        else:
            self.image_encoder = _ImageEncoder(int(input_size), patch, width, layers, embed)
            self.image_encoder_depth = layers
            for p in self.image_encoder.parameters():
                p.requires_grad = False
            self.num_vpt = int(num_vpt)
            self.deep_vpt = bool(deep_vpt)
            self.input_size = int(input_size)
            val = math.sqrt(6.0 / float(3 * patch + width))  # model.py:70-75
            for idx in range(layers if self.deep_vpt else 1):
                p = nn.Parameter(torch.empty(self.num_vpt, width))
                nn.init.uniform_(p, -val, val)
                setattr(self, f"vpt_{idx}", p)
            self.vpt_drop = float(vpt_drop)  # identity in eval mode; the inference path has no dropout
            self.encoder_reduction = patch
            self.reduction = self.encoder_reduction if reduction is None else int(reduction)
            self.channels = width
            self.image_decoder = nn.Sequential(_BasicBlock(width))
        self.clip_embed_dim = embed
        _init_decoder(self.image_decoder)
        self.projection = nn.Conv2d(self.channels, embed, kernel_size=1)
        _init_decoder(self.projection)

        self.prompt_type = prompt_type
        self.freeze_text_encoder = freeze_text_encoder
        self.bins = bins
        self.anchor_points = torch.tensor(anchor_points, dtype=torch.float32, requires_grad=False).view(1, -1, 1, 1)
        self.text_features: Optional[Tensor] = None
        self.set_text_features(text_features)
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07), requires_grad=True)

        assert operand_dtype in ("fp16", "bf16"), f"operand_dtype must be 'fp16' or 'bf16', got {operand_dtype}"
        self.operand_dtype = operand_dtype  # 16-bit tensor-core operand format (accumulation / residual stay fp32)
        self.decoder_conv1_fine = bool(decoder_conv1_fine)
        self._text_encoder_state: "OrderedDict[str, Tensor]" = OrderedDict()
        self._window_chunk = int(window_chunk)
        self._handle: Optional[C.c_void_p] = None
        self._handle_device: Optional[int] = None
        self._packed_key = None
        self.use_cuda_graphs = True   # forward(): replay a captured CUDA graph per input shape (see _forward_graphed)
        self._graph_cache: dict = {}
        self._capture_streams: dict = {}  # device index -> the side stream graphs of that device are captured on
        self._warned_train = False

    # ------------------------------------------------------------------ text features (constant input of the head)
    def set_text_features(self, text_features: Tensor) -> None:
        """[N, embed_dim] output of the (out-of-scope) CLIP text tower for the bin prompts (model.py:127-129)."""
        tf = torch.as_tensor(text_features, dtype=torch.float32).detach()
        assert tf.dim() == 2 and tf.shape[0] == len(self.bins) and tf.shape[1] == self.embed_dim, \
            f"Expected text_features of shape ({len(self.bins)}, {self.embed_dim}), got {tuple(tf.shape)}"
        self.text_features = tf
        self._packed_key = None

    # ------------------------------------------------------------------ state_dict compatibility
    def state_dict(self, *args, **kwargs):
        sd = super().state_dict(*args, **kwargs)
        prefix = kwargs.get("prefix", args[1] if len(args) > 1 else "")
        for k, v in self._text_encoder_state.items():
            sd[prefix + k] = v
        return sd

    def load_state_dict(self, state_dict, strict: bool = True, **kwargs):
        own = OrderedDict()
        self._text_encoder_state = OrderedDict()
        for k, v in state_dict.items():
            if k.startswith("text_encoder."):
                self._text_encoder_state[k] = v  # kept verbatim; the text tower is not on the hot path
            else:
                own[k] = v
        out = super().load_state_dict(own, strict=strict, **kwargs)
        self._packed_key = None
        return out

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .float()
        out = super()._apply(fn, *args, **kwargs)
        self._packed_key = None
        return out

    # ------------------------------------------------------------------ native handle
    def _device(self) -> torch.device:
        return self.logit_scale.device

    def _ensure_packed(self) -> None:
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("clip_ebc_b200.CLIP_EBC runs on a CUDA device only (no CPU fallback): call .to('cuda')")
        tensors = self._hot_path_tensors()
        dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        key = (dev_index, tuple((k, t.data_ptr(), t._version) for k, t in tensors.items()))
        if self._handle is not None and key == self._packed_key:
            return
        lib = _lib.load()
        with torch.cuda.device(dev):
            if self._handle is not None and self._handle_device != dev_index:
                # .to(another GPU): a native handle owns buffers on the device it was created on -- start over there
                self._release_handle()
            if self._handle is None:
                cfg = _lib.make_config(self.input_size, self.reduction, self.num_vpt, int(self.deep_vpt), len(self.bins),
                                       self._window_chunk, int(self.operand_dtype == "fp16"), self.patch, self.width,
                                       self.layers, self.embed_dim, int(self.decoder_conv1_fine), int(self.is_resnet))
                h = C.c_void_p()
                _lib.check(lib.clipebc_model_create(C.byref(cfg), C.byref(h)), "model_create")
                self._handle, self._handle_device = h, dev_index
            for name, t in tensors.items():
                t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
                shape = (C.c_int64 * t.dim())(*t.shape)
                _lib.check(lib.clipebc_model_set_tensor(self._handle, name.encode(), t.data_ptr(), shape, t.dim()),
                           f"set_tensor({name})")
            _lib.check(lib.clipebc_model_pack(self._handle, torch.cuda.current_stream().cuda_stream), "model_pack")
        self._packed_key = key
        self._graph_cache.clear()  # captured graphs read the packed weights of the previous state

    def _hot_path_tensors(self) -> "OrderedDict[str, Tensor]":
        out = OrderedDict()
        for k, v in super().state_dict().items():
            if k.endswith("num_batches_tracked") or v.numel() == 0:  # num_vpt == 0: the (0, width) prompts carry nothing
                continue
            out[k] = v
        out["text_features"] = self.text_features
        out["anchor_points"] = self.anchor_points.reshape(-1)
        return out

    def _release_handle(self) -> None:
        if self._handle is not None:
            self._graph_cache.clear()
            _lib.load().clipebc_model_destroy(self._handle)
            self._handle, self._handle_device, self._packed_key = None, None, None

    def __del__(self):
        try:
            self._release_handle()
        except Exception:
            pass

    # ------------------------------------------------------------------ forward (models/clip/model.py:191-217)
    def forward(self, x: Tensor) -> Union[Tensor, Tuple[Tensor, Tensor]]:
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"Expected input of shape (B, 3, H, W), got {tuple(x.shape)}")
        self._ensure_packed()
        dev = self._device()
        if x.device != dev:
            raise RuntimeError(f"input is on {x.device} but the model is on {dev}")
        if self.training and torch.is_grad_enabled() and not self._warned_train:
            self._warned_train = True
            warnings.warn("clip_ebc_b200.CLIP_EBC is an inference build: in train mode forward() returns the reference's "
                          "(logits, exp) pair, but with eval semantics (BatchNorm running statistics, no VPT dropout) and "
                          "without an autograd graph -- the outputs carry no gradients.", stacklevel=2)
        x = x.detach()
        with torch.cuda.device(dev):
            if self._graphs_usable():
                return self._forward_graphed(x)
            return self._forward_eager(x.to(torch.float32).contiguous())

    def _forward_eager(self, x: Tensor):
        """One C-ABI call on torch's current stream; x: float32, contiguous, on the model's device."""
        dev = x.device
        B, _, h, w = x.shape
        r = self.reduction
        exp = torch.empty((B, 1, h // r, w // r), dtype=torch.float32, device=dev)
        logits = torch.empty((B, len(self.bins), h // r, w // r), dtype=torch.float32, device=dev) if self.training else None
        _lib.check(_lib.load().clipebc_forward_windows(
            self._handle, x.data_ptr(), B, h, w, exp.data_ptr(), None if logits is None else logits.data_ptr(),
            torch.cuda.current_stream().cuda_stream), "forward_windows")
        if self.training:
            return logits, exp
        return exp

    # A forward is ~85 dependent launches (PDL-chained). Replaying them as one CUDA graph takes 2-6 % off a call
    # (profiles/graph_probe.py: 0.99 -> 0.94 ms at 1 window, 4.43 -> 4.32 ms at 64): the second call with a given input
    # shape captures the launches (static input / output buffers owned by the cache), later calls copy the input in,
    # replay and return copies of the outputs. Bit-identical to the eager path (same kernels, same order).
    _GRAPH_CACHE_ENTRIES = 4

    def _graphs_usable(self) -> bool:
        if not self.use_cuda_graphs or os.environ.get("CLIPEBC_NO_GRAPHS"):
            return False
        lib = _lib.load()
        # per-launch profiling brackets every kernel with events; an outer capture must see our launches directly
        return not lib.clipebc_profile_enabled() and not torch.cuda.is_current_stream_capturing()

    def _forward_graphed(self, x: Tensor):
        lib = _lib.load()
        key = (tuple(x.shape), bool(self.training), self._packed_key is not None and id(self._packed_key),
               int(lib.clipebc_config_epoch()), torch.cuda.current_device())
        entry = self._graph_cache.get(key)
        if entry is None:
            # first sighting of this shape: plain call (it also sizes the library's workspaces, uploads index tables and
            # raises the L2 carve-out -- none of which may happen inside a capture)
            while len(self._graph_cache) >= self._GRAPH_CACHE_ENTRIES:
                self._graph_cache.pop(next(iter(self._graph_cache)))
            self._graph_cache[key] = {}
            return self._forward_eager(x.to(torch.float32).contiguous())
        if entry.get("disabled"):
            return self._forward_eager(x.to(torch.float32).contiguous())
        if "graph" not in entry:
            static_x = torch.empty(x.shape, dtype=torch.float32, device=x.device)
            static_x.copy_(x)
            graph = torch.cuda.CUDAGraph()
            l0 = lib.clipebc_launch_count()
            try:
                # thread_local: only THIS thread's CUDA calls are checked during the capture. The default ("global") makes
                # an unrelated cudaMalloc / event query of any other thread -- a DataLoader's pin-memory thread, another
                # model -- fail with "operation not permitted when stream is capturing" while we capture.
                # an explicit capture stream ON THE MODEL'S DEVICE: torch's default capture stream is created once, on
                # whichever device captured first, and a capture on a stream of another device records nothing
                dev_index = x.device.index if x.device.index is not None else torch.cuda.current_device()
                cap = self._capture_streams.get(dev_index)
                if cap is None:
                    cap = self._capture_streams[dev_index] = torch.cuda.Stream(x.device)
                with torch.cuda.graph(graph, stream=cap, capture_error_mode="thread_local"):
                    outs = self._forward_eager(static_x)
            except Exception:
                # a capture that cannot be completed (e.g. an allocation inside the library for a shape the eager call
                # did not size) must not take the call down: this shape stays eager from now on
                entry["disabled"] = True
                torch.cuda.synchronize()
                return self._forward_eager(static_x)
            entry.update(graph=graph, x=static_x, outs=outs, launches=int(lib.clipebc_launch_count() - l0))
        else:
            entry["x"].copy_(x)
        entry["graph"].replay()
        lib.clipebc_note_replayed_launches(entry["launches"])
        outs = entry["outs"]
        if isinstance(outs, tuple):
            return tuple(o.clone() for o in outs)
        return outs.clone()

    # ------------------------------------------------------------------ fused sliding-window entry (eval_utils.py:26-96)
    def sliding_window_density(self, image: Tensor, window_size: Tuple[int, int], stride: Tuple[int, int],
                               with_count: bool = False):
        """image [1,3,H,W] on the model's device -> density [1,1,H//r,W//r] (device) and optionally its sum [1]."""
        self._ensure_packed()
        dev = self._device()
        if image.device != dev:
            raise RuntimeError(f"image is on {image.device} but the model is on {dev}")
        if image.dim() != 4 or image.shape[0] != 1 or image.shape[1] != 3:
            raise RuntimeError(f"Expected image of shape (1, 3, H, W), got {tuple(image.shape)}")
        image = image.detach().to(torch.float32).contiguous()
        H, W = image.shape[-2:]
        r = self.reduction
        dens = torch.empty((1, 1, H // r, W // r), dtype=torch.float32, device=dev)
        cnt = torch.empty((1,), dtype=torch.float32, device=dev) if with_count else None
        with torch.cuda.device(dev):
            _lib.check(_lib.load().clipebc_sliding_window_predict(
                self._handle, image.data_ptr(), H, W, window_size[0], window_size[1], stride[0], stride[1],
                dens.data_ptr(), None if cnt is None else cnt.data_ptr(), torch.cuda.current_stream().cuda_stream),
                "sliding_window_predict")
        return (dens, cnt) if with_count else dens


    def sliding_window_density_batch(self, images, window_size: Tuple[int, int], stride: Tuple[int, int]):
        """images: list of [1,3,H_i,W_i] tensors on the model's device (sizes may differ) -> (list of densities
        [1,1,H_i//r,W_i//r], counts [n] on the device). One C-ABI call: the windows of all images share the ViT passes."""
        self._ensure_packed()
        dev = self._device()
        n = len(images)
        if n == 0:
            raise RuntimeError("empty image batch")
        imgs, hs, ws, dens = [], [], [], []
        r = self.reduction
        for im in images:
            if im.device != dev:
                raise RuntimeError(f"image is on {im.device} but the model is on {dev}")
            if im.dim() != 4 or im.shape[0] != 1 or im.shape[1] != 3:
                raise RuntimeError(f"Expected image of shape (1, 3, H, W), got {tuple(im.shape)}")
            im = im.detach().to(torch.float32).contiguous()
            imgs.append(im)
            hs.append(int(im.shape[-2])); ws.append(int(im.shape[-1]))
            dens.append(torch.empty((1, 1, hs[-1] // r, ws[-1] // r), dtype=torch.float32, device=dev))
        cnt = torch.empty((n,), dtype=torch.float32, device=dev)
        img_ptrs = (C.c_void_p * n)(*[im.data_ptr() for im in imgs])
        den_ptrs = (C.c_void_p * n)(*[d.data_ptr() for d in dens])
        with torch.cuda.device(dev):
            _lib.check(_lib.load().clipebc_sliding_window_predict_batch(
                self._handle, n, img_ptrs, _lib.int_array(hs), _lib.int_array(ws), window_size[0], window_size[1], stride[0],
                stride[1], den_ptrs, cnt.data_ptr(), torch.cuda.current_stream().cuda_stream), "sliding_window_predict_batch")
        return dens, cnt


def _clip_ebc(backbone: str, bins, anchor_points, reduction=None, freeze_text_encoder=True, prompt_type="number",
              input_size=None, num_vpt=None, deep_vpt=None, vpt_drop=None, decoder_block=None, decoder_cfg=None,
              **extra) -> CLIP_EBC:
    """models/clip/model.py:220-270."""
    return CLIP_EBC(backbone=backbone, bins=bins, anchor_points=anchor_points, reduction=reduction,
                    freeze_text_encoder=freeze_text_encoder, prompt_type=prompt_type, input_size=input_size,
                    num_vpt=num_vpt, deep_vpt=deep_vpt, vpt_drop=vpt_drop, decoder_block=decoder_block,
                    decoder_cfg=decoder_cfg, **extra)


def get_model(backbone: str, input_size: int, reduction: int, bins: Optional[List[Tuple[float, float]]] = None,
              anchor_points: Optional[List[float]] = None, **kwargs: Any) -> CLIP_EBC:
    """models/__init__.py:10-44, CLIP branch. kwargs: prompt_type, num_vpt, vpt_drop, deep_vpt (+ text_features)."""
    backbone = backbone.lower()
    if "clip" in backbone:
        backbone = backbone[5:]
        assert backbone in clip_names, f"Expected backbone to be in {clip_names}, got {backbone}"
        return _clip_ebc(backbone=backbone, input_size=input_size, reduction=reduction, bins=bins,
                         anchor_points=anchor_points, **kwargs)
    raise NotImplementedError(
        f"backbone '{backbone}': clip_ebc_b200 provides the CLIP-EBC path only; the reference's non-CLIP Classifier/"
        "Regressor models are outside the scope of this build")
