"""Single-kernel parity tests (B200 only): each sm_100a kernel, called through the C-ABI, against a plain PyTorch
fp32 statement of the same op on the same seeded inputs. Tolerances are written next to each check."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from clip_ebc_b200 import ops as _ops

    return _ops


def _rand(shape, seed, scale=1.0, device="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(device)


def _bf(x):
    return x.to(torch.bfloat16)


DT16 = [torch.bfloat16, torch.float16]
DT16_IDS = ["bf16", "fp16"]
# relative size of one rounding of a value to the 16-bit format (half an ulp of the worst-case mantissa)
ROUND16 = {torch.bfloat16: 2.0 ** -8, torch.float16: 2.0 ** -11}


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20)).item()


# ----------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("block_n", [0, 128, 192, 256])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 768), (300, 768, 768), (1000, 2304, 768),
                                   (777, 768, 3072), (32, 2304, 768), (12608, 768, 768)])
@pytest.mark.parametrize("dt", DT16, ids=DT16_IDS)
def test_gemm_plain_f32(ops, M, N, K, block_n, dt):
    if block_n and N % block_n:
        pytest.skip("N is not a multiple of this tile width")
    a, w = _rand((M, K), 1).to(dt), _rand((N, K), 2, 0.05).to(dt)
    out = ops.gemm(a, w, ops.EPI_F32, block_n=block_n)
    ref = a.float() @ w.float().t()
    # the 16-bit inputs are exact in both; only fp32 accumulation order differs -> 1e-5 relative to the output scale
    assert _rel(out, ref) < 2e-5


@pytest.mark.parametrize("dt", DT16, ids=DT16_IDS)
@pytest.mark.parametrize("block_n", [128, 192, 256])
def test_gemm_bias_16_and_gelu(ops, block_n, dt):
    M, N, K = 1234, 3072, 768
    a, w, b = _rand((M, K), 3).to(dt), _rand((N, K), 4, 0.04).to(dt), _rand((N,), 5)
    ref = a.float() @ w.float().t() + b
    out = ops.gemm(a, w, ops.EPI_BIAS_BF16, bias=b, block_n=block_n)
    assert out.dtype == dt
    assert _rel(out, ref) < 1.5 * ROUND16[dt]  # one rounding of the output
    out = ops.gemm(a, w, ops.EPI_BIAS_BF16, bias=b, block_n=block_n, out_fp16=False)  # QKV: bf16 out for the attention
    assert out.dtype == torch.bfloat16 and _rel(out, ref) < 1.5 * ROUND16[torch.bfloat16]
    out = ops.gemm(a, w, ops.EPI_BIAS_GELU_BF16, bias=b, block_n=block_n)
    refg = ref * torch.sigmoid(1.702 * ref)
    assert _rel(out, refg) < 1.5 * ROUND16[dt] + 2e-6


@pytest.mark.parametrize("block_n", [128, 192, 256])
def test_gemm_bias_resid_inplace(ops, block_n):
    M, N, K = 2000, 768, 3072
    a, w, b = _bf(_rand((M, K), 6)), _bf(_rand((N, K), 7, 0.02)), _rand((N,), 8)
    x = _rand((M, N), 9)
    ref = x + a.float() @ w.float().t() + b
    out = ops.gemm(a, w, ops.EPI_BIAS_RESID_F32, bias=b, resid=x, out=x, block_n=block_n)  # in place on the residual stream
    assert out.data_ptr() == x.data_ptr()
    assert _rel(out, ref) < 2e-5


# ------------------------------------------------------------------- LayerNorm folded into the GEMMs either side of it


@pytest.mark.parametrize("dt", DT16, ids=DT16_IDS)
def test_gemm_conv_segments_match_conv2d(ops, dt):
    """3x3 conv over the zero-bordered NHWC grid as 9 row-shifted K-segments == F.conv2d(padding=1)."""
    B, g, Cc = 3, 14, 768
    Hp = Wp = g + 2
    x = _rand((B, Cc, g, g), 10).to(dt)
    wt = _rand((Cc, Cc, 3, 3), 11, 0.02).to(dt)
    bias = _rand((Cc,), 12)
    ref = F.relu(F.conv2d(x.float(), wt.float(), padding=1) + bias.view(1, -1, 1, 1))
    xp = torch.zeros((B, Hp, Wp, Cc), dtype=dt, device="cuda")
    xp[:, 1:-1, 1:-1, :] = x.permute(0, 2, 3, 1)
    wk = wt.permute(0, 2, 3, 1).reshape(Cc, 9 * Cc).contiguous()  # [O, (ky, kx, I)]
    shifts = [(ky - 1) * Wp + (kx - 1) for ky in range(3) for kx in range(3)]
    out = ops.gemm(xp.view(-1, Cc), wk, ops.EPI_BIAS_RELU_MASK_BF16, bias=bias, K=9 * Cc, seg_row_shift=shifts,
                   seg_col_start=[0] * 9, mask_hw=(Hp, Wp))
    out = out.view(B, Hp, Wp, Cc)
    # border rows must be exactly zero so the result can feed the next conv
    border = out.clone()
    border[:, 1:-1, 1:-1, :] = 0
    assert border.abs().max().item() == 0.0
    got = out[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2)
    assert _rel(got, ref) < 1.5 * ROUND16[dt]  # output rounding


@pytest.mark.parametrize("dt", DT16, ids=DT16_IDS)
def test_gemm_resid_relu_split_and_split_projection(ops, dt):
    """conv2-style epilogue writes hi|lo; the [hi|lo|hi] x [Whi|Whi|Wlo] GEMM reproduces an fp32 linear to ~1e-5."""
    M, Cc, E = 900, 768, 512
    a, w, b = _rand((M, Cc), 13).to(dt), _rand((Cc, Cc), 14, 0.03).to(dt), _rand((Cc,), 15)
    u = _rand((M, Cc), 16)
    t_ref = F.relu(a.float() @ w.float().t() + b + u)
    split = ops.gemm(a, w, ops.EPI_BIAS_RESID_RELU_SPLIT, bias=b, resid=u)
    hi, lo = split[:, :Cc].float(), split[:, Cc:].float()
    assert _rel(hi + lo, t_ref) < 3e-5  # hi + lo carries ~16 mantissa bits
    wp = _rand((E, Cc), 17, 0.05)
    bp = _rand((E,), 18)
    w_hi = wp.to(dt)
    w_lo = (wp - w_hi.float()).to(dt)
    w3 = torch.cat([w_hi, w_hi, w_lo], dim=1).contiguous()
    out = ops.gemm(split, w3, ops.EPI_BIAS_F32, bias=bp, K=3 * Cc, seg_row_shift=[0, 0, 0],
                   seg_col_start=[0, Cc, 0])
    ref = (hi + lo).double() @ wp.double().t() + bp.double()
    assert _rel(out, ref.float()) < 5e-5


@pytest.mark.parametrize("dt", DT16, ids=DT16_IDS)
@pytest.mark.parametrize("rem", [8, 64, 100, 128])
def test_gemm_remainder_rows_all_epilogues(ops, rem, dt):
    """M = 2 full 256-row tiles + a small remainder: every fused epilogue must mask the padded rows of the last tile and
    give the same answer on the remainder rows as on full tiles."""
    M, N, K = 512 + rem, 768, 768
    a, w, b = _rand((M, K), 30).to(dt), _rand((N, K), 31, 0.04).to(dt), _rand((N,), 32)
    x = _rand((M, N), 33)
    ref = a.float() @ w.float().t() + b
    out = ops.gemm(a, w, ops.EPI_BIAS_F32, bias=b)
    assert _rel(out, ref) < 2e-5 and _rel(out[512:], ref[512:]) < 2e-5
    out = ops.gemm(a, w, ops.EPI_BIAS_BF16, bias=b)
    assert _rel(out[512:], ref[512:]) < 1.5 * ROUND16[dt]
    out = ops.gemm(a, w, ops.EPI_BIAS_GELU_BF16, bias=b)
    refg = ref * torch.sigmoid(1.702 * ref)
    assert _rel(out, refg) < 1.5 * ROUND16[dt] + 2e-6 and _rel(out[512:], refg[512:]) < 1.5 * ROUND16[dt] + 2e-6
    xr = x.clone()
    out = ops.gemm(a, w, ops.EPI_BIAS_RESID_F32, bias=b, resid=xr, out=xr)
    assert _rel(out, ref + x) < 2e-5 and _rel(out[512:], (ref + x)[512:]) < 2e-5
    split = ops.gemm(a, w, ops.EPI_BIAS_RESID_RELU_SPLIT, bias=b, resid=x)
    t_ref = F.relu(ref + x)
    got = split[:, :N].float() + split[:, N:].float()
    assert _rel(got, t_ref) < 3e-5 and _rel(got[512:], t_ref[512:]) < 3e-5


@pytest.mark.parametrize("dt", DT16, ids=DT16_IDS)
def test_gemm_conv_segments_on_shared_border_grid(ops, dt):
    """The decoder's layout: one trailing zero column per line and one zero row per image ((g+1)^2 rows instead of
    (g+2)^2); rows before the buffer are zero-filled by TMA. 3x3 conv as 9 row-shifted K-segments == F.conv2d(padding=1)."""
    B, g, Cc = 3, 14, 768
    Wp = g + 1
    x = _rand((B, Cc, g, g), 20).to(dt)
    wt = _rand((Cc, Cc, 3, 3), 21, 0.02).to(dt)
    bias = _rand((Cc,), 22)
    ref = F.relu(F.conv2d(x.float(), wt.float(), padding=1) + bias.view(1, -1, 1, 1))
    xp = torch.zeros((B, Wp, Wp, Cc), dtype=dt, device="cuda")
    xp[:, :-1, :-1, :] = x.permute(0, 2, 3, 1)
    wk = wt.permute(0, 2, 3, 1).reshape(Cc, 9 * Cc).contiguous()
    shifts = [(ky - 1) * Wp + (kx - 1) for ky in range(3) for kx in range(3)]
    out = ops.gemm(xp.view(-1, Cc), wk, ops.EPI_BIAS_RELU_MASK_BF16, bias=bias, K=9 * Cc, seg_row_shift=shifts,
                   seg_col_start=[0] * 9, mask_hw=(Wp, Wp), mask_lead=False).view(B, Wp, Wp, Cc)
    border = out.clone()
    border[:, :-1, :-1, :] = 0
    assert border.abs().max().item() == 0.0
    assert _rel(out[:, :-1, :-1, :].permute(0, 3, 1, 2), ref) < 1.5 * ROUND16[dt]


def test_gemm_remainder_rows_conv_segments(ops):
    """Row-shifted K-segments + border mask on a shape with a remainder (2 windows x 30 x 30 = 1800 = 7 x 256 + 8)."""
    dt = torch.float16
    B, g, Cc = 2, 28, 768
    Hp = Wp = g + 2
    x = _rand((B, Cc, g, g), 40).to(dt)
    wt = _rand((Cc, Cc, 3, 3), 41, 0.02).to(dt)
    bias = _rand((Cc,), 42)
    ref = F.relu(F.conv2d(x.float(), wt.float(), padding=1) + bias.view(1, -1, 1, 1))
    xp = torch.zeros((B, Hp, Wp, Cc), dtype=dt, device="cuda")
    xp[:, 1:-1, 1:-1, :] = x.permute(0, 2, 3, 1)
    wk = wt.permute(0, 2, 3, 1).reshape(Cc, 9 * Cc).contiguous()
    shifts = [(ky - 1) * Wp + (kx - 1) for ky in range(3) for kx in range(3)]
    out = ops.gemm(xp.view(-1, Cc), wk, ops.EPI_BIAS_RELU_MASK_BF16, bias=bias, K=9 * Cc, seg_row_shift=shifts,
                   seg_col_start=[0] * 9, mask_hw=(Hp, Wp)).view(B, Hp, Wp, Cc)
    border = out.clone()
    border[:, 1:-1, 1:-1, :] = 0
    assert border.abs().max().item() == 0.0
    assert _rel(out[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2), ref) < 1.5 * ROUND16[dt]


@pytest.mark.parametrize("dt", DT16, ids=DT16_IDS)
def test_gemm_resid16_relu_mask(ops, dt):
    """Epilogue 9 (last 1x1 conv of a ResNet bottleneck without a downsample conv): relu(acc + bias + identity) on the interior
    cells of a shared-border grid, zero on its border row / column; the identity is a 16-bit tensor."""
    n, g, cin, cout = 3, 7, 128, 256
    rows = n * (g + 1) * (g + 1)
    a = _rand((rows, cin), 90).to(dt)
    w = _rand((cout, cin), 91, 0.1).to(dt)
    bias = _rand((cout,), 92)
    ident = _rand((rows, cout), 93).to(dt)
    out = ops.gemm(a, w, ops.EPI_BIAS_RESID16_RELU_MASK_BF16, bias=bias, resid=ident, mask_hw=(g + 1, g + 1), mask_lead=False)
    ref = torch.relu(a.float() @ w.float().t() + bias + ident.float()).view(n, g + 1, g + 1, cout)
    ref[:, g, :, :] = 0
    ref[:, :, g, :] = 0
    out = out.float().view(n, g + 1, g + 1, cout)
    assert out[:, g].abs().max().item() == 0.0 and out[:, :, g].abs().max().item() == 0.0
    assert _rel(out, ref) < 1.5 * ROUND16[dt]


def test_gemm_rejects_bad_shapes(ops):
    a, w = _bf(_rand((64, 100), 1)), _bf(_rand((256, 100), 2))
    with pytest.raises(RuntimeError):
        ops.gemm(a, w, ops.EPI_F32)  # K not a multiple of 64


# ----------------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("width", [768, 1024], ids=["vit_b", "vit_l"])
def test_layernorm(ops, width):
    x = _rand((1000, width), 20, 3.0) + 0.5
    g, b = _rand((width,), 21) * 0.1 + 1.0, _rand((width,), 22) * 0.1
    ref = F.layer_norm(x, (width,), g, b, 1e-5)
    out = ops.layernorm(x, g, b, out_dtype=torch.float32)
    assert (out - ref).abs().max().item() < 2e-5
    for dt in DT16:
        out16 = ops.layernorm(x, g, b, out_dtype=dt)          # >= 64 contiguous rows: the streaming kernel
        assert torch.equal(out16, out.to(dt))
        out16s = ops.layernorm(x[:40].contiguous(), g, b, out_dtype=dt)  # short input: the warp-per-row kernel
        assert torch.equal(out16s, out[:40].to(dt))
    # row map: take the last 196 rows of every group of 229 (ln_post on the patch rows)
    x = _rand((3 * 229, width), 23)
    out = ops.layernorm(x, g, b, out_dtype=torch.float32, n_rows_out=3 * 196, rows_out_per_group=196, rows_in_per_group=229,
                        in_row_offset=33)
    ref = F.layer_norm(x.view(3, 229, width)[:, 33:], (width,), g, b, 1e-5).reshape(-1, width)
    assert (out - ref).abs().max().item() < 2e-5


# ----------------------------------------------------------------------------------------------- attention
def _attention_ref(qkv, ckv, n_win, t_live, n_const, heads):
    q, k, v = qkv.float().view(n_win, t_live, 3, heads, 64).unbind(2)
    if n_const:
        ck, cv = ckv.float().view(n_const, 3, heads, 64)[:, 1], ckv.float().view(n_const, 3, heads, 64)[:, 2]
        k = torch.cat([k, ck.expand(n_win, -1, -1, -1)], 1)
        v = torch.cat([v, cv.expand(n_win, -1, -1, -1)], 1)
    return F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2)


# (2, 17, 5) and (1, 300, 7): constant-key counts that are not multiples of 8 take the streamed-K/V kernel
@pytest.mark.parametrize("heads", [12, 16], ids=["h12", "h16"])
@pytest.mark.parametrize("n_win,t_live,n_const", [(2, 197, 32), (3, 229, 0), (1, 50, 0), (2, 17, 5), (1, 256, 0),
                                                  (5, 197, 32), (2, 128, 8), (1, 129, 0), (40, 197, 32)])
@pytest.mark.parametrize("out_fp16", [False, True], ids=["out_bf16", "out_fp16"])
def test_attention(ops, n_win, t_live, n_const, out_fp16, heads):
    qkv = _bf(_rand((n_win * t_live, 3 * 64 * heads), 30))
    ckv = _bf(_rand((n_const, 3 * 64 * heads), 31)) if n_const else None
    out = ops.attention(qkv, n_win, t_live, ckv, out_fp16=out_fp16, heads=heads)
    assert out.dtype == (torch.float16 if out_fp16 else torch.bfloat16)
    out = out.float().view(n_win, t_live, heads, 64)
    ref = _attention_ref(qkv, ckv, n_win, t_live, n_const, heads)
    # P is rounded to bf16 before P@V and the output is bf16: 2^-8 relative
    assert (out - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("n_win,t_live,n_const,heads", [(2, 401, 32, 12), (1, 785, 32, 12), (2, 817, 0, 12), (3, 257, 0, 12),
                                                        (1, 300, 7, 12), (3, 257, 32, 16), (2, 289, 0, 16), (1, 288, 32, 12),
                                                        (3, 300, 16, 12), (2, 225, 32, 12), (1, 320, 0, 12), (40, 257, 32, 16),
                                                        (1, 300, 24, 12), (7, 193, 64, 12), (5, 321, 0, 12), (3, 264, 32, 16),
                                                        (2, 136, 128, 12)])
def test_attention_beyond_256_tokens(ops, n_win, t_live, n_const, heads):
    """Windows with more than 256 keys. 257..320 keys with a constant-key count that is a multiple of 16 (ViT-L/14 224 x 224:
    257 + 32 deep, 289 shallow; the 320-key maximum; 1 and 64 keys in the second block; two and three query tiles; enough
    windows for the K / V stages and the query ring to wrap; 1 and 8 query rows in the last tile) take the two-block tcgen05 kernel (attention_ppl.cu); everything
    else (ViT-B/16 320 x 320: 401 live + 32 prompt keys; 448 x 448: 785 + 32; 817 live; 321 keys; prompt counts 7 and 24) the
    streamed-K/V kernel (64-query chunks, 64-key blocks, online softmax)."""
    qkv = _bf(_rand((n_win * t_live, 3 * 64 * heads), 34, 1.5))
    ckv = _bf(_rand((n_const, 3 * 64 * heads), 35, 1.5)) if n_const else None
    out = ops.attention(qkv, n_win, t_live, ckv, out_fp16=True, heads=heads).float().view(n_win, t_live, heads, 64)
    ref = _attention_ref(qkv, ckv, n_win, t_live, n_const, heads)
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() < 2e-2 * ref.abs().max().item()  # P rounded to bf16, 16-bit output


def test_attention_two_block_kernel_is_deterministic_and_batch_invariant(ops):
    """The two-block kernel gives a window the same bits whatever batch it is part of (work is dealt per CTA in tile ranges
    that depend on the batch) and from call to call: the cross-image batching of the evaluation loop relies on it."""
    heads, t_live, n_const = 16, 257, 32
    qkv = _bf(_rand((24 * t_live, 3 * 64 * heads), 36, 1.5))
    ckv = _bf(_rand((n_const, 3 * 64 * heads), 37, 1.5))
    full = ops.attention(qkv, 24, t_live, ckv, out_fp16=True, heads=heads).view(24, t_live, heads * 64)
    again = ops.attention(qkv, 24, t_live, ckv, out_fp16=True, heads=heads).view(24, t_live, heads * 64)
    assert torch.equal(full, again)
    for w0, n in ((0, 1), (5, 3), (17, 7)):
        part = ops.attention(qkv[w0 * t_live:(w0 + n) * t_live].contiguous(), n, t_live, ckv, out_fp16=True, heads=heads)
        assert torch.equal(part.view(n, t_live, heads * 64), full[w0:w0 + n])


def test_attention_two_block_kernel_large_scores(ops):
    """Peaky softmax over 289 keys: the reference maximum of the first block also scales the second one."""
    n_win, t_live, n_const, heads = 2, 257, 32, 16
    qkv = _bf(_rand((n_win * t_live, 3 * 64 * heads), 38, 3.0))
    ckv = _bf(_rand((n_const, 3 * 64 * heads), 39, 3.0))
    out = ops.attention(qkv, n_win, t_live, ckv, heads=heads).float().view(n_win, t_live, heads, 64)
    ref = _attention_ref(qkv, ckv, n_win, t_live, n_const, heads)
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


def test_attention_large_scores(ops):
    """Peaky softmax (|score| up to ~40): the single-pass reference-max scheme must stay exact up to bf16 rounding."""
    n_win, t_live, n_const = 3, 197, 32
    qkv = _bf(_rand((n_win * t_live, 2304), 32, 3.0))
    ckv = _bf(_rand((n_const, 2304), 33, 3.0))
    out = ops.attention(qkv, n_win, t_live, ckv).float().view(n_win, t_live, 12, 64)
    ref = _attention_ref(qkv, ckv, n_win, t_live, n_const, 12)
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


# ----------------------------------------------------------------------------------------------- stem / decoder
@pytest.mark.parametrize("dt", DT16, ids=DT16_IDS)
@pytest.mark.parametrize("patch", [16, 32, 14])
def test_patchify_matches_unfold(ops, dt, patch):
    H, W = (64, 96) if patch != 14 else (56, 98)
    img = _rand((2, 3, H, W), 40)
    k = 3 * patch * patch
    kp = (k + 63) // 64 * 64   # ViT-L/14: 588 real columns padded to 640 (the GEMM's K granularity), pad columns zero
    out = ops.patchify(img, fp16=dt == torch.float16, patch=patch).float().view(2, (H // patch) * (W // patch), 2, kp)
    assert out[..., k:].abs().max().item() == 0.0 if kp > k else True
    out = out[..., :k]
    ref = F.unfold(img, kernel_size=patch, stride=patch).transpose(1, 2)  # [n, L, c*P*P + py*P + px]
    hi = ref.to(dt).float()
    assert torch.equal(out[:, :, 0], hi)
    assert torch.equal(out[:, :, 1], (ref - hi).to(dt).float())
    # hi + lo carries the pixels to ~2 roundings of the 16-bit format
    assert (out[:, :, 0] + out[:, :, 1] - ref).abs().max().item() < 4 * ROUND16[dt] ** 2 * ref.abs().max().item()
    # the single-precision form (fp16 operands): hi only, rows of kp columns
    single = ops.patchify(img, fp16=dt == torch.float16, patch=patch, split=False).float()
    assert single.shape == (2 * (H // patch) * (W // patch), kp)
    assert torch.equal(single.view(2, -1, kp)[..., :k], hi)


@pytest.mark.parametrize("g", [28, 14, 7])
def test_resample_from_the_7x7_grid_of_vit_b_32(ops, g):
    """ViT-B/32 windows have 7 x 7 patches: x4 / x2 bilinear resample (or none) onto the reduction-8 / 16 / 32 grid."""
    n = 3
    Y = _rand((n * 49, 768), 42)
    ub, uf = ops.resample_to_padded(Y, n, 7, 7, g, g, fp16=True)
    x = Y.view(n, 7, 7, 768).permute(0, 3, 1, 2)
    ref = x if g == 7 else F.interpolate(x, scale_factor=g / 7, mode="bilinear")
    uf = uf.view(n, g + 1, g + 1, 768)
    assert (uf[:, :-1, :-1].permute(0, 3, 1, 2) - ref).abs().max().item() < 1e-5
    assert torch.equal(ub.view_as(uf), uf.to(torch.float16))


def test_resample_x1_75_of_vit_l_14(ops):
    """ViT-L/14: 16 x 16 patches of 1024 channels -> the 28 x 28 reduction-8 grid, scale_factor 14 / 8 = 1.75."""
    n = 2
    Y = _rand((n * 256, 1024), 43)
    ub, uf = ops.resample_to_padded(Y, n, 16, 16, 28, 28, fp16=True)
    x = Y.view(n, 16, 16, 1024).permute(0, 3, 1, 2)
    ref = F.interpolate(x, scale_factor=14 / 8, mode="bilinear")
    assert ref.shape[-2:] == (28, 28)
    uf = uf.view(n, 29, 29, 1024)
    assert (uf[:, :-1, :-1].permute(0, 3, 1, 2) - ref).abs().max().item() < 1e-5
    assert torch.equal(ub.view_as(uf), uf.to(torch.float16))


@pytest.mark.parametrize("g", [28, 14, 7])
@pytest.mark.parametrize("dt", DT16, ids=DT16_IDS)
def test_resample_matches_interpolate(ops, g, dt):
    n = 2
    Y = _rand((n * 196, 768), 41)
    ub, uf = ops.resample_to_padded(Y, n, 14, 14, g, g, fp16=dt == torch.float16)
    x = Y.view(n, 14, 14, 768).permute(0, 3, 1, 2)
    ref = x if g == 14 else F.interpolate(x, scale_factor=g / 14, mode="bilinear")
    uf = uf.view(n, g + 1, g + 1, 768)  # shared-border grid: one trailing zero column per line, one zero row per window
    assert (uf[:, :-1, :-1].permute(0, 3, 1, 2) - ref).abs().max().item() < 1e-5
    border = uf.clone()
    border[:, :-1, :-1] = 0
    assert border.abs().max().item() == 0.0
    assert torch.equal(ub.view_as(uf), uf.to(dt))


# ----------------------------------------------------------------------------------------------- fold
def _numpy_fold(preds, H, W, wh, ww, sh, sw, r):
    """The reference loop (utils/eval_utils.py:54-95) verbatim in numpy, for bit-exact comparison."""
    import numpy as np

    nr = int(np.ceil((H - wh) / sh) + 1)
    nc = int(np.ceil((W - ww) / sw) + 1)
    pm = np.zeros((1, H // r, W // r), np.float32)
    cm = np.zeros((1, H // r, W // r), np.float32)
    idx = 0
    for i in range(nr):
        for j in range(nc):
            xs, ys = i * sh, j * sw
            xe, ye = xs + wh, ys + ww
            if xe > H:
                xs, xe = H - wh, H
            if ye > W:
                ys, ye = W - ww, W
            pm[:, xs // r: xe // r, ys // r: ye // r] += preds[idx]
            cm[:, xs // r: xe // r, ys // r: ye // r] += 1.0
            idx += 1
    return pm / cm


@pytest.mark.parametrize("H,W,s,r", [(448, 448, 224, 8), (448, 672, 112, 8), (1536, 2048, 112, 8), (448, 672, 112, 32),
                                     (480, 700, 100, 16)])
def test_fold_bit_exact(ops, H, W, s, r):
    ro, co = ops.window_origins(H, W, (224, 224), (s, s))
    g = 224 // r
    preds = _rand((len(ro) * len(co), 1, g, g), 60).abs()
    dens, cnt = ops.fold_average(preds, [v // r for v in ro], [v // r for v in co], H // r, W // r, want_count=True)
    ref = _numpy_fold(preds.cpu().numpy(), H, W, 224, 224, s, s, r)[0]
    import numpy as np

    assert np.array_equal(dens.cpu().numpy(), ref)  # bit-exact: same fp32 additions in the same order
    assert abs(cnt.item() - float(ref.sum(dtype=np.float64))) < 1e-3 * max(1.0, abs(float(ref.sum())))
