"""Single-kernel parity tests (B200 only): each sm_100a kernel, called through the C-ABI, against a plain PyTorch
fp32 statement of the same op on the same seeded inputs. Tolerances are written next to each check."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[1, 2], ids=["gemm1cta", "gemm2cta"])
def ops(request):
    """All kernel tests run once per GEMM implementation (single CTA / CTA pair with cta_group::2)."""
    from clip_ebc_b200 import ops as _ops

    _ops.set_gemm_impl(request.param)
    yield _ops
    _ops.set_gemm_impl(2)


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
def _ln_ref(x, gamma, beta):
    """nn.LayerNorm(768, eps=1e-5) in float64 (reference: _clip/blocks.py:8-14)."""
    xd = x.double()
    mu = xd.mean(-1, keepdim=True)
    var = ((xd - mu) ** 2).mean(-1, keepdim=True)
    return ((xd - mu) / torch.sqrt(var + 1e-5) * gamma.double() + beta.double()).float()


def _merge_stats(stats, parts):
    """Chan merge of the per-row partials (equal counts), in float64 -> (mean, variance)."""
    st = stats[:, :parts].double()
    n = 768 // parts
    mean = st[..., 0].mean(1)
    m2 = st[..., 1].sum(1) + n * ((st[..., 0] - mean[:, None]) ** 2).sum(1)
    return mean, m2 / 768


@pytest.mark.parametrize("dt", DT16, ids=DT16_IDS)
def test_rowstats_and_fold_ln_linear(dt):
    from clip_ebc_b200 import ops as _ops

    fp16 = dt == torch.float16
    x = _rand((333, 768), 60, 2.0) + 3.0 * _rand((333, 1), 61)   # rows with a non-zero mean
    x16, stats = _ops.rowstats(x, fp16=fp16)
    assert torch.equal(x16, x.to(dt))
    mean, var = _merge_stats(stats, 1)
    assert (mean - x.double().mean(1)).abs().max().item() < 1e-5
    assert ((var - x.double().var(1, unbiased=False)).abs() / x.double().var(1, unbiased=False)).max().item() < 1e-5
    W, b = _rand((2304, 768), 62, 0.03), _rand((2304,), 63)
    gamma, beta = 1.0 + 0.3 * _rand((768,), 64), 0.2 * _rand((768,), 65)
    wf, colsum, bias_f = _ops.fold_ln_linear(W, b, gamma, beta, fp16=fp16)
    assert torch.equal(wf, (W * gamma).to(dt))
    assert (colsum - wf.double().sum(1)).abs().max().item() < 1e-4   # sums of the ROUNDED weights
    assert (bias_f - (b.double() + W.double() @ beta.double())).abs().max().item() < 1e-4


@pytest.mark.parametrize("block_n", [0, 192])
@pytest.mark.parametrize("M,K", [(2000, 768), (777, 3072), (12608, 768), (31, 768)])
def test_gemm_resid_stats(block_n, M, K):
    """out_proj / c_proj with the statistics epilogue (residual tiles through TMA): same residual update as epilogue 4 plus
    the 16-bit rows and the (mean, M2) partials of every 96 columns."""
    from clip_ebc_b200 import ops as _ops

    dt = torch.float16
    a, w, b = _rand((M, K), 66).to(dt), _rand((768, K), 67, 0.02).to(dt), _rand((768,), 68)
    x = _rand((M, 768), 69) + 2.0
    ref = x + a.float() @ w.float().t() + b
    x16, stats = _ops.gemm_resid_stats(a, w, x, b, block_n=block_n)
    assert _rel(x, ref) < 2e-5                                   # in place, as EPI_BIAS_RESID_F32
    assert torch.equal(x16, x.to(dt))                            # the 16-bit copy is the rounding of what was written
    xc = x.double().view(M, 8, 96)
    assert (stats[..., 0].double() - xc.mean(2)).abs().max().item() < 2e-5              # partial means
    assert _rel(stats[..., 1], ((xc - xc.mean(2, keepdim=True)) ** 2).sum(2)) < 2e-5      # partial M2
    mean, var = _merge_stats(stats, 8)
    xd = x.double()
    assert (mean - xd.mean(1)).abs().max().item() < 2e-5
    assert ((var - xd.var(1, unbiased=False)).abs() / xd.var(1, unbiased=False)).max().item() < 2e-5


@pytest.mark.parametrize("dt", DT16, ids=DT16_IDS)
@pytest.mark.parametrize("gelu", [False, True], ids=["qkv", "c_fc"])
@pytest.mark.parametrize("M,N", [(1234, 2304), (12608, 3072), (50, 768)])
def test_gemm_ln_matches_layernorm_linear(dt, gelu, M, N):
    """LN folded into the GEMM (raw 16-bit rows, W diag(gamma), epilogue rstd * (acc - mean * colsum) + b') against
    LayerNorm -> Linear (-> QuickGELU) in fp32/fp64 on the same rows. The only differences are one 16-bit rounding of the
    raw rows (instead of the normalised ones) and of the folded weights: a few output roundings."""
    from clip_ebc_b200 import ops as _ops

    fp16 = dt == torch.float16
    x = _rand((M, 768), 70, 1.5) + 2.0 * _rand((M, 1), 71) + 0.5 * _rand((1, 768), 72)
    W, b = _rand((N, 768), 73, 0.03), _rand((N,), 74)
    gamma, beta = 1.0 + 0.3 * _rand((768,), 75), 0.2 * _rand((768,), 76)
    ref = _ln_ref(x, gamma, beta) @ W.t() + b
    if gelu:
        ref = ref * torch.sigmoid(1.702 * ref)
    x16, stats = _ops.rowstats(x, fp16=fp16)
    wf, colsum, bias_f = _ops.fold_ln_linear(W, b, gamma, beta, fp16=fp16)
    out = _ops.gemm_ln(x16, wf, bias_f, stats, colsum, 1, gelu=gelu)
    # what the separate-kernel path computes: 16-bit LN output x 16-bit weights
    xn16 = _ln_ref(x, gamma, beta).to(dt).float()
    unfused = xn16 @ W.to(dt).float().t() + b
    if gelu:
        unfused = unfused * torch.sigmoid(1.702 * unfused)
    err, err_unfused = _rel(out, ref), _rel(unfused.to(dt), ref)
    assert err < 4 * ROUND16[dt], (err, err_unfused)
    assert err < 3 * err_unfused + 1e-4, (err, err_unfused)      # no worse than the path it replaces (same error class)


def test_gemm_resid_stats_feeds_gemm_ln():
    """The pair as the hot path chains it: residual GEMM -> (x16, chunk statistics) -> LN-folded GEMM."""
    from clip_ebc_b200 import ops as _ops

    dt = torch.float16
    M = 3000
    a, w, b = _rand((M, 768), 77).to(dt), _rand((768, 768), 78, 0.03).to(dt), _rand((768,), 79)
    x = _rand((M, 768), 80)
    W2, b2 = _rand((3072, 768), 81, 0.03), _rand((3072,), 82)
    gamma, beta = 1.0 + 0.3 * _rand((768,), 83), 0.2 * _rand((768,), 84)
    wf, colsum, bias_f = _ops.fold_ln_linear(W2, b2, gamma, beta, fp16=True)
    xc = x.clone()
    x16, stats = _ops.gemm_resid_stats(a, w, xc, b)
    out = _ops.gemm_ln(x16, wf, bias_f, stats, colsum, 8, gelu=True)
    r = _ln_ref(xc, gamma, beta) @ W2.t() + b2
    r = r * torch.sigmoid(1.702 * r)
    assert _rel(out, r) < 4 * ROUND16[dt]


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


def test_gemm_rejects_bad_shapes(ops):
    a, w = _bf(_rand((64, 100), 1)), _bf(_rand((256, 100), 2))
    with pytest.raises(RuntimeError):
        ops.gemm(a, w, ops.EPI_F32)  # K not a multiple of 64


# ----------------------------------------------------------------------------------------------- LayerNorm
def test_layernorm(ops):
    x = _rand((1000, 768), 20, 3.0) + 0.5
    g, b = _rand((768,), 21) * 0.1 + 1.0, _rand((768,), 22) * 0.1
    ref = F.layer_norm(x, (768,), g, b, 1e-5)
    out = ops.layernorm(x, g, b, out_dtype=torch.float32)
    assert (out - ref).abs().max().item() < 2e-5
    for dt in DT16:
        out16 = ops.layernorm(x, g, b, out_dtype=dt)
        assert torch.equal(out16, out.to(dt))
    # row map: take the last 196 rows of every group of 229 (ln_post on the patch rows)
    x = _rand((3 * 229, 768), 23)
    out = ops.layernorm(x, g, b, out_dtype=torch.float32, n_rows_out=3 * 196, rows_out_per_group=196, rows_in_per_group=229,
                        in_row_offset=33)
    ref = F.layer_norm(x.view(3, 229, 768)[:, 33:], (768,), g, b, 1e-5).reshape(-1, 768)
    assert (out - ref).abs().max().item() < 2e-5


# ----------------------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("impl", [1, 2, 3, 4], ids=["mma_sync", "tcgen05", "tcgen05_persistent", "tcgen05_two_chains"])
@pytest.mark.parametrize("n_win,t_live,n_const", [(2, 197, 32), (3, 229, 0), (1, 50, 0), (2, 17, 5), (1, 256, 0),
                                                  (5, 197, 32), (2, 128, 8), (1, 129, 0), (40, 197, 32)])
@pytest.mark.parametrize("out_fp16", [False, True], ids=["out_bf16", "out_fp16"])
def test_attention(ops, n_win, t_live, n_const, impl, out_fp16):
    ops.set_attention_impl(impl)  # impl 2 falls back to 1 when n_const % 8 != 0
    qkv = _bf(_rand((n_win * t_live, 2304), 30))
    ckv = _bf(_rand((n_const, 2304), 31)) if n_const else None
    out = ops.attention(qkv, n_win, t_live, ckv, out_fp16=out_fp16)
    assert out.dtype == (torch.float16 if out_fp16 else torch.bfloat16)
    out = out.float().view(n_win, t_live, 12, 64)
    q, k, v = qkv.float().view(n_win, t_live, 3, 12, 64).unbind(2)
    if n_const:
        ck, cv = ckv.float().view(n_const, 3, 12, 64)[:, 1], ckv.float().view(n_const, 3, 12, 64)[:, 2]
        k = torch.cat([k, ck.expand(n_win, -1, -1, -1)], 1)
        v = torch.cat([v, cv.expand(n_win, -1, -1, -1)], 1)
    ref = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2)
    ops.set_attention_impl(4)
    # P is rounded to bf16 before P@V and the output is bf16: 2^-8 relative
    assert (out - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("n_win,t_live,n_const", [(2, 401, 32), (1, 785, 32), (2, 817, 0), (3, 257, 0), (1, 300, 7)])
def test_attention_beyond_256_tokens(n_win, t_live, n_const):
    """Windows with more than 256 tokens (320 x 320: 401 live + 32 prompt keys; 448 x 448: 785 + 32; shallow VPT: all 817
    live) go through the streamed-K/V kernel (64-query chunks, 64-key blocks, online softmax)."""
    from clip_ebc_b200 import ops as _ops

    qkv = _bf(_rand((n_win * t_live, 2304), 34, 1.5))
    ckv = _bf(_rand((n_const, 2304), 35, 1.5)) if n_const else None
    out = _ops.attention(qkv, n_win, t_live, ckv, out_fp16=True).float().view(n_win, t_live, 12, 64)
    q, k, v = qkv.float().view(n_win, t_live, 3, 12, 64).unbind(2)
    if n_const:
        ck, cv = ckv.float().view(n_const, 3, 12, 64)[:, 1], ckv.float().view(n_const, 3, 12, 64)[:, 2]
        k = torch.cat([k, ck.expand(n_win, -1, -1, -1)], 1)
        v = torch.cat([v, cv.expand(n_win, -1, -1, -1)], 1)
    ref = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2)
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() < 2e-2 * ref.abs().max().item()  # P rounded to bf16, 16-bit output


@pytest.mark.parametrize("impl", [1, 2, 3, 4], ids=["mma_sync", "tcgen05", "tcgen05_persistent", "tcgen05_two_chains"])
def test_attention_large_scores(ops, impl):
    """Peaky softmax (|score| up to ~40): the single-pass reference-max scheme must stay exact up to bf16 rounding."""
    ops.set_attention_impl(impl)
    n_win, t_live, n_const = 3, 197, 32
    qkv = _bf(_rand((n_win * t_live, 2304), 32, 3.0))
    ckv = _bf(_rand((n_const, 2304), 33, 3.0))
    out = ops.attention(qkv, n_win, t_live, ckv).float().view(n_win, t_live, 12, 64)
    ops.set_attention_impl(4)
    q, k, v = qkv.float().view(n_win, t_live, 3, 12, 64).unbind(2)
    ck, cv = ckv.float().view(n_const, 3, 12, 64)[:, 1], ckv.float().view(n_const, 3, 12, 64)[:, 2]
    k = torch.cat([k, ck.expand(n_win, -1, -1, -1)], 1)
    v = torch.cat([v, cv.expand(n_win, -1, -1, -1)], 1)
    ref = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2)
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


# ----------------------------------------------------------------------------------------------- stem / decoder
@pytest.mark.parametrize("dt", DT16, ids=DT16_IDS)
@pytest.mark.parametrize("patch", [16, 32])
def test_patchify_matches_unfold(ops, dt, patch):
    img = _rand((2, 3, 64, 96), 40)
    kp = 3 * patch * patch
    out = ops.patchify(img, fp16=dt == torch.float16, patch=patch).float().view(2, (64 // patch) * (96 // patch), 2, kp)
    ref = F.unfold(img, kernel_size=patch, stride=patch).transpose(1, 2)  # [n, L, c*P*P + py*P + px]
    hi = ref.to(dt).float()
    assert torch.equal(out[:, :, 0], hi)
    assert torch.equal(out[:, :, 1], (ref - hi).to(dt).float())
    # hi + lo carries the pixels to ~2 roundings of the 16-bit format
    assert (out[:, :, 0] + out[:, :, 1] - ref).abs().max().item() < 4 * ROUND16[dt] ** 2 * ref.abs().max().item()


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


@pytest.mark.parametrize("n_bins", [3, 5, 20])
def test_ebc_head(ops, n_bins):
    n, g = 2, 7
    Fm = _rand((n * (g + 1) * (g + 1), 512), 50)
    text = _rand((n_bins, 512), 51)
    anchors = torch.arange(n_bins, dtype=torch.float32, device="cuda") * 1.25
    scale = math.log(1 / 0.07)
    tmat = math.exp(scale) * F.normalize(text, dim=-1)
    exp, logits = ops.ebc_head(Fm, tmat.contiguous(), anchors, n, g, g, want_logits=True)
    f = Fm.view(n, g + 1, g + 1, 512)[:, :-1, :-1]
    ref_logits = (math.exp(scale) * F.normalize(f, dim=-1)) @ F.normalize(text, dim=-1).t()
    ref_logits = ref_logits.permute(0, 3, 1, 2)
    ref_exp = (ref_logits.softmax(1) * anchors.view(1, -1, 1, 1)).sum(1, keepdim=True)
    assert (logits - ref_logits).abs().max().item() < 1e-4
    assert (exp - ref_exp).abs().max().item() < 1e-4


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
