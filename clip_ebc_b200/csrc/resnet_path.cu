// The CLIP-ResNet encoder path behind the same boundary: CLIP_EBC with a ModifiedResNet image
// encoder (resnet50, resnet101: stem width 64; resnet50x4 / x16 / x64: 80 / 96 / 128) and its Bottleneck decoder.
//
// Reference: /root/reference/models/clip/_clip/image_encoder.py:10-115 (stem: three 3x3 convs + avgpool; four layers of
// Bottlenecks; layer4 keeps stride 1 when reduction <= 16), _clip/blocks.py:56-101 (Bottleneck: 1x1 -> 3x3 -> avgpool(stride) ->
// 1x1, downsample = avgpool(stride) -> 1x1 -> BN), models/clip/model.py:50-52,193-198,228-239 (features -> bilinear resample to
// the reduction grid -> decoder Bottleneck(s) of models/utils.py:334-390 -> 1x1 projection -> EBC head).
//
// Design. Activations are 16-bit NHWC on shared-border grids (kernels.h), so that EVERY convolution is a launch of the
// tcgen05 GEMM: 1x1 convs are plain GEMMs over the rows, 3x3 convs implicit GEMMs with 9 row-shifted K-segments, the stride-2
// stem conv a GEMM over im2col rows; BatchNorm is folded into the weights, ReLU and the zero border into the epilogue. The
// identity branch is either the 16-bit residual of the epilogue (EPI_BIAS_RESID16_RELU_MASK_BF16) or -- when the block has a
// downsample conv -- folded into the SAME GEMM: conv3 and the downsample conv are both 1x1 on the block's output grid, so
// their operands are concatenated along K ([conv2 output | (pooled) block input] x [W3' | Wd']) and their biases added.
// Anti-aliased strides are average pools between grids (resnet.cu). Channel counts are zero-padded to the GEMM's granularity:
// a tensor of C channels lives in a buffer of n128(C) columns (a multiple of 128, what a GEMM writes) whose columns beyond C
// are exactly zero (zero weight rows, zero bias, relu(0) = 0), and is read as a K-segment of k64(C) columns.
#include <algorithm>
#include <string>

#include "model_state.h"

namespace cebc {

namespace {

int k64(int c) { return (c + 63) / 64 * 64; }      // channels as a K-segment of the GEMM
int n128(int c) { return (c + 127) / 128 * 128; }  // channels as the N of a GEMM = pitch of the buffer it writes

bool has(clipebc_model* m, const std::string& name) { return m->raw.find(name) != m->raw.end(); }
int64_t dim(clipebc_model* m, const std::string& name, int i) {
  auto it = m->raw.find(name);
  if (it == m->raw.end() || static_cast<int>(it->second.shape.size()) <= i) return -1;
  return it->second.shape[i];
}

// conv `wname` [O, I, k, k] + BatchNorm `bn` -> cp (or, with col_off > 0 / accumulate, into the columns behind an earlier fold)
int fold(clipebc_model* m, cudaStream_t s, ConvPack* cp, const std::string& wname, const std::string& bn, int O, int I, int taps,
         int i_pad, int n_pad, int k_total, int col_off, bool accumulate, std::string* err) {
  const int ksz = taps == 9 ? 3 : 1;
  bool ok = taps == 1 && ksz == 1 && wname == "image_encoder.conv1.weight"
                ? true
                : check_shape(m, wname, {O, I, ksz, ksz}, err);
  ok = ok && check_shape(m, bn + ".weight", {O}, err) && check_shape(m, bn + ".bias", {O}, err) &&
       check_shape(m, bn + ".running_mean", {O}, err) && check_shape(m, bn + ".running_var", {O}, err);
  if (!ok) return fail(CLIPEBC_ESTATE, "pack: " + *err);
  const int fp16 = m->cfg.operand_fp16 != 0;
  if (!accumulate) {
    cp->n = O; cp->n_pad = n_pad; cp->k = k_total;
    CUDA_TRY(cp->w.reserve(static_cast<size_t>(n_pad) * k_total * 2));
    CUDA_TRY(cp->b.reserve(static_cast<size_t>(n_pad) * 4));
    CUDA_TRY(cudaMemsetAsync(cp->w.p, 0, static_cast<size_t>(n_pad) * k_total * 2, s));
    CUDA_TRY(cudaMemsetAsync(cp->b.p, 0, static_cast<size_t>(n_pad) * 4, s));
  }
  K_TRY(fold_conv_bn_general(s, raw_ptr(m, wname), raw_ptr(m, bn + ".weight"), raw_ptr(m, bn + ".bias"),
                             raw_ptr(m, bn + ".running_mean"), raw_ptr(m, bn + ".running_var"), 1e-5f, O, I, taps, i_pad, cp->w.p,
                             k_total, col_off, cp->b.as<float>(), accumulate ? 1 : 0, fp16));
  return CLIPEBC_OK;
}

int pack_block(clipebc_model* m, cudaStream_t s, RnBlock* b, std::string* err) {
  const std::string& p = b->prefix;
  const int np = n128(b->planes), kp = k64(b->planes), kin = k64(b->c_in);
  int rc;
  if ((rc = fold(m, s, &b->c1, p + "conv1.weight", p + "bn1", b->planes, b->c_in, 1, kin, np, kin, 0, false, err))) return rc;
  if ((rc = fold(m, s, &b->c2, p + "conv2.weight", p + "bn2", b->planes, b->planes, 9, kp, np, 9 * kp, 0, false, err))) return rc;
  const int k3 = kp + (b->down ? kin : 0);
  if ((rc = fold(m, s, &b->c3, p + "conv3.weight", p + "bn3", b->c_out, b->planes, 1, kp, n128(b->c_out), k3, 0, false, err))) return rc;
  if (b->down &&
      (rc = fold(m, s, &b->c3, p + "downsample.0.weight", p + "downsample.1", b->c_out, b->c_in, 1, kin, n128(b->c_out), k3, kp, true, err)))
    return rc;
  return CLIPEBC_OK;
}

GemmParams conv_params(int fp16, int M, int N, int K, void* out, int ldo, const float* bias, int gh, int gw) {
  GemmParams p = gemm_params_plain(M, N, K);
  p.out = out; p.ldo = ldo; p.bias = bias; p.ab_fp16 = fp16; p.out_fp16 = fp16;
  p.mask_hp = gh + 1; p.mask_wp = gw + 1; p.mask_lead = 0;
  return p;
}
// 3x3 conv, padding 1, over a shared-border grid: 9 K-segments of `c` channels, segment (ky, kx) shifted by (ky-1, kx-1)
void conv3x3_segments(GemmParams* p, int c, int gw) {
  p->n_seg = 9; p->seg_kblocks = c / 64;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      p->seg_row_shift[ky * 3 + kx] = (ky - 1) * (gw + 1) + (kx - 1);
      p->seg_col_start[ky * 3 + kx] = 0;
    }
}

// One Bottleneck on `n` units: X16 [n * (gh+1) * (gw+1), ldx] -> out [n * (gh/s + 1) * (gw/s + 1), n128(c_out)]
int run_block(clipebc_model* m, cudaStream_t s, const RnBlock& b, int n, int gh, int gw, const void* X, int ldx, void* out) {
  ResNetPack& R = *m->resnet;
  const int fp16 = m->cfg.operand_fp16 != 0;
  const int np = n128(b.planes), kp = k64(b.planes), kin = k64(b.c_in), no = n128(b.c_out);
  const int64_t rows_in = static_cast<int64_t>(n) * (gh + 1) * (gw + 1);
  const int oh = gh / b.stride, ow = gw / b.stride;
  const int64_t rows_out = static_cast<int64_t>(n) * (oh + 1) * (ow + 1);
  if (rows_in > 0x7fffffff) return fail(CLIPEBC_EINVAL, "resnet: pass too large");
  const __nv_bfloat16* Xb = static_cast<const __nv_bfloat16*>(X);

  CUDA_TRY(R.t1.reserve(static_cast<size_t>(rows_in) * np * 2));
  K_TRY(gemm_dispatch(s, EPI_BIAS_RELU_MASK_BF16, Xb, rows_in, kin, ldx, b.c1.w.as<__nv_bfloat16>(), b.c1.k,
                      conv_params(fp16, static_cast<int>(rows_in), np, kin, R.t1.p, np, b.c1.b.as<float>(), gh, gw), 0));

  const int kcat = kp + kin;
  // conv2 may write straight into the concatenated operand of the last GEMM when the block does not stride and its padded
  // output fits the first part exactly (np == kp: planes is a multiple of 128)
  const bool direct = b.down && b.stride == 1 && np == kp;
  if (b.down) CUDA_TRY(R.cc.reserve(static_cast<size_t>(rows_out) * kcat * 2));
  if (!direct) CUDA_TRY(R.t2.reserve(static_cast<size_t>(rows_in) * np * 2));
  GemmParams p2 = conv_params(fp16, static_cast<int>(rows_in), np, 9 * kp, direct ? R.cc.p : R.t2.p, direct ? kcat : np,
                              b.c2.b.as<float>(), gh, gw);
  conv3x3_segments(&p2, kp, gw);
  K_TRY(gemm_dispatch(s, EPI_BIAS_RELU_MASK_BF16, R.t1.as<__nv_bfloat16>(), rows_in, np, np, b.c2.w.as<__nv_bfloat16>(), b.c2.k, p2, 0));

  if (b.down) {
    // [conv2 output | block input] on the output grid, average-pooled when the block strides (blocks.py:71,83)
    if (!direct) K_TRY(pool_copy(s, R.t2.p, np, 0, R.cc.p, kcat, 0, kp, n, oh, ow, b.stride, fp16));
    K_TRY(pool_copy(s, X, ldx, 0, R.cc.p, kcat, kp, kin, n, oh, ow, b.stride, fp16));
    K_TRY(gemm_dispatch(s, EPI_BIAS_RELU_MASK_BF16, R.cc.as<__nv_bfloat16>(), rows_out, kcat, kcat, b.c3.w.as<__nv_bfloat16>(), b.c3.k,
                        conv_params(fp16, static_cast<int>(rows_out), no, kcat, out, no, b.c3.b.as<float>(), oh, ow), 0));
  } else {
    GemmParams p3 = conv_params(fp16, static_cast<int>(rows_in), no, kp, out, no, b.c3.b.as<float>(), gh, gw);
    p3.resid16 = X; p3.ldr = ldx;
    K_TRY(gemm_dispatch(s, EPI_BIAS_RESID16_RELU_MASK_BF16, R.t2.as<__nv_bfloat16>(), rows_in, np, np, b.c3.w.as<__nv_bfloat16>(),
                        b.c3.k, p3, 0));
  }
  return CLIPEBC_OK;
}

}  // namespace

int resnet_default_chunk(const clipebc_model* m, int h, int w) {
  if (m->cfg.window_chunk > 0) return m->cfg.window_chunk;
  // ~15 MB of 16-bit activations per 224 x 224 window (the three stem maps at 113 x 113 x 128 dominate): 96 windows per pass
  // are 1.4 GB of workspace and give the GEMMs of layer3 / layer4 (15 x 15 grids) 85 row tiles for the 74 CTA pairs
  const int64_t px = static_cast<int64_t>(h) * w;
  return static_cast<int>(std::max<int64_t>(1, 96 * 224 * 224 / px));
}

int resnet_pack(clipebc_model* m, cudaStream_t s) {
  const clipebc_config& c = m->cfg;
  const int fp16 = c.operand_fp16 != 0;
  std::string err;
  int rc;
  m->resnet.reset(new ResNetPack());
  ResNetPack& R = *m->resnet;
  R.enc_reduction = c.reduction <= 16 ? 16 : 32;  // image_encoder.py:50,68
  const int sw = static_cast<int>(dim(m, "image_encoder.conv3.weight", 0));  // stem width: 64 (RN50 / 101), 80, 96, 128
  if (sw < 16 || sw > 128 || (sw & 15)) return fail(CLIPEBC_ESTATE, "pack: unsupported CLIP-ResNet stem width");
  R.stem_width = sw;
  if (!check_shape(m, "image_encoder.conv1.weight", {sw / 2, 3, 3, 3}, &err)) return fail(CLIPEBC_ESTATE, "pack: " + err);
  // stem: conv1 as [w/2, 27] over im2col rows (k = c * 9 + ky * 3 + kx, the memory order of the weight), K padded to 64; the
  // three stem maps have w/2, w/2, w <= 128 channels in 128-column buffers, read as 64-column K-segments
  if ((rc = fold(m, s, &R.stem1, "image_encoder.conv1.weight", "image_encoder.bn1", sw / 2, 27, 1, 64, 128, 64, 0, false, &err))) return rc;
  if ((rc = fold(m, s, &R.stem2, "image_encoder.conv2.weight", "image_encoder.bn2", sw / 2, sw / 2, 9, 64, 128, 576, 0, false, &err))) return rc;
  if ((rc = fold(m, s, &R.stem3, "image_encoder.conv3.weight", "image_encoder.bn3", sw, sw / 2, 9, 64, 128, 576, 0, false, &err))) return rc;
  int c_prev = sw;
  for (int layer = 1; layer <= 4; ++layer)
    for (int i = 0;; ++i) {
      const std::string p = "image_encoder.layer" + std::to_string(layer) + "." + std::to_string(i) + ".";
      if (!has(m, p + "conv1.weight")) {
        if (i == 0) return fail(CLIPEBC_ESTATE, "pack: missing tensor '" + p + "conv1.weight'");
        break;
      }
      std::unique_ptr<RnBlock> b(new RnBlock());
      b->prefix = p;
      b->planes = static_cast<int>(dim(m, p + "conv1.weight", 0));
      b->c_in = static_cast<int>(dim(m, p + "conv1.weight", 1));
      b->c_out = static_cast<int>(dim(m, p + "conv3.weight", 0));
      b->stride = (i == 0 && (layer == 2 || layer == 3 || (layer == 4 && c.reduction > 16))) ? 2 : 1;
      b->down = has(m, p + "downsample.0.weight");
      if (b->c_in != c_prev || b->planes < 16 || (b->planes & 7) || b->c_out != 4 * b->planes || (b->stride > 1 && !b->down) ||
          (!b->down && b->c_in != b->c_out))
        return fail(CLIPEBC_ESTATE, "pack: unexpected bottleneck shape at '" + p + "'");
      if ((rc = pack_block(m, s, b.get(), &err))) return rc;
      c_prev = b->c_out;
      R.enc.push_back(std::move(b));
    }
  R.c_feat = c_prev;
  for (int j = 0;; ++j) {
    const std::string p = "image_decoder." + std::to_string(j) + ".";
    if (!has(m, p + "conv1.weight")) {
      if (j == 0) return fail(CLIPEBC_ESTATE, "pack: missing tensor '" + p + "conv1.weight'");
      break;
    }
    std::unique_ptr<RnBlock> b(new RnBlock());
    b->prefix = p;
    b->planes = static_cast<int>(dim(m, p + "conv1.weight", 0));   // models/utils.py:348: width = out_channels, expansion 1
    b->c_in = static_cast<int>(dim(m, p + "conv1.weight", 1));
    b->c_out = static_cast<int>(dim(m, p + "conv3.weight", 0));
    b->stride = 1;
    b->down = has(m, p + "downsample.0.weight");
    if (b->c_in != c_prev || (b->planes & 7) || b->planes < 16 || b->c_out != b->planes || (!b->down && b->c_in != b->c_out))
      return fail(CLIPEBC_ESTATE, "pack: unexpected decoder block shape at '" + p + "'");
    if ((rc = pack_block(m, s, b.get(), &err))) return rc;
    c_prev = b->c_out;
    R.dec.push_back(std::move(b));
  }
  R.c_dec = c_prev;
  // projection: embed_dim rows padded to a multiple of 256 (the head epilogue works on 256-wide tiles; resnet50x4 has 640):
  // zero weight rows, zero bias and zero text columns add nothing to ||f||^2 or to the bin dot products
  const int E = c.embed_dim;
  R.e_pad = (E + 255) / 256 * 256;
  const int kd = k64(R.c_dec);
  if (!check_shape(m, "projection.weight", {E, R.c_dec, 1, 1}, &err) || !check_shape(m, "projection.bias", {E}, &err) ||
      !check_shape(m, "text_features", {c.num_bins, E}, &err) || !check_shape(m, "anchor_points", {c.num_bins}, &err))
    return fail(CLIPEBC_ESTATE, "pack: " + err);
  CUDA_TRY(R.w_proj.reserve(static_cast<size_t>(R.e_pad) * kd * 2));
  CUDA_TRY(R.b_proj.reserve(static_cast<size_t>(R.e_pad) * 4));
  CUDA_TRY(cudaMemsetAsync(R.w_proj.p, 0, static_cast<size_t>(R.e_pad) * kd * 2, s));
  CUDA_TRY(cudaMemsetAsync(R.b_proj.p, 0, static_cast<size_t>(R.e_pad) * 4, s));
  K_TRY(fold_conv_bn_general(s, raw_ptr(m, "projection.weight"), nullptr, nullptr, nullptr, nullptr, 0.f, E, R.c_dec, 1, kd,
                             R.w_proj.p, kd, 0, nullptr, 0, fp16));
  CUDA_TRY(cudaMemcpyAsync(R.b_proj.p, raw_ptr(m, "projection.bias"), static_cast<size_t>(E) * 4, cudaMemcpyDeviceToDevice, s));
  return CLIPEBC_OK;
}

int resnet_run_windows(clipebc_model* m, cudaStream_t s, const float* image_dev, int H, int W, const int* origins_yx_dev, int nw,
                       int h, int w, float* exp_out, float* logits_out) {
  const clipebc_config& c = m->cfg;
  ResNetPack& R = *m->resnet;
  const int fp16 = c.operand_fp16 != 0;
  // ---- stem: three 3x3 convs on the (h/2) x (w/2) grid, then avgpool(2) (image_encoder.py:77-82) ----
  const int g1h = h / 2, g1w = w / 2;
  const int64_t rows1 = static_cast<int64_t>(nw) * (g1h + 1) * (g1w + 1);
  if (rows1 > 0x7fffffff) return fail(CLIPEBC_EINVAL, "resnet: pass too large");
  CUDA_TRY(R.col.reserve(static_cast<size_t>(rows1) * 64 * 2));
  CUDA_TRY(R.s1.reserve(static_cast<size_t>(rows1) * 128 * 2));
  CUDA_TRY(R.s2.reserve(static_cast<size_t>(rows1) * 128 * 2));
  K_TRY(stem_im2col(s, image_dev, nw, H, W, origins_yx_dev, h, w, R.col.p, fp16));
  set_launch_tag("rn_stem");
  K_TRY(gemm_dispatch(s, EPI_BIAS_RELU_MASK_BF16, R.col.as<__nv_bfloat16>(), rows1, 64, 64, R.stem1.w.as<__nv_bfloat16>(), 64,
                      conv_params(fp16, static_cast<int>(rows1), 128, 64, R.s1.p, 128, R.stem1.b.as<float>(), g1h, g1w), 0));
  GemmParams ps = conv_params(fp16, static_cast<int>(rows1), 128, 576, R.s2.p, 128, R.stem2.b.as<float>(), g1h, g1w);
  conv3x3_segments(&ps, 64, g1w);
  K_TRY(gemm_dispatch(s, EPI_BIAS_RELU_MASK_BF16, R.s1.as<__nv_bfloat16>(), rows1, 128, 128, R.stem2.w.as<__nv_bfloat16>(), 576, ps, 0));
  ps.out = R.s1.p; ps.bias = R.stem3.b.as<float>();  // conv3 writes over conv1's output (already consumed)
  K_TRY(gemm_dispatch(s, EPI_BIAS_RELU_MASK_BF16, R.s2.as<__nv_bfloat16>(), rows1, 128, 128, R.stem3.w.as<__nv_bfloat16>(), 576, ps, 0));
  set_launch_tag(nullptr);
  int gh = h / 4, gw = w / 4;
  int ld = n128(R.stem_width);  // pitch of the current map (its channels beyond the real ones are zero)
  CUDA_TRY(R.xa.reserve(static_cast<size_t>(nw) * (gh + 1) * (gw + 1) * ld * 2));
  K_TRY(pool_copy(s, R.s1.p, 128, 0, R.xa.p, ld, 0, ld, nw, gh, gw, 2, fp16));

  // ---- layers 1-4 ----
  DevBuf* cur = &R.xa;
  DevBuf* nxt = &R.xb;
  int rc;
  set_launch_tag("rn_encoder");
  for (const auto& b : R.enc) {
    const int oh = gh / b->stride, ow = gw / b->stride;
    CUDA_TRY(nxt->reserve(static_cast<size_t>(nw) * (oh + 1) * (ow + 1) * n128(b->c_out) * 2));
    if ((rc = run_block(m, s, *b, nw, gh, gw, cur->p, ld, nxt->p))) { set_launch_tag(nullptr); return rc; }
    std::swap(cur, nxt);
    gh = oh; gw = ow; ld = n128(b->c_out);
  }
  set_launch_tag(nullptr);
  // ---- F.interpolate to the reduction grid (model.py:195-196) ----
  const int dh = h / c.reduction, dw = w / c.reduction;
  const void* X = cur->p;
  if (dh != gh || dw != gw) {
    CUDA_TRY(R.up.reserve(static_cast<size_t>(nw) * (dh + 1) * (dw + 1) * ld * 2));
    K_TRY(resample16(s, cur->p, R.up.p, ld, nw, gh, gw, dh, dw, fp16));
    gh = dh; gw = dw;
    X = R.up.p;  // xa / xb stay free for the decoder's ping-pong
  }
  // ---- decoder Bottleneck(s) (models/utils.py:334-390) ----
  set_launch_tag("rn_decoder");
  DevBuf* out = nxt;
  DevBuf* other = cur;
  for (const auto& b : R.dec) {
    CUDA_TRY(out->reserve(static_cast<size_t>(nw) * (gh + 1) * (gw + 1) * n128(b->c_out) * 2));
    if ((rc = run_block(m, s, *b, nw, gh, gw, X, ld, out->p))) { set_launch_tag(nullptr); return rc; }
    X = out->p;
    ld = n128(b->c_out);
    std::swap(out, other);
  }
  set_launch_tag(nullptr);
  // ---- 1x1 projection fused with the EBC head (model.py:198-212) ----
  const int E = R.e_pad, kd = k64(R.c_dec);
  const int64_t Mp = static_cast<int64_t>(nw) * (gh + 1) * (gw + 1);
  const int kParts = 2 * (E / 256);
  CUDA_TRY(m->ws_F.reserve(static_cast<size_t>(Mp) * kParts * (1 + c.num_bins) * 4));
  GemmParams pp = gemm_params_plain(static_cast<int>(Mp), E, kd);
  pp.ab_fp16 = fp16; pp.out_fp16 = fp16;
  pp.bias = R.b_proj.as<float>();
  pp.out = m->ws_F.p; pp.ldo = kParts * (1 + c.num_bins);
  pp.head_tmat = m->tmat.as<float>(); pp.head_bins = c.num_bins;
  set_launch_tag("projection+head");
  K_TRY(gemm_dispatch(s, EPI_BIAS_HEAD_PARTIAL, static_cast<const __nv_bfloat16*>(X), Mp, kd, ld, R.w_proj.as<__nv_bfloat16>(), kd, pp, 256));
  set_launch_tag(nullptr);
  K_TRY(ebc_head_finish(s, m->ws_F.as<float>(), kParts, raw_ptr(m, "anchor_points"), c.num_bins, nw, gh, gw, exp_out, logits_out));
  return CLIPEBC_OK;
}

}  // namespace cebc
