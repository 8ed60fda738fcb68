// Internal (C++) launcher contracts for the sm_100a kernels of the CLIP-EBC hot path.
// The public C-ABI lives in include/clipebc_b200.h; api.cu maps one onto the other.
// Launchers return nullptr on success or a static error string (no exceptions, no torch types).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace cebc {

// ------------------------------------------------------------------ GEMM ---------------------------------------
enum GemmEpilogue : int {
  EPI_F32 = 0,                  // out f32 = acc (+ bias if given)
  EPI_BIAS_F32 = 1,             // out f32 = acc + bias
  EPI_BIAS_BF16 = 2,            // out bf16 = acc + bias
  EPI_BIAS_GELU_BF16 = 3,       // out bf16 = quickgelu(acc + bias)
  EPI_BIAS_RESID_F32 = 4,       // out f32 = acc + bias + resid   (out may alias resid)
  EPI_BIAS_RELU_MASK_BF16 = 5,  // out bf16 = border ? 0 : relu(acc + bias)          (decoder conv1 on the padded grid)
  EPI_BIAS_RESID_RELU_SPLIT = 6,// t = relu(acc + bias + resid); out[:, n] = hi(t) and, when split_lo, out[:, N + n] = lo(t)
  EPI_BIAS_HEAD_PARTIAL = 7,    // f = acc + bias is never written: per row and per half tile (CTA-pair kernel, 256-wide
                                // tiles) out[row][2 * n_blk + half][0] = sum f^2, [1 + b] = sum f * head_tmat[b][n]
                                // -- the projection fused with the first half of the EBC head (ebc_head_finish)
  EPI_BIAS_RESID16_RELU_MASK_BF16 = 9,  // out 16-bit = border ? 0 : relu(acc + bias + resid16[row, n]); resid16 is a 16-bit
                                // [M, ldr] tensor in the output format (the identity branch of a ResNet bottleneck)
  EPI_BIAS_UPSKIP_RELU_SPLIT = 8  // EPI_BIAS_RESID_RELU_SPLIT whose residual is the bilinear upsample of the coarse map,
                                // evaluated on the fly: resid = Y f32 [n_win * up_hp * up_wp, ldr] (ln_post rows), the row's
                                // cell comes from the shared-border grid (mask_hp x mask_wp rows per window); border rows add 0.
                                // The BasicBlock skip of the decoder without materialising the fine-grid map (CTA-pair kernel)
};

constexpr int kMaxGemmSegs = 9;

struct GemmParams {
  int M, N, K;                        // K = n_seg * seg_kblocks * 64
  int n_seg, seg_kblocks;             // A-operand K-segments
  int seg_row_shift[kMaxGemmSegs];    // A row offset of each segment (may be negative: TMA zero-fills OOB rows)
  int seg_col_start[kMaxGemmSegs];    // A column start of each segment
  void* out;                          // f32 or bf16, see epilogue
  int ldo;                            // elements
  const float* bias;                  // [N]
  const float* resid;                 // f32 [M, ldr]
  const void* resid16;                // EPI_BIAS_RESID16_RELU_MASK_BF16: 16-bit [M, ldr]
  int ldr;
  int mask_hp, mask_wp;               // padded grid (rows per image = mask_hp * mask_wp), EPI_BIAS_RELU_MASK_BF16
  int up_hp, up_wp;                   // EPI_BIAS_UPSKIP_RELU_SPLIT: patch grid of the coarse map
  int mask_lead;                      // 1: first AND last row / column of the grid are border; 0: only the last ones
                                      // (shared-border grid: the trailing zero column / row of one line / image is the
                                      // leading border of the next)
  int ab_fp16;                        // 16-bit format of A and W: 0 = bf16, 1 = fp16
  int out_fp16;                       // 16-bit format written by the *_BF16 / SPLIT epilogues: 0 = bf16, 1 = fp16
  int split_lo;                       // *_SPLIT epilogues: 1 = also write lo(t) = t - hi(t) at column N + n (split precision for
                                      // the GEMM behind it), 0 = hi only
  const float* head_tmat;             // EPI_BIAS_HEAD_PARTIAL: f32 [head_bins, N] (logit_scale * normalised text features)
  int head_bins;                      // 1..32
  int w_prefetch;                     // 1 (gemm_params_plain default): W does not depend on the previous kernel of the
                                      // stream, so its first tiles are requested BEFORE the programmatic-dependency wait.
                                      // Must be 0 when the launch just before this one wrote W (weight packing).
  int tma_out;                        // set by the launcher: outputs leave through bulk tensor stores / reductions
};

inline GemmParams gemm_params_plain(int M, int N, int K) {
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.n_seg = 1; p.seg_kblocks = K / 64;
  p.w_prefetch = 1;
  return p;
}

// D[M,N] = A[M,K] * W[N,K]^T. A: 16-bit [a_rows, a_cols] pitch lda; W: 16-bit [N, K] pitch ldw. block_n: 0 = auto, else
// 128 / 192 / 256. CTA-pair (cta_group::2) tcgen05 kernel, gemm2_tcgen05.cu.
const char* gemm2_bf16_tn(cudaStream_t stream, int epi, const __nv_bfloat16* A, int64_t a_rows, int64_t a_cols,
                          int64_t lda, const __nv_bfloat16* W, int64_t ldw, GemmParams p, int block_n);
int gemm2_pick_block_n(int M, int N);  // the tile width gemm2_bf16_tn chooses for block_n = 0
int device_num_sms();  // of the current device (cached per device)
// Per-device, once: cudaFuncSetAttribute(MaxDynamicSharedMemorySize). The attribute is per device, so a process that
// uses several GPUs must set it on each of them; `done_mask` is the caller's static bit mask of devices already served.
template <class Kern>
inline cudaError_t ensure_dyn_smem(Kern kern, int bytes, unsigned long long* done_mask) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (*done_mask & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) *done_mask |= bit;
  return e;
}
// Every kernel launch of this library goes through a LaunchScope: it counts the launch (clipebc_launch_count) and, when
// profiling is enabled (clipebc_profile_enable), brackets it with CUDA events on the launching stream and books the
// duration, algorithmic FLOPs and bytes under `tag` (the current tag set by api.cu, else `kind`).
struct LaunchScope {
  LaunchScope(cudaStream_t stream, const char* kind, double flops = 0.0, double bytes = 0.0);
  ~LaunchScope();
  LaunchScope(const LaunchScope&) = delete;
  LaunchScope& operator=(const LaunchScope&) = delete;
  cudaStream_t stream_;
  int slot_;
};
void set_launch_tag(const char* tag);  // nullptr clears
bool profiling_on();

// Launch with programmatic stream serialization (see common.cuh: pdl_wait) and an optional cluster width.
bool pdl_enabled();  // false when CLIPEBC_NO_PDL is set (A/B experiments)
// Persisting-L2 access-policy window attached to every launch of the calling thread while set (api.cu: L2Window), or nullptr
const cudaAccessPolicyWindow* current_l2_window();
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                       Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[3];
  unsigned n = 0;
  if (const cudaAccessPolicyWindow* w = current_l2_window()) {
    at[n].id = cudaLaunchAttributeAccessPolicyWindow;
    at[n].val.accessPolicyWindow = *w;
    ++n;
  }
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = static_cast<unsigned>(cluster_x); at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------ LayerNorm ----------------------------------
// out[r] = LN(in[map(r)]) * gamma + beta over D = `width` channels (768: ViT-B, 1024: ViT-L/14), eps 1e-5, fp32 statistics
// (two-pass, in registers). out_kind: 0 = f32, 1 = bf16, 2 = fp16.
// Row map: in_row = (r / rows_out_per_group) * rows_in_per_group + in_row_offset + r % rows_out_per_group.
// out16_extra (f32 output only, nullable): also the 16-bit rounding of the rows, in the format fp16_extra selects.
const char* layernorm_rows(cudaStream_t stream, int width, const float* in, const float* gamma, const float* beta, void* out,
                           int out_kind, int64_t n_rows_out, int rows_out_per_group, int rows_in_per_group,
                           int in_row_offset, void* out16_extra = nullptr, int fp16_extra = 0);

// ------------------------------------------------------------------ stem ---------------------------------------
// image f32 [n_img, 3, H, W] -> patch rows (16-bit, fp16 flag): split = 1: [n_img * gh * gw, 2 * KP] = [hi | lo] split
// of the pixels (hi = round16(x), lo = round16(x - hi)); split = 0: [n_img * gh * gw, KP] = hi only. KP = kp_pad >= 3 * patch^2 (columns beyond 3 * patch^2 are zero: the GEMM needs K % 64 == 0, ViT-L/14 has
// 3 * 14^2 = 588 -> 640), k = c * patch^2 + py * patch + px (= conv1.weight.view(width, 3 * patch^2)), on the grid whose
// (0,0) patch starts at pixel (y0, x0) of each image (gh, gw patches). patch = 16 / 32 (ViT-B) or 14 (ViT-L/14).
const char* patchify(cudaStream_t stream, const float* image, int n_img, int H, int W, int y0, int x0, int gh, int gw,
                     int patch, int kp_pad, int split, void* out, int fp16);
// per-window patchify when window origins are not on the patch grid: out rows [n_win * hp * wp, (1 + split) * KP]
const char* patchify_windows(cudaStream_t stream, const float* image, int H, int W, const int* origins_yx_dev,
                             int n_win, int hp, int wp, int patch, int kp_pad, int split, void* out, int fp16);

// Assemble the residual stream X f32 [n_win * t_live, width]:
//   row 0            : LN_pre(class_emb + pos[0])
//   rows 1..n_prompt : vpt0 rows (shallow VPT only; n_prompt = 0 for deep)            (model.py:161-168)
//   remaining rows   : LN_pre(patch_embed[src_row(win, p)] + pos[1 + p])              (model.py:147-157)
// src_row = win_base[win] + (p / wp) * pitch + p % wp  (gather from a shared per-image patch grid or per-window rows);
// pitch = win_pitch_dev[win] when given (windows of several images in one pass), else src_pitch
const char* assemble_tokens(cudaStream_t stream, int width, const float* patch_embed, const int* win_base_dev, int src_pitch,
                            const int* win_pitch_dev, const float* class_emb, const float* pos, const float* ln_g, const float* ln_b,
                            const float* vpt0, int n_prompt, int n_win, int hp, int wp, float* X);

// ------------------------------------------------------------------ attention ----------------------------------
// qkv bf16 [n_win * t_live, 3 * width] (q | k | v, head h at columns 64h..64h+63 of each third, width = 64 * heads);
// const_kv bf16 [n_const, 3 * width] rows appended as extra keys/values for every window (deep-VPT prompt tokens); out
// [n_win * t_live, width] (bf16, or fp16 when out_fp16). softmax(q k^T / 8) v per (window, head), no mask.
//
// attention_h64_pp (attention_pp.cu): tcgen05 / TMEM, persistent, two independent chains per CTA, one per TMEM buffer: a
// thread owns a whole query row, the softmax of one chain overlaps the tensor work of the other. Needs
// t_live + n_const <= 256 and n_const % 8 == 0.
const char* attention_h64_pp(cudaStream_t stream, const __nv_bfloat16* qkv, const __nv_bfloat16* const_kv, int n_const,
                             int n_win, int t_live, int heads, void* out, int out_fp16);
// attention_h64_ppl (attention_ppl.cu): the two-chain tcgen05 kernel with the keys of a tile taken in two blocks (256 + up to
// 64) inside the chain's TMEM buffer: 257..320 keys (ViT-L/14 at 224 x 224: 257 + 32), n_const % 16 == 0, t_live <= 384.
bool attention_h64_ppl_takes(int n_const, int t_live);
const char* attention_h64_ppl(cudaStream_t stream, const __nv_bfloat16* qkv, const __nv_bfloat16* const_kv, int n_const,
                              int n_win, int t_live, int heads, void* out, int out_fp16);
// attention_h64_long (attention.cu): any sequence length and any n_const -- 64-query chunks, K / V streamed in 64-key
// blocks, online softmax, mma.sync. Windows with more than 256 tokens (ViT-L/14 224-windows: 289; 448 x 448 windows).
const char* attention_h64_long(cudaStream_t stream, const __nv_bfloat16* qkv, const __nv_bfloat16* const_kv, int n_const,
                               int n_win, int t_live, int heads, void* out, int out_fp16);

// ------------------------------------------------------------------ decoder / head -----------------------------
// Y f32 [n_win * hp * wp, width] (ln_post rows) -> bilinear resample (align_corners = False, scale = gh/hp) into the
// shared-border NHWC grids U_16 / U_f32 [n_win, gh + 1, gw + 1, width]: cell (y, x) of a window at row y * (gw + 1) + x,
// column gw and row gh are zero. The zero column that ends one line is the left border of the next line, the zero row
// that ends one window is the top border of the next (rows before the buffer are zero-filled by TMA): every 3x3 tap
// r + dy * (gw + 1) + dx of an interior cell lands on the right neighbour or on a zero  (model.py:195-196).
const char* resample_to_padded(cudaStream_t stream, int width, const float* Y, int n_win, int hp, int wp, int gh, int gw,
                               void* U_16, float* U_f32, int fp16);

// conv1 of the decoder from the coarse patch grid (elementwise.cu: conv1_from_coarse_kernel): Z 16-bit
// [n_win * hp * wp, 9 * width] = Y_16 x Wz^T (Wz from fold_conv3x3_bn_tapout, column = tap * width + o) ->
// D1 16-bit [n_win * (gh+1) * (gw+1), width] = relu(conv3x3(bilinear_up(Y)) + bias) on the shared-border grid (border rows 0)
const char* conv1_from_coarse(cudaStream_t stream, int width, const void* Z, const float* bias, int n_win, int hp, int wp,
                              int gh, int gw, void* D1, int fp16);
const char* fold_conv3x3_bn_tapout(cudaStream_t stream, const float* W, const float* gamma, const float* var, float eps, int O,
                                   int I, void* Wz, int fp16);

// Second half of the fused head: partial f32 [n_win * (gh+1) * (gw+1), n_part, 1 + n_bins] written by the projection GEMM
// with EPI_BIAS_HEAD_PARTIAL -> sums over the n_part partials in a fixed order, 1 / max(||f||, 1e-12), softmax over the
// bins, anchor expectation on the interior cells: exp_out f32 [n_win, 1, gh, gw]; logits_out (nullable) f32
// [n_win, n_bins, gh, gw]  (model.py:200-212).
const char* ebc_head_finish(cudaStream_t stream, const float* partial, int n_part, const float* anchors, int n_bins, int n_win,
                            int gh, int gw, float* exp_out, float* logits_out);

// ------------------------------------------------------------------ CLIP-ResNet encoder path (resnet.cu) ---------
// Stem conv1 (3 -> 32, 3x3, stride 2, pad 1) as im2col rows for the GEMM: 16-bit [n_units * (h/2+1) * (w/2+1), 64] on the
// shared-border grid, k = c * 9 + ky * 3 + kx < 27, rest zero. Units: images of a batch [n, 3, H, W] (origins == nullptr,
// h == H, w == W) or windows of one image [3, H, W] at origins_yx_dev (device, y then x per unit).
const char* stem_im2col(cudaStream_t stream, const float* image, int n_units, int H, int W, const int* origins_yx_dev, int h,
                        int w, void* out, int fp16);
// 16-bit NHWC maps on shared-border grids: dst[.., dst_col + c] = S x S mean of src[.., src_col + c] (S = 1 copy, S = 2
// nn.AvgPool2d(2)); dst grid (go_h + 1) x (go_w + 1) per unit, src grid (S go_h + 1) x (S go_w + 1); dst border = 0.
const char* pool_copy(cudaStream_t stream, const void* src, int ld_src, int src_col, void* dst, int ld_dst, int dst_col, int C,
                      int n_units, int go_h, int go_w, int S, int fp16);
// bilinear resample (align_corners = False) of a 16-bit NHWC map [.., C] between shared-border grids gi -> go
const char* resample16(cudaStream_t stream, const void* src, void* dst, int C, int n_units, int gi_h, int gi_w, int go_h, int go_w,
                       int fp16);
// BatchNorm (eval) folded into the conv before it, written as a K-major GEMM operand with padding / concatenation:
// Wp[o * ldw + col_off + tap * i_pad + i] = W[o, i, tap] * s(o); bias[o] (+)= beta - mean * s(o), s = gamma / sqrt(var + eps)
const char* fold_conv_bn_general(cudaStream_t stream, const float* W, const float* gamma, const float* beta, const float* mean,
                                 const float* var, float eps, int O, int I, int taps, int i_pad, void* Wp, int ldw, int col_off,
                                 float* bias, int accumulate_bias, int fp16);

// ------------------------------------------------------------------ fold ---------------------------------------
// preds f32 [n_rows * n_cols, 1, gh, gw] -> density f32 [Ho, Wo]: average of overlapping windows, summed in ascending
// window order (bit-exact vs the numpy loop of eval_utils.py:79-95); optional per-image sum (count).
const char* fold_average(cudaStream_t stream, const float* preds, const int* row_cells_dev, const int* col_cells_dev,
                         int n_rows, int n_cols, int gh, int gw, int Ho, int Wo, float* density, float* count_out);

// ------------------------------------------------------------------ before / after the hot path ----------------
// in [C, h, w] (uint8 when in_is_u8, else f32 in [0,1]) -> out f32 [C, H, W]: bicubic antialiased resize as
// TF.resize(BICUBIC, antialias=True) (datasets/transforms.py:27-35); tmp f32 [C, h, W]; mean/std (host, nullable
// together) apply torchvision Normalize to the result (datasets/crowd.py:64).
const char* resize_bicubic_aa(cudaStream_t stream, const void* in, int in_is_u8, int C, int h, int w, float* tmp, float* out,
                              int H, int W, const float* mean_host, const float* std_host);
// right / bottom zero padding of the [0,1] image to [C, H, W] (datasets/transforms.py:138-140), then Normalize
const char* pad_normalize(cudaStream_t stream, const void* in, int in_is_u8, int C, int h, int w, float* out, int H, int W,
                          const float* mean_host, const float* std_host);
// x f32 [h, w] -> out f32 [H, W] = bilinear(x) * nan_to_num(sum(bilinear(x)) / sum(x))   (utils/eval_utils.py:19-23);
// workspace: resize_density_workspace_floats() floats; sums_out (nullable) <- [sum(x), sum(bilinear(x))]
int resize_density_workspace_floats();
const char* resize_density_map(cudaStream_t stream, const float* x, int h, int w, int H, int W, float* out, float* workspace,
                               float* sums_out);

// ------------------------------------------------------------------ pack-time helpers --------------------------
const char* f32_to_16(cudaStream_t stream, const float* in, void* out, int64_t n, int fp16);
// W f32 [O, I, 3, 3] + BN(gamma, beta, mean, var, eps) -> Wp bf16 [O, 9*I] (tap-major K: k = (ky*3+kx)*I + i), bias f32 [O]
const char* fold_conv3x3_bn(cudaStream_t stream, const float* W, const float* gamma, const float* beta, const float* mean,
                            const float* var, float eps, int O, int I, void* Wp, float* bias, int fp16);
// W f32 [O, I] -> 16-bit [O, 3 * Ip] = [hi | hi | lo], every third zero-padded from I to Ip >= I columns
const char* split_weight_hi_hi_lo(cudaStream_t stream, const float* W, int O, int I, int Ip, void* out, int fp16);
// text f32 [n, d] -> tmat [n, ld_out] (ld_out = 0: d; columns beyond d are left untouched) = exp(logit_scale) * text / max(||text||, 1e-12)
const char* pack_text(cudaStream_t stream, const float* text, const float* logit_scale, int n, int d, float* tmat, int ld_out = 0);

}  // namespace cebc
