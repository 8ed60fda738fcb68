/*
 * clipebc_b200 -- C-ABI of the B200-native (sm_100a) CLIP-EBC inference hot path.
 *
 * Drop-in boundary for the two reference entry points (the reference has no FFI layer of its own; both are Python
 * callables, see INTEGRATION.md for the ctypes binding a maintainer would add):
 *
 *   get_model(...)/model(x)            /root/reference/models/__init__.py:10-29, models/clip/model.py:191-217
 *   sliding_window_predict(...)        /root/reference/utils/eval_utils.py:26-96
 *
 * Conventions
 *   - every function returns 0 on success, a CLIPEBC_E* code otherwise; clipebc_last_error() gives the message of the
 *     last failure on the calling thread (reference convention: Python AssertionError / RuntimeError; the Python host
 *     side in clip_ebc_b200/ keeps the reference's asserts and maps non-zero codes to RuntimeError).
 *   - plain pointers and sizes only. "dev" pointers are CUDA device pointers on the current device, "host" pointers
 *     are ordinary host memory; `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - all work is enqueued on `stream`; inputs are borrowed and never written; outputs are caller-allocated.
 *   - there is NO CPU fallback: without an sm_100 device every compute entry point fails with CLIPEBC_ECUDA.
 */
#ifndef CLIPEBC_B200_H_
#define CLIPEBC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLIPEBC_OK 0
#define CLIPEBC_EINVAL 1   /* bad argument / shape / unsupported configuration */
#define CLIPEBC_ECUDA 2    /* CUDA runtime or driver error (message carries cudaGetErrorString) */
#define CLIPEBC_ESTATE 3   /* call order: tensors missing, model not packed, ... */

#define CLIPEBC_ABI_VERSION 9

typedef struct clipebc_model clipebc_model;

/* Hyper-parameters of CLIP_EBC(backbone="vit_b_16" | "vit_b_32" | "vit_l_14" | "resnet50" | "resnet101") -- models/clip/model.py:16-24,31-45,
 * _clip_ebc :220-270. The backbone is given by its dimensions:
 *     vit_b_16: patch 16, width 768,  layers 12, embed_dim 512      vit_b_32: patch 32, width 768, layers 12, embed_dim 512
 *     vit_l_14: patch 14, width 1024, layers 24, embed_dim 768      (heads = width / 64, hidden = 4 * width)
 * `struct_size` MUST be set to sizeof(clipebc_config) by the caller: clipebc_model_create rejects any other value with
 * CLIPEBC_EINVAL, so a binding written against an older (shorter) struct fails loudly instead of being read past its end. */
typedef struct clipebc_config {
  uint32_t struct_size; /* sizeof(clipebc_config) as the caller sees it                                              */
  int input_size;   /* side of the square the positional embedding was built for (224)                              */
  int reduction;    /* 8, 16 or 32 (model.reduction; encoder reduction is the patch size)                            */
  int num_vpt;      /* visual prompt tokens per layer (32)                                                          */
  int deep_vpt;     /* 1: per-layer prompts vpt_0..vpt_{layers-1}, 0: shallow (vpt_0 only, propagated)                */
  int num_bins;     /* N = len(bins) = len(anchor_points), 1..32                                                    */
  int window_chunk; /* windows per internal pass (0 = library default)                                               */
  int operand_fp16; /* 16-bit tensor-core operand format: 1 = fp16 (11-bit mantissa; CLIP's released weights are fp16,
                       all operands are range-bounded and saturated), 0 = bf16. Accumulation, residual stream,
                       LayerNorm statistics, softmax and the head are fp32 either way. With bf16 the patch-embedding and
                       projection GEMMs run in split precision (hi + lo operands, 3 K-segments), with fp16 in one.   */
  int patch;        /* ViT patch size = encoder reduction: 16 (also when 0), 32 or 14
                       -- _clip/image_encoder.py:141, models/clip/model.py:78                                        */
  int width;        /* transformer width = decoder channels: 768 (also when 0) or 1024                               */
  int layers;       /* transformer blocks: 12 (also when 0) or 24                                                    */
  int embed_dim;    /* CLIP embedding = projection outputs = text feature length: 512 (also when 0) or 768             */
  int decoder_conv1_fine; /* 0 (default): conv1 of the decoder is computed from the coarse patch grid whenever the decoder
                       grid is at least twice as fine (conv3x3(bilinear_up(Y)) = 9 per-tap channel contractions on the patch
                       grid + a bilinear gather; DESIGN.md section 2 rewrite 8); 1: always the implicit GEMM on the fine
                       grid. Same result to a few 16-bit roundings; per model, for A/B measurements.                  */
  int encoder;      /* 0 (default): CLIP VisionTransformer + VPT, described by patch / width / layers above. 1: CLIP
                       ModifiedResNet of width 64 (resnet50, resnet101 -- _clip/image_encoder.py:10-115) with the Bottleneck
                       decoder of models/clip/model.py:228-239: its depth, channel counts and decoder blocks are read off the
                       tensors loaded with clipebc_model_set_tensor (image_encoder.{conv1..3,bn1..3,layer{1..4}.{i}.*},
                       image_decoder.{j}.*, projection.*); input_size / num_vpt / deep_vpt / patch / width / layers are
                       ignored, embed_dim is the CLIP embedding (1024 / 512), windows must be multiples of 32 pixels.  */
} clipebc_config;

const char* clipebc_last_error(void);
int clipebc_abi_version(void);
/* Number of kernels this library has launched so far in this process (bench.py's gpu_launches). */
int64_t clipebc_launch_count(void);

/* Optional per-launch profiling: when enabled every kernel launch is bracketed by CUDA events on its stream.
 * clipebc_profile_dump synchronises the device and writes a JSON object {"<kernel>[:<use>]": {"ms", "launches",
 * "flops", "bytes"}} (algorithmic work of the launches, summed since enable) into buf. Enabling clears the records. */
int clipebc_profile_enable(int on);
int clipebc_profile_dump(char* buf, int cap);
int clipebc_profile_enabled(void);
/* For host layers that replay captured CUDA graphs of this library's launches (clip_ebc_b200/model.py): the epoch
 * changes whenever a device buffer of the library is (re)allocated or released (a captured graph holds the workspace
 * addresses used at capture time), and a replay reports the launches it contains so that clipebc_launch_count stays the
 * number of kernels actually run. */
int64_t clipebc_config_epoch(void);
void clipebc_note_replayed_launches(int64_t n);

/* ---- model lifetime: mirrors get_model() + load_state_dict() + .eval() ---------------------------------------- */
/* A handle belongs to the CUDA device that is current when it is created: its weights and workspaces live there, and every
 * later call on it must be made with that device current (CLIPEBC_ESTATE otherwise). One handle per device. */
int clipebc_model_create(const clipebc_config* cfg, clipebc_model** out);
void clipebc_model_destroy(clipebc_model* m);
/* Upload one fp32 tensor under its reference state_dict key (models/clip/model.py state_dict):
 *   vpt_{l} (only when num_vpt > 0), logit_scale, image_encoder.{class_embedding,positional_embedding,conv1.weight,ln_pre.*,ln_post.*,
 *   transformer.resblocks.{l}.{attn.in_proj_weight,attn.in_proj_bias,attn.out_proj.*,ln_1.*,ln_2.*,mlp.c_fc.*,
 *   mlp.c_proj.*}}, image_decoder.0.{conv1.weight,bn1.*,conv2.weight,bn2.*}, projection.{weight,bias}
 * plus the two plain attributes of the reference module: text_features [N, embed_dim] and anchor_points [N].
 * `data` may be a host or a device pointer (copied, caller keeps ownership). Invalidates a previous pack. */
int clipebc_model_set_tensor(clipebc_model* m, const char* name, const float* data, const int64_t* shape, int ndim);
/* Build the device-resident packed form: bf16 GEMM layouts, BatchNorm folded into the decoder convs, hi/lo split of
 * the projection, exp(logit_scale) * normalised text matrix, constant prompt K/V rows (deep VPT). */
int clipebc_model_pack(clipebc_model* m, void* stream);

/* ---- the hot path --------------------------------------------------------------------------------------------- */
/* model(x) in eval mode: x_dev f32 [B,3,h,w] -> exp_out_dev f32 [B,1,h/r,w/r]; logits_out_dev (nullable) f32
 * [B,N,h/r,w/r] is the train-mode first output (models/clip/model.py:214-217). */
int clipebc_forward_windows(clipebc_model* m, const float* x_dev, int B, int h, int w, float* exp_out_dev,
                            float* logits_out_dev, void* stream);
/* sliding_window_predict(model, image[1,3,H,W], (wh,ww), (sh,sw)) -> density_out_dev f32 [H/r, W/r] (the [1,1,.,.]
 * tensor of the reference, on the device) and, if count_out_dev != NULL, its sum (eval.py:35, test_nwpu.py:100). */
int clipebc_sliding_window_predict(clipebc_model* m, const float* image_dev, int H, int W, int wh, int ww, int sh,
                                   int sw, float* density_out_dev, float* count_out_dev, void* stream);

/* The same for a batch of images of arbitrary (different) sizes: the windows of all images share the passes of the
 * ViT / decoder / head, so small images (a handful of windows each) still fill the GPU; each image is folded on its
 * own. This is the cross-image window batching of the evaluation loop (eval.py:25-35 processes one image per
 * iteration because sizes vary). images_dev / heights / widths / density_out_dev: HOST arrays of n_images entries
 * (device pointers / sizes); counts_out_dev: device array of n_images floats or NULL. Results are identical to
 * n_images separate clipebc_sliding_window_predict calls. */
int clipebc_sliding_window_predict_batch(clipebc_model* m, int n_images, const float* const* images_dev,
                                         const int* heights, const int* widths, int window_h, int window_w,
                                         int stride_h, int stride_w, float* const* density_out_dev,
                                         float* counts_out_dev, void* stream);

/* ---- integer part of sliding_window_predict, host only (utils/eval_utils.py:54-66) ----------------------------- */
/* Writes n_rows/n_cols and, when the arrays are non-NULL (capacity >= n_rows / n_cols), the clamped window origins. */
int clipebc_window_origins(int H, int W, int wh, int ww, int sh, int sw, int* n_rows, int* n_cols, int* row_origins,
                           int* col_origins);

/* ---- single kernels (unit/parity tests and profiling; all pointers are device pointers) ------------------------ */
/* "16" = a 16-bit floating format selected by an fp16 flag: 0 = bf16, 1 = fp16 (saturating). */
int clipebc_f32_to_16(const float* in_dev, void* out_16_dev, int64_t n, int fp16, void* stream);
/* D[M,N] = A[M,K] W[N,K]^T on the CTA-pair tcgen05 kernel with a fused epilogue.
 * epi: 0 f32, 1 bias f32, 2 bias ->16, 3 bias+quickgelu ->16, 4 bias+resid f32, 5 bias+relu+border-mask ->16,
 *      6 bias+resid+relu hi/lo split ->16, 9 bias + 16-bit resid (resid_dev points to a 16-bit [M, ldr] tensor) + relu +
 *      border-mask ->16. ab_fp16: format of A and W; out_fp16: format of a 16-bit output.
 *      n_seg K-segments of A with per-segment row shift / column start (implicit-GEMM 3x3 taps, split precision).
 *      mask_hp x mask_wp: rows per image of the zero-bordered grid of epi 5; mask_lead 1: first and last row/column are
 *      border, 0: only the last ones (shared-border grid, see clipebc_resample_to_padded). block_n: 0 (auto), 128, 192, 256.
 *      See clip_ebc_b200/csrc/kernels.h for the contract. */
int clipebc_gemm_bf16(int epi, const void* A_16_dev, int64_t a_rows, int64_t a_cols, int64_t lda,
                      const void* W_16_dev, int64_t ldw, int M, int N, int K, int n_seg, const int* seg_row_shift,
                      const int* seg_col_start, void* out_dev, int ldo, const float* bias_dev, const float* resid_dev,
                      int ldr, int mask_hp, int mask_wp, int mask_lead, int block_n, int ab_fp16, int out_fp16,
                      void* stream);
/* nn.LayerNorm(width, eps 1e-5) over rows of `width` = 768 or 1024 channels; out_kind: 0 = f32, 1 = bf16, 2 = fp16.
 * Row map: in_row = (r / rows_out_per_group) * rows_in_per_group + in_row_offset + r % rows_out_per_group. */
int clipebc_layernorm(const float* in_dev, const float* gamma_dev, const float* beta_dev, int width, void* out_dev,
                      int out_kind, int64_t n_rows_out, int rows_out_per_group, int rows_in_per_group,
                      int in_row_offset, void* stream);
/* softmax(q k^T / 8) v per (window, head): qkv bf16 [n_win * t_live, 3 * 64 * heads], const_kv bf16 [n_const, 3 * 64 * heads]
 * extra keys / values of every window (deep-VPT prompts), out 16-bit [n_win * t_live, 64 * heads]. tcgen05 kernels when
 * t_live + n_const <= 256 and n_const % 8 == 0, or 257..320 keys with n_const % 16 == 0 (ViT-L/14); streamed-K/V kernel
 * otherwise. */
int clipebc_attention(const void* qkv_bf16_dev, const void* const_kv_bf16_dev, int n_const, int n_win, int t_live,
                      int heads, void* out_16_dev, int out_fp16, void* stream);
/* out: split = 1: [n_img*gh*gw, 2*kp_pad] = [hi | lo] split of the pixels in the 16-bit format (what the path uses with bf16
 * operands); split = 0: [n_img*gh*gw, kp_pad] = hi only (fp16 operands). kp_pad >= 3 * patch^2 (a multiple of 8; pad columns
 * are not written), patch 14, 16 or 32 */
int clipebc_patchify(const float* image_dev, int n_img, int H, int W, int y0, int x0, int gh, int gw, int patch,
                     int kp_pad, int split, void* out_16_dev, int fp16, void* stream);
/* Shared-border grid [n_win, gh+1, gw+1, width]: cell (y, x) at row y*(gw+1)+x, column gw and row gh are zero; the zero
 * column ending a line is the left border of the next line, the zero row ending a window the top border of the next.
 * U_16_dev may be NULL. */
int clipebc_resample_to_padded(const float* Y_dev, int n_win, int hp, int wp, int gh, int gw, int width, void* U_16_dev,
                               float* U_f32_dev, int fp16, void* stream);
/* preds_dev f32 [n_rows*n_cols, gh, gw]; row_cells/col_cells: HOST arrays of window origins // reduction. */
int clipebc_fold_average(const float* preds_dev, const int* row_cells_host, const int* col_cells_host, int n_rows,
                         int n_cols, int gh, int gw, int Ho, int Wo, float* density_out_dev, float* count_out_dev,
                         void* stream);

/* ---- the steps either side of the hot path ------------------------------------------------
 * Pre-step: what datasets/crowd.py:213-228 does to an image before sliding_window_predict -- uint8 -> [0,1]
 * (`/ 255.`), Resize2Multiple (datasets/transforms.py:69-105: TF.resize BICUBIC antialias=True) or ZeroPad2Multiple
 * (:108-140: right/bottom zero pad), torchvision Normalize (datasets/crowd.py:64). Post-step: resize_density_map
 * (utils/eval_utils.py:19-23).
 *
 * in_dev: [C, h, w], uint8 when in_is_u8 != 0, else f32 already in [0,1]; C <= 4. mean/std: HOST arrays of C floats, or
 * both NULL for no normalisation. tmp_dev: f32 [C, h, W] scratch. out_dev: f32 [C, H, W]. */
int clipebc_resize_bicubic_aa(const void* in_dev, int in_is_u8, int C, int h, int w, float* tmp_dev, float* out_dev,
                              int H, int W, const float* mean_host, const float* std_host, void* stream);
int clipebc_pad_normalize(const void* in_dev, int in_is_u8, int C, int h, int w, float* out_dev, int H, int W,
                          const float* mean_host, const float* std_host, void* stream);
/* x_dev f32 [h, w] (one density map, the only shape the reference function broadcasts for) -> out_dev f32 [H, W] =
 * bilinear(x) * nan_to_num(sum(bilinear(x)) / sum(x), 0, 0, 0). workspace_dev: clipebc_resize_density_workspace_floats()
 * floats. sums_out_dev (nullable): [sum(x), sum(bilinear(x))]. Sums are two-stage with a fixed order (deterministic). */
int clipebc_resize_density_workspace_floats(void);
int clipebc_resize_density_map(const float* x_dev, int h, int w, int H, int W, float* out_dev, float* workspace_dev,
                               float* sums_out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIPEBC_B200_H_ */
