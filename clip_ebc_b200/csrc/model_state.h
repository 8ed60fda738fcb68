// Internal (C++) state of a model handle, shared by the translation units that implement the C-ABI: api.cu (ViT path,
// entry points) and resnet_path.cu (CLIP-ResNet encoder path). Not part of the public boundary (include/clipebc_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <atomic>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/clipebc_b200.h"
#include "kernels.h"

namespace cebc {

// error convention of the C-ABI: a code + a per-thread message (clipebc_last_error)
int fail(int code, const std::string& msg);
int fail_cuda(cudaError_t e, const char* what);
#define CUDA_TRY(expr)                                         \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::cebc::fail_cuda(_e, #expr); \
  } while (0)
// kernel launchers return nullptr or a message
#define K_TRY(expr)                                                                \
  do {                                                                             \
    const char* _m = (expr);                                                       \
    if (_m != nullptr) return ::cebc::fail(CLIPEBC_ECUDA, std::string(_m));        \
  } while (0)

// Bumped by every (re)allocation or release of a device buffer of this library: host layers that cache captured CUDA
// graphs of the library's launches key them on it (a graph holds the buffer addresses used at capture time -- replaying
// it after a workspace has moved would write to freed memory).
extern std::atomic<int64_t> g_config_epoch;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { if (p) { cudaFree(p); g_config_epoch.fetch_add(1); } }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  cudaError_t reserve(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; bytes = 0; }
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) bytes = n;
    g_config_epoch.fetch_add(1);
    return e;
  }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

struct RawTensor {
  DevBuf buf;
  std::vector<int64_t> shape;
  int64_t numel = 0;
};

struct LayerPack {
  DevBuf w_qkv, w_out, w_fc, w_proj;  // 16-bit, nn.Linear layout [N, K]
  DevBuf const_kv;                    // bf16 [num_vpt, 3 * width] (deep VPT): in_proj(LN1_l(vpt_l)), input-independent
  const float *b_qkv, *b_out, *b_fc, *b_proj, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
};

// ---- CLIP-ResNet encoder path (resnet_path.cu) ----------------------------------------------------------------------
// One convolution with its BatchNorm folded in, as a GEMM operand: W 16-bit [n_pad, k] K-major (k = taps * per-tap padded
// input channels, or the concatenation [conv3 | downsample] of a block's last GEMM), bias f32 [n_pad]
struct ConvPack {
  DevBuf w, b;
  int n = 0, n_pad = 0, k = 0;
};
// Bottleneck of the encoder (_clip/blocks.py:56-101) or of the decoder (models/utils.py:334-390): 1x1 -> 3x3 -> [avgpool] ->
// 1x1 (+ identity, or + BN(conv1x1([avgpool] x)) folded into the same GEMM by concatenating the operands along K)
struct RnBlock {
  std::string prefix;
  int c_in = 0, planes = 0, c_out = 0, stride = 1;
  bool down = false;
  ConvPack c1, c2, c3;
};
struct ResNetPack {
  ConvPack stem1, stem2, stem3;
  std::vector<std::unique_ptr<RnBlock>> enc, dec;  // encoder blocks in execution order; decoder blocks
  int enc_reduction = 16, stem_width = 64, c_feat = 0, c_dec = 0, e_pad = 0;
  DevBuf w_proj, b_proj;                           // 16-bit [e_pad, k64(c_dec)], f32 [e_pad]: embed_dim padded to 256 with zeros
  // workspaces of one pass (activations are 16-bit NHWC on shared-border grids)
  DevBuf col, s1, s2, s3, xa, xb, t1, t2, cc, up;
};

}  // namespace cebc

struct clipebc_model {
  clipebc_config cfg;       // normalised: patch / width / layers / embed_dim filled in
  int device = 0;           // the CUDA device the handle was created on: every buffer below lives there
  int kp_pad = 0;           // patch row length 3 * patch^2 rounded up to the GEMM's K granularity (64)
  // Split precision (hi + lo operands, three K-segments [hi | lo | hi] x [Whi | Whi | Wlo]) of the two GEMMs whose rounding
  // reaches the head directly: patch embedding and the 1x1 projection. On for bf16 operands (8-bit mantissa: a single
  // segment costs 0.6-1.3e-2 on the logits and 0.1-0.3 % of the bin argmax at the projection alone); off for fp16 operands
  // (11 bits: 0.7-1.5e-3 and >= 99.97 %, measured against the fp32 oracle stage by stage) -- there the split would be 1.8 of
  // 48.8 GFLOP per window spent on bits the other twelve 16-bit roundings per block have already given up.
  bool split_precision = false;
  std::map<std::string, cebc::RawTensor> raw;
  bool packed = false;
  // packed (ViT path)
  std::unique_ptr<cebc::LayerPack[]> layer;  // [cfg.layers]
  cebc::DevBuf w_patch;           // 16-bit [width, 3 * kp_pad] = hi | hi | lo
  cebc::DevBuf w_c1z;             // 16-bit [9 * width, width]: conv1 with the tap on the output side (coarse-grid form)
  cebc::DevBuf zero_bias;         // f32 [9 * width] zeros
  cebc::DevBuf ws_Y16, ws_Z;      // coarse-grid conv1: 16-bit ln_post rows, per-tap products [n * hp * wp, 9 * width]
  cebc::DevBuf w_c1, w_c2;        // 16-bit [width, 9 * width]
  cebc::DevBuf b_c1, b_c2;        // f32 [width]
  cebc::DevBuf w_p3;              // 16-bit [embed, 3 * width] = hi | hi | lo
  cebc::DevBuf tmat;              // f32 [N, embed]
  cebc::DevBuf pack_tmp_16;
  std::map<int, cebc::DevBuf> pos_cache;  // key hp * 4096 + wp -> f32 [1 + hp*wp, width]; bounded
  // packed (CLIP-ResNet path)
  std::unique_ptr<cebc::ResNetPack> resnet;
  // workspace
  cebc::DevBuf ws_patch_rows, ws_patch_embed, ws_X, ws_Xn, ws_QKV, ws_AO, ws_Hid, ws_Y, ws_Ub, ws_Uf, ws_D1, ws_D2, ws_F,
      ws_preds;
  // device-resident index tables (window -> patch-grid row, window origins, fold cell origins), cached per geometry so
  // the steady state has no host->device upload and no host synchronisation; bounded
  std::map<std::string, cebc::DevBuf> idx_cache;
};

namespace cebc {

// Tensor by its state_dict key, or nullptr: pack() has checked every key the path reads, so a miss can only be a tensor
// the configuration does not need (never throws across the C ABI)
const float* raw_ptr(clipebc_model* m, const std::string& name);
bool check_shape(clipebc_model* m, const std::string& name, std::initializer_list<int64_t> want, std::string* err);
const char* gemm_dispatch(cudaStream_t stream, int epi, const __nv_bfloat16* A, int64_t a_rows, int64_t a_cols, int64_t lda,
                          const __nv_bfloat16* W, int64_t ldw, GemmParams p, int block_n);

// ---- resnet_path.cu ---------------------------------------------------------------------------------------------------
// Validate the loaded tensors of a CLIP-ResNet model, fold every BatchNorm into its convolution and build the 16-bit GEMM
// operands (m->resnet). Called by clipebc_model_pack when cfg.encoder == 1.
int resnet_pack(clipebc_model* m, cudaStream_t s);
// model(x) for `nw` windows of h x w pixels: units are whole images of a batch [nw, 3, h, w] (origins_yx_dev == nullptr) or
// windows of ONE image [3, H, W] with origins origins_yx_dev[2 * i] / [2 * i + 1]. exp_out f32 [nw, 1, h/r, w/r];
// logits_out (nullable) f32 [nw, N, h/r, w/r].
int resnet_run_windows(clipebc_model* m, cudaStream_t s, const float* image_dev, int H, int W, const int* origins_yx_dev,
                       int nw, int h, int w, float* exp_out, float* logits_out);
int resnet_default_chunk(const clipebc_model* m, int h, int w);

}  // namespace cebc
