// C-ABI of the CLIP-EBC hot path (include/clipebc_b200.h): model object (reference state_dict in, packed device
// weights out), the two reference-facing entry points (model(x), sliding_window_predict) and single-kernel test hooks.
// Host orchestration only -- every arithmetic step runs in the sm_100a kernels of this directory.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/clipebc_b200.h"
#include "kernels.h"
#include "model_state.h"

namespace cebc {

static std::atomic<int64_t> g_launches{0};

// ---- launch accounting / optional per-launch CUDA-event profiling (bench.py roofline breakdown) -------------------
namespace {
struct ProfRec {
  std::string tag;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double flops = 0, bytes = 0;
};
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
thread_local const char* g_tag = nullptr;
}  // namespace

void set_launch_tag(const char* tag) { g_tag = tag; }
bool profiling_on() { return g_prof_on; }

bool pdl_enabled() {
  static const bool on = std::getenv("CLIPEBC_NO_PDL") == nullptr;
  return on;
}

LaunchScope::LaunchScope(cudaStream_t stream, const char* kind, double flops, double bytes) : stream_(stream), slot_(-1) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  r.tag = g_tag ? std::string(kind) + ":" + g_tag : std::string(kind);
  r.flops = flops; r.bytes = bytes;
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, stream);
  g_prof.push_back(r);
  slot_ = static_cast<int>(g_prof.size()) - 1;
}
LaunchScope::~LaunchScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot_ < static_cast<int>(g_prof.size())) cudaEventRecord(g_prof[slot_].e1, stream_);
}

const char* gemm_dispatch(cudaStream_t stream, int epi, const __nv_bfloat16* A, int64_t a_rows, int64_t a_cols, int64_t lda,
                          const __nv_bfloat16* W, int64_t ldw, GemmParams p, int block_n) {
  return gemm2_bf16_tn(stream, epi, A, a_rows, a_cols, lda, W, ldw, p, block_n);
}

// tcgen05 / TMEM kernels: windows of at most 256 keys whose constant-key count is a multiple of 8 (every stock configuration
// of ViT-B/16 and ViT-B/32) take the one-block kernel, windows of 257..320 keys (ViT-L/14: 289 keys per 224 window) the
// two-block one; the streamed-K/V kernel takes everything else (windows larger than 224 x 224, odd prompt counts)
const char* attention_dispatch(cudaStream_t stream, const __nv_bfloat16* qkv, const __nv_bfloat16* const_kv, int n_const,
                               int n_win, int t_live, int heads, void* out, int out_fp16) {
  if (t_live + n_const <= 256 && n_const % 8 == 0)
    return attention_h64_pp(stream, qkv, const_kv, n_const, n_win, t_live, heads, out, out_fp16);
  if (attention_h64_ppl_takes(n_const, t_live))
    return attention_h64_ppl(stream, qkv, const_kv, n_const, n_win, t_live, heads, out, out_fp16);
  return attention_h64_long(stream, qkv, const_kv, n_const, n_win, t_live, heads, out, out_fp16);
}

namespace { thread_local std::string g_err; }
namespace { thread_local const cudaAccessPolicyWindow* g_l2_window_tls = nullptr; }
const cudaAccessPolicyWindow* current_l2_window() { return g_l2_window_tls; }
static void set_l2_window(const cudaAccessPolicyWindow* w) { g_l2_window_tls = w; }

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int fail_cuda(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return CLIPEBC_ECUDA;
}
std::atomic<int64_t> g_config_epoch{0};

}  // namespace cebc

using namespace cebc;

namespace cebc {

const float* raw_ptr(clipebc_model* m, const std::string& name) {
  auto it = m->raw.find(name);
  return it == m->raw.end() ? nullptr : it->second.buf.as<float>();
}

bool check_shape(clipebc_model* m, const std::string& name, std::initializer_list<int64_t> want, std::string* err) {
  auto it = m->raw.find(name);
  if (it == m->raw.end()) { *err = "missing tensor '" + name + "'"; return false; }
  const auto& s = it->second.shape;
  std::vector<int64_t> w(want);
  if (s != w) {
    std::string got = "[", exp = "[";
    for (auto v : s) got += std::to_string(v) + ",";
    for (auto v : w) exp += std::to_string(v) + ",";
    *err = "tensor '" + name + "' has shape " + got + "] expected " + exp + "]";
    return false;
  }
  return true;
}

}  // namespace cebc

namespace {

constexpr size_t kMaxPosCache = 16, kMaxIdxCache = 64;

// Every entry point that touches a handle: the handle's buffers live on the device it was created on, and the kernels'
// per-device attributes are keyed on the current device, so the two must agree.
int check_device(const clipebc_model* m);

bool check_shape_ok(clipebc_model* m, const std::string& scalar_name) {
  auto it = m->raw.find(scalar_name);
  return it != m->raw.end() && it->second.numel == 1;
}

std::string blk(int l, const char* tail) {
  return "image_encoder.transformer.resblocks." + std::to_string(l) + "." + tail;
}

int to_16(cudaStream_t s, const float* src, int64_t n, DevBuf* dst, int fp16) {
  CUDA_TRY(dst->reserve(static_cast<size_t>(n) * 2));
  K_TRY(f32_to_16(s, src, dst->p, n, fp16));
  return CLIPEBC_OK;
}

// PyTorch upsample_bicubic2d (align_corners=False, A=-0.75) of the patch part of the positional embedding
// (reference: _clip/image_encoder.py:183-198). Runs once per (hp, wp) on the host.
void cubic_coeffs(float t, float w[4]) {
  const float A = -0.75f;
  auto c1 = [&](float x) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; };
  auto c2 = [&](float x) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; };
  w[0] = c2(t + 1.f); w[1] = c1(t); w[2] = c1(1.f - t); w[3] = c2(2.f - t);
}

int get_pos(clipebc_model* m, int hp, int wp, cudaStream_t stream, const float** out) {
  const int g0 = m->cfg.input_size / m->cfg.patch, D = m->cfg.width;
  if (hp == g0 && wp == g0) { *out = raw_ptr(m, "image_encoder.positional_embedding"); return CLIPEBC_OK; }
  const int key = hp * 4096 + wp;
  auto it = m->pos_cache.find(key);
  if (it != m->pos_cache.end()) { *out = it->second.as<float>(); return CLIPEBC_OK; }
  std::vector<float> src(static_cast<size_t>(1 + g0 * g0) * D);
  CUDA_TRY(cudaMemcpy(src.data(), raw_ptr(m, "image_encoder.positional_embedding"), src.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<float> dst(static_cast<size_t>(1 + hp * wp) * D);
  std::memcpy(dst.data(), src.data(), static_cast<size_t>(D) * 4);
  const float sy = static_cast<float>(g0) / hp, sx = static_cast<float>(g0) / wp;
  for (int oy = 0; oy < hp; ++oy) {
    const float fy = (oy + 0.5f) * sy - 0.5f;
    const int iy = static_cast<int>(std::floor(fy));
    float wy[4]; cubic_coeffs(fy - iy, wy);
    for (int ox = 0; ox < wp; ++ox) {
      const float fx = (ox + 0.5f) * sx - 0.5f;
      const int ix = static_cast<int>(std::floor(fx));
      float wx[4]; cubic_coeffs(fx - ix, wx);
      float* o = &dst[static_cast<size_t>(1 + oy * wp + ox) * D];
      for (int c = 0; c < D; ++c) o[c] = 0.f;
      for (int a = 0; a < 4; ++a) {
        const int yy = std::min(std::max(iy - 1 + a, 0), g0 - 1);
        for (int b = 0; b < 4; ++b) {
          const int xx = std::min(std::max(ix - 1 + b, 0), g0 - 1);
          const float wgt = wy[a] * wx[b];
          const float* sp = &src[static_cast<size_t>(1 + yy * g0 + xx) * D];
          for (int c = 0; c < D; ++c) o[c] += wgt * sp[c];
        }
      }
    }
  }
  // bounded: a stream of differently sized windows must not grow the cache without limit. Entries are only dropped
  // between calls that use them (the stream is synchronised below), never while a launch may still read one.
  if (m->pos_cache.size() >= kMaxPosCache) {
    CUDA_TRY(cudaStreamSynchronize(stream));
    m->pos_cache.clear();
  }
  DevBuf& buf = m->pos_cache[key];
  CUDA_TRY(buf.reserve(dst.size() * 4));
  CUDA_TRY(cudaMemcpyAsync(buf.p, dst.data(), dst.size() * 4, cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  *out = buf.as<float>();
  return CLIPEBC_OK;
}

// fp16: 16-bit format of the operands (A, W); out16_fp16: format of a 16-bit output (QKV stays bf16 for the attention)
GemmParams plain(int fp16, int out16_fp16, int M, int N, int K, void* out, int ldo, const float* bias,
                 const float* resid = nullptr, int ldr = 0) {
  GemmParams p = gemm_params_plain(M, N, K);
  p.out = out; p.ldo = ldo; p.bias = bias; p.resid = resid; p.ldr = ldr;
  p.ab_fp16 = fp16; p.out_fp16 = out16_fp16;
  return p;
}

// patch embedding: split precision (rows [hi | lo] (2 * kp) x W3 = [Whi | Whi | Wlo] (3 * kp), segments hi, lo, hi) or a
// single segment (rows [hi] (kp) x the Whi columns of W3); kp = 3 * patch^2 rounded up to 64 (768 for ViT-B/16, 3072 for
// ViT-B/32, 640 for ViT-L/14)
GemmParams patch_embed_params(int fp16, bool split, int rows, int width, void* out, int kp) {
  GemmParams p = gemm_params_plain(rows, width, (split ? 3 : 1) * kp);
  if (split) {
    p.n_seg = 3; p.seg_kblocks = kp / 64;
    p.seg_col_start[0] = 0; p.seg_col_start[1] = kp; p.seg_col_start[2] = 0;
  }
  p.out = out; p.ldo = width; p.bias = nullptr; p.ab_fp16 = fp16; p.out_fp16 = fp16;
  return p;
}

// RAII: persisting-L2 access-policy window over [ptr, ptr + bytes), attached as a LAUNCH attribute to every kernel this
// thread launches while it lives (the caller's stream state is not touched). The per-device carve-out
// (cudaLimitPersistingL2CacheSize) only ever grows, up to kL2PersistCapMB / the device maximum.
constexpr int kL2PersistCapMB = 60;
struct L2Window {
  cudaAccessPolicyWindow win_ = {};
  bool active_ = false;
  L2Window(const void* ptr, size_t bytes) {
    static const int env_mb = std::getenv("CLIPEBC_L2_PERSIST") ? std::atoi(std::getenv("CLIPEBC_L2_PERSIST")) : -1;
    if (env_mb == 0 || bytes == 0 || cebc::current_l2_window() != nullptr) return;
    // the carve-out limit and its maximum are per device
    constexpr int kMaxDev = 64;
    static std::mutex mu;
    static long long max_persist[kMaxDev];  // 0 = not queried yet, -1 = none
    static long long max_window[kMaxDev];   // largest access-policy window of the device (a larger one fails the LAUNCH)
    static size_t limit_now[kMaxDev];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) { cudaGetLastError(); return; }
    std::lock_guard<std::mutex> lock(mu);
    if (max_persist[dev] == 0) {
      int v = 0;
      if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, dev) != cudaSuccess) { cudaGetLastError(); v = 0; }
      max_persist[dev] = v > 0 ? v : -1;
      int w = 0;
      if (cudaDeviceGetAttribute(&w, cudaDevAttrMaxAccessPolicyWindowSize, dev) != cudaSuccess) { cudaGetLastError(); w = 0; }
      max_window[dev] = w;
    }
    if (max_window[dev] <= 0) return;
    bytes = std::min(bytes, static_cast<size_t>(max_window[dev]));  // a buffer beyond the limit gets a window over its head
    size_t carve = std::min(bytes, static_cast<size_t>(env_mb > 0 ? env_mb : kL2PersistCapMB) << 20);
    carve = std::min(carve, static_cast<size_t>(std::max(0ll, max_persist[dev])));
    if (carve == 0) return;
    if (carve > limit_now[dev]) {
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) { cudaGetLastError(); return; }
      limit_now[dev] = carve;
    }
    win_.base_ptr = const_cast<void*>(ptr);
    win_.num_bytes = bytes;
    win_.hitRatio = static_cast<float>(std::min(1.0, static_cast<double>(carve) / static_cast<double>(bytes)));
    win_.hitProp = cudaAccessPropertyPersisting;
    win_.missProp = cudaAccessPropertyStreaming;
    set_l2_window(&win_);
    active_ = true;
  }
  ~L2Window() { if (active_) set_l2_window(nullptr); }
  L2Window(const L2Window&) = delete;
  L2Window& operator=(const L2Window&) = delete;
};

// The ViT blocks + decoder + head for `nw` windows whose patch embeddings are already in m->ws_patch_embed.
int run_windows(clipebc_model* m, cudaStream_t s, const int* win_base_dev, int src_pitch, int nw, int hp, int wp,
                const float* pos, float* exp_out, float* logits_out, const int* win_pitch_dev = nullptr) {
  const clipebc_config& c = m->cfg;
  const int D = c.width, heads = D / 64, hidden = 4 * D, E = c.embed_dim;
  const bool deep = c.deep_vpt != 0;
  const int fp16 = c.operand_fp16 != 0;   // 16-bit operand format of every GEMM of the path
  const int ln16 = fp16 ? 2 : 1;          // layernorm_rows out_kind
  const int n_prompt_live = deep ? 0 : c.num_vpt;
  const int n_const = deep ? c.num_vpt : 0;
  const int npatch = hp * wp;
  const int T = 1 + n_prompt_live + npatch;
  const int M = nw * T;
  const int gh = hp * c.patch / c.reduction, gw = wp * c.patch / c.reduction;
  const int Hp = gh + 1, Wp = gw + 1;  // shared-border decoder grid (kernels.h: resample_to_padded): 841 rows per r8
  const int Mp = nw * Hp * Wp;         // window instead of 900 on a grid bordered on all four sides
  // conv1 from the coarse grid (elementwise.cu: conv1_from_coarse_kernel) whenever the decoder grid is at least twice as
  // fine as the patch grid (reduction 8 with ViT-B/16; 8 / 16 with ViT-B/32); else the implicit GEMM on the fine grid
  const bool coarse1 = c.decoder_conv1_fine == 0 && gh >= 2 * hp && gw >= 2 * wp;

  CUDA_TRY(m->ws_X.reserve(static_cast<size_t>(M) * D * 4));
  CUDA_TRY(m->ws_Xn.reserve(static_cast<size_t>(M) * D * 2));
  CUDA_TRY(m->ws_QKV.reserve(static_cast<size_t>(M) * 3 * D * 2));
  CUDA_TRY(m->ws_AO.reserve(static_cast<size_t>(M) * D * 2));
  CUDA_TRY(m->ws_Hid.reserve(static_cast<size_t>(M) * hidden * 2));
  CUDA_TRY(m->ws_Y.reserve(static_cast<size_t>(nw) * npatch * D * 4));
  if (coarse1) {
    CUDA_TRY(m->ws_Y16.reserve(static_cast<size_t>(nw) * npatch * D * 2));
    CUDA_TRY(m->ws_Z.reserve(static_cast<size_t>(nw) * npatch * 9 * D * 2));
  } else {
    CUDA_TRY(m->ws_Ub.reserve(static_cast<size_t>(Mp) * D * 2));
    CUDA_TRY(m->ws_Uf.reserve(static_cast<size_t>(Mp) * D * 4));
  }
  CUDA_TRY(m->ws_D1.reserve(static_cast<size_t>(Mp) * D * 2));
  const bool split = m->split_precision;
  const int d2_cols = (split ? 2 : 1) * D;  // conv2 output: hi | lo, or hi only
  CUDA_TRY(m->ws_D2.reserve(static_cast<size_t>(Mp) * d2_cols * 2));
  const int kParts = 2 * (E / 256);  // head partials per cell: one per half of a 256-wide projection tile
  CUDA_TRY(m->ws_F.reserve(static_cast<size_t>(Mp) * kParts * (1 + c.num_bins) * 4));

  float* X = m->ws_X.as<float>();
  // The fp32 residual stream is read / updated four times per block while QKV (58 MB at 64 windows) and Hid (77 MB)
  // stream through L2 once: an access-policy window on the launching stream keeps X in the persisting carve-out of L2
  // for the duration of the pass (measured on B200: +4 % on 2048x1536 images, +0.3..1.3 % at 64 windows;
  // profiles/r01i_l2_persist.txt). CLIPEBC_L2_PERSIST=<MB> overrides the carve-out (0 = off).
  const L2Window l2win(X, static_cast<size_t>(M) * D * 4);
  __nv_bfloat16* Xn = m->ws_Xn.as<__nv_bfloat16>();
  __nv_bfloat16* QKV = m->ws_QKV.as<__nv_bfloat16>();
  __nv_bfloat16* AO = m->ws_AO.as<__nv_bfloat16>();
  __nv_bfloat16* Hid = m->ws_Hid.as<__nv_bfloat16>();

  K_TRY(assemble_tokens(s, D, m->ws_patch_embed.as<float>(), win_base_dev, src_pitch, win_pitch_dev,
                        raw_ptr(m, "image_encoder.class_embedding"), pos, raw_ptr(m, "image_encoder.ln_pre.weight"),
                        raw_ptr(m, "image_encoder.ln_pre.bias"), n_prompt_live > 0 ? raw_ptr(m, "vpt_0") : nullptr,
                        n_prompt_live, nw, hp, wp, X));

  for (int l = 0; l < c.layers; ++l) {
    const LayerPack& L = m->layer[l];
    K_TRY(layernorm_rows(s, D, X, L.ln1_g, L.ln1_b, Xn, ln16, M, 1, 1, 0));
    set_launch_tag("qkv");
    K_TRY(gemm_dispatch(s, EPI_BIAS_BF16, Xn, M, D, D, L.w_qkv.as<__nv_bfloat16>(), D,
                        plain(fp16, 0, M, 3 * D, D, QKV, 3 * D, L.b_qkv), 0));
    set_launch_tag(nullptr);
    K_TRY(attention_dispatch(s, QKV, n_const > 0 ? L.const_kv.as<__nv_bfloat16>() : nullptr, n_const, nw, T, heads, AO, fp16));
    set_launch_tag("out_proj");
    K_TRY(gemm_dispatch(s, EPI_BIAS_RESID_F32, AO, M, D, D, L.w_out.as<__nv_bfloat16>(), D,
                        plain(fp16, fp16, M, D, D, X, D, L.b_out, X, D), 0));
    set_launch_tag(nullptr);
    K_TRY(layernorm_rows(s, D, X, L.ln2_g, L.ln2_b, Xn, ln16, M, 1, 1, 0));
    set_launch_tag("c_fc");
    K_TRY(gemm_dispatch(s, EPI_BIAS_GELU_BF16, Xn, M, D, D, L.w_fc.as<__nv_bfloat16>(), D,
                        plain(fp16, fp16, M, hidden, D, Hid, hidden, L.b_fc), 0));
    set_launch_tag("c_proj");
    K_TRY(gemm_dispatch(s, EPI_BIAS_RESID_F32, Hid, M, hidden, hidden, L.w_proj.as<__nv_bfloat16>(), hidden,
                        plain(fp16, fp16, M, D, hidden, X, D, L.b_proj, X, D), 0));
    set_launch_tag(nullptr);
  }

  // ln_post on the patch rows only (cls / prompt rows are dropped, model.py:185-188), fp32 out; the coarse-grid conv1 also
  // gets the 16-bit rows its GEMM reads
  float* Y = m->ws_Y.as<float>();
  K_TRY(layernorm_rows(s, D, X, raw_ptr(m, "image_encoder.ln_post.weight"), raw_ptr(m, "image_encoder.ln_post.bias"), Y, 0,
                       static_cast<int64_t>(nw) * npatch, npatch, T, T - npatch, coarse1 ? m->ws_Y16.p : nullptr, fp16));
  __nv_bfloat16* Ub = m->ws_Ub.as<__nv_bfloat16>();
  float* Uf = m->ws_Uf.as<float>();
  // coarse-grid conv1: the fine-grid map is never materialised -- conv1 reads the per-tap products on the patch grid and
  // conv2's epilogue evaluates the BasicBlock skip (bilinear_up(Y)) on the fly (EPI_BIAS_UPSKIP_RELU_SPLIT)
  if (!coarse1) K_TRY(resample_to_padded(s, D, Y, nw, hp, wp, gh, gw, Ub, Uf, fp16));

  // decoder BasicBlock as two implicit GEMMs over the zero-bordered grid: 9 taps = 9 row-shifted K-segments
  GemmParams pc = gemm_params_plain(Mp, D, 9 * D);
  pc.ab_fp16 = fp16; pc.out_fp16 = fp16;
  pc.n_seg = 9; pc.seg_kblocks = D / 64;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      pc.seg_row_shift[ky * 3 + kx] = (ky - 1) * Wp + (kx - 1);
      pc.seg_col_start[ky * 3 + kx] = 0;
    }
  __nv_bfloat16* D1 = m->ws_D1.as<__nv_bfloat16>();
  __nv_bfloat16* D2 = m->ws_D2.as<__nv_bfloat16>();
  set_launch_tag("dec_conv1");
  if (coarse1) {
    const int64_t rows_c = static_cast<int64_t>(nw) * npatch;
    K_TRY(gemm_dispatch(s, EPI_BIAS_BF16, m->ws_Y16.as<__nv_bfloat16>(), rows_c, D, D, m->w_c1z.as<__nv_bfloat16>(), D,
                        plain(fp16, fp16, static_cast<int>(rows_c), 9 * D, D, m->ws_Z.p, 9 * D, m->zero_bias.as<float>()), 0));
    set_launch_tag(nullptr);
    K_TRY(conv1_from_coarse(s, D, m->ws_Z.p, m->b_c1.as<float>(), nw, hp, wp, gh, gw, D1, fp16));
  } else {
    GemmParams p1 = pc;
    p1.out = D1; p1.ldo = D; p1.bias = m->b_c1.as<float>(); p1.mask_hp = Hp; p1.mask_wp = Wp; p1.mask_lead = 0;
    K_TRY(gemm_dispatch(s, EPI_BIAS_RELU_MASK_BF16, Ub, Mp, D, D, m->w_c1.as<__nv_bfloat16>(), 9 * D, p1, 0));
  }
  GemmParams p2 = pc;
  p2.out = D2; p2.ldo = d2_cols; p2.split_lo = split; p2.bias = m->b_c2.as<float>(); p2.resid = Uf; p2.ldr = D;
  if (coarse1) { p2.resid = Y; p2.mask_hp = Hp; p2.mask_wp = Wp; p2.up_hp = hp; p2.up_wp = wp; }
  set_launch_tag("dec_conv2");
  K_TRY(gemm_dispatch(s, coarse1 ? EPI_BIAS_UPSKIP_RELU_SPLIT : EPI_BIAS_RESID_RELU_SPLIT, D1, Mp, D, D,
                      m->w_c2.as<__nv_bfloat16>(), 9 * D, p2, 0));

  // projection 1x1 -- in split precision, [hi | lo | hi] x [Whi | Whi | Wlo] (the A segments re-use the hi columns), or one
  // segment hi x Whi -- fused with the head: the `embed` projected features of a cell are never written; the GEMM epilogue
  // leaves ||f||^2 and the N bin dot products per half tile (kParts partials per row), ebc_head_finish does the rest
  GemmParams pp = gemm_params_plain(Mp, E, (split ? 3 : 1) * D);
  pp.ab_fp16 = fp16; pp.out_fp16 = fp16;
  if (split) {
    pp.n_seg = 3; pp.seg_kblocks = D / 64;
    pp.seg_col_start[0] = 0; pp.seg_col_start[1] = D; pp.seg_col_start[2] = 0;
  }
  float* F = m->ws_F.as<float>();
  pp.bias = raw_ptr(m, "projection.bias");
  pp.out = F; pp.ldo = kParts * (1 + c.num_bins);
  pp.head_tmat = m->tmat.as<float>(); pp.head_bins = c.num_bins;
  set_launch_tag("projection+head");
  K_TRY(gemm_dispatch(s, EPI_BIAS_HEAD_PARTIAL, D2, Mp, d2_cols, d2_cols, m->w_p3.as<__nv_bfloat16>(), 3 * D, pp, 256));
  set_launch_tag(nullptr);
  K_TRY(ebc_head_finish(s, F, kParts, raw_ptr(m, "anchor_points"), c.num_bins, nw, gh, gw, exp_out, logits_out));
  return CLIPEBC_OK;
}

// windows per internal pass: 148 windows of 197 / 229 tokens by default (profiles/r02/chunk_sweep.txt: 96 / 128 / 148 / 192
// windows per pass give 75.1 / 75.4 / 76.6 / 76.3 images/s on 2048x1536 at stride 112, 66.4 / 64.5 / 67.2 / 67.5 on 4096x3072 at
// stride 224; an equalised split of an image's windows is not better); windows with more tokens get proportionally fewer per
// pass so that the workspaces stay the same size
int default_chunk(const clipebc_model* m, int hp, int wp) {
  if (m->cfg.window_chunk > 0) return m->cfg.window_chunk;
  const int64_t tokens = 1 + m->cfg.num_vpt + static_cast<int64_t>(hp) * wp;
  if (m->cfg.width > 768) {
    // ViT-L/14: as many windows as fill one 256-row tile per CTA pair (74 row tiles), so that every GEMM of a pass runs whole
    // rounds whatever its N: 73 windows of 257 live rows (65 of 289 with shallow VPT). profiles/r02/l14_chunk_sweep.txt: 292
    // windows at 48 / 58 / 64 / 73 / 96 / 146 per pass -> 4538 / 4480 / 4627 / 4776 / 4589 / 4680 windows/s
    const int64_t live = m->cfg.deep_vpt ? 1 + static_cast<int64_t>(hp) * wp : tokens;
    return static_cast<int>(std::max<int64_t>(1, static_cast<int64_t>(device_num_sms() / 2) * 256 / live));
  }
  if (tokens <= 128) return 256;  // ViT-B/32 windows (82 tokens)
  if (tokens <= 256) return 148;
  return static_cast<int>(std::max<int64_t>(1, 148 * 229 / tokens));
}

int check_window_geometry(clipebc_model* m, int h, int w) {
  if (m->cfg.encoder == 1) {
    if (h <= 0 || w <= 0 || h % 32 != 0 || w % 32 != 0)
      return fail(CLIPEBC_EINVAL, "window height/width must be positive multiples of 32 for the CLIP-ResNet encoders");
    if (h > 2048 || w > 2048) return fail(CLIPEBC_EINVAL, "window too large for the CLIP-ResNet path (max 2048 pixels a side)");
    return CLIPEBC_OK;
  }
  const int kPatch = m->cfg.patch;
  if (h <= 0 || w <= 0 || h % kPatch != 0 || w % kPatch != 0)
    return fail(CLIPEBC_EINVAL, "window height/width must be positive multiples of the patch size");
  if ((h % m->cfg.reduction) != 0 || (w % m->cfg.reduction) != 0)
    return fail(CLIPEBC_EINVAL, "window height/width must be multiples of the reduction");
  const int64_t T = 1 + m->cfg.num_vpt + static_cast<int64_t>(h / kPatch) * (w / kPatch);
  if (T > 16384) return fail(CLIPEBC_EINVAL, "window too large: 1 + num_vpt + patches must be <= 16384 tokens");
  return CLIPEBC_OK;
}

int check_device(const clipebc_model* m) {
  int dev = -1;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev != m->device)
    return fail(CLIPEBC_ESTATE, "the model handle lives on CUDA device " + std::to_string(m->device) + " but the current device is " +
                                    std::to_string(dev) + ": make its device current, or create a new handle on this one");
  return CLIPEBC_OK;
}

// patch rows of a call: [rows, 2 * kp_pad] 16-bit; when 3 * patch^2 is not a multiple of 64 (ViT-L/14: 588 -> 640) the
// pad columns must be zero, and since patchify never writes them they are cleared whenever the buffer is (re)allocated
int reserve_patch_rows(clipebc_model* m, int64_t rows, cudaStream_t s) {
  const size_t bytes = static_cast<size_t>(rows) * (m->split_precision ? 2 : 1) * m->kp_pad * 2;
  const void* before = m->ws_patch_rows.p;
  CUDA_TRY(m->ws_patch_rows.reserve(bytes));
  if (m->ws_patch_rows.p != before && m->kp_pad != 3 * m->cfg.patch * m->cfg.patch)
    CUDA_TRY(cudaMemsetAsync(m->ws_patch_rows.p, 0, m->ws_patch_rows.bytes, s));
  CUDA_TRY(m->ws_patch_embed.reserve(static_cast<size_t>(rows) * m->cfg.width * 4));
  return CLIPEBC_OK;
}

// stem of a call: patch rows are in ws_patch_rows -> patch embeddings f32 [rows, width] in ws_patch_embed
int patch_embed(clipebc_model* m, cudaStream_t s, int64_t rows) {
  const int fp16 = m->cfg.operand_fp16 != 0, kp = m->kp_pad;
  set_launch_tag("patch_embed");
  const int a_cols = (m->split_precision ? 2 : 1) * kp;
  K_TRY(gemm_dispatch(s, EPI_F32, m->ws_patch_rows.as<__nv_bfloat16>(), rows, a_cols, a_cols, m->w_patch.as<__nv_bfloat16>(),
                      3 * kp, patch_embed_params(fp16, m->split_precision, static_cast<int>(rows), m->cfg.width,
                                                 m->ws_patch_embed.p, kp), 0));
  set_launch_tag(nullptr);
  return CLIPEBC_OK;
}

// device copy of a host index table, cached under `key`
int cached_table(clipebc_model* m, const std::string& key, const std::vector<int>& tab, cudaStream_t s, const int** out) {
  auto it = m->idx_cache.find(key);
  if (it == m->idx_cache.end()) {
    if (m->idx_cache.size() >= kMaxIdxCache) {  // bound the cache for streams of differently sized images / batches
      CUDA_TRY(cudaStreamSynchronize(s));      // no launch may still be reading an entry
      m->idx_cache.clear();
    }
    DevBuf& buf = m->idx_cache[key];
    CUDA_TRY(buf.reserve(tab.size() * 4));
    CUDA_TRY(cudaMemcpy(buf.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    it = m->idx_cache.find(key);
  }
  *out = it->second.as<int>();
  return CLIPEBC_OK;
}

// sliding_window_predict with a CLIP-ResNet encoder: the windows are convolved independently (the reference slices a window
// first and zero-pads it, so overlapping windows share no activations), chunk by chunk, then folded like the ViT windows
int sliding_resnet(clipebc_model* m, cudaStream_t s, const float* image_dev, int H, int W, int wh, int ww, int sh, int sw,
                   const std::vector<int>& ro, const std::vector<int>& co, float* density_out_dev, float* count_out_dev) {
  const int r = m->cfg.reduction, nr = static_cast<int>(ro.size()), nc = static_cast<int>(co.size()), n_win = nr * nc;
  const int gh = wh / r, gw = ww / r;
  // index table: origins (y, x) per window | row cells | col cells
  std::vector<int> tab(static_cast<size_t>(2) * n_win + nr + nc);
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j) { tab[2 * (i * nc + j)] = ro[i]; tab[2 * (i * nc + j) + 1] = co[j]; }
  for (int i = 0; i < nr; ++i) tab[2 * n_win + i] = ro[i] / r;
  for (int j = 0; j < nc; ++j) tab[2 * n_win + nr + j] = co[j] / r;
  const std::string key = "rsw:" + std::to_string(H) + ":" + std::to_string(W) + ":" + std::to_string(wh) + ":" + std::to_string(ww) +
                          ":" + std::to_string(sh) + ":" + std::to_string(sw);
  const int* d_tab;
  int rc;
  if ((rc = cached_table(m, key, tab, s, &d_tab))) return rc;
  CUDA_TRY(m->ws_preds.reserve(static_cast<size_t>(n_win) * gh * gw * 4));
  float* preds = m->ws_preds.as<float>();
  const int chunk = resnet_default_chunk(m, wh, ww);
  for (int b0 = 0; b0 < n_win; b0 += chunk) {
    const int nw = std::min(chunk, n_win - b0);
    if ((rc = resnet_run_windows(m, s, image_dev, H, W, d_tab + 2 * b0, nw, wh, ww, preds + static_cast<int64_t>(b0) * gh * gw, nullptr)))
      return rc;
  }
  K_TRY(fold_average(s, preds, d_tab + 2 * n_win, d_tab + 2 * n_win + nr, nr, nc, gh, gw, H / r, W / r, density_out_dev, count_out_dev));
  return CLIPEBC_OK;
}

}  // namespace

// =================================================================================================================
extern "C" {

const char* clipebc_last_error(void) { return g_err.c_str(); }
int clipebc_abi_version(void) { return CLIPEBC_ABI_VERSION; }
int64_t clipebc_launch_count(void) { return g_launches.load(); }

int64_t clipebc_config_epoch(void) { return g_config_epoch.load(); }
int clipebc_profile_enabled(void) { return cebc::profiling_on() ? 1 : 0; }
void clipebc_note_replayed_launches(int64_t n) { if (n > 0) g_launches.fetch_add(n); }

int clipebc_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  for (auto& r : g_prof) { if (r.e0) cudaEventDestroy(r.e0); if (r.e1) cudaEventDestroy(r.e1); }
  g_prof.clear();
  return CLIPEBC_OK;
}

int clipebc_profile_dump(char* buf, int cap) {
  if (!buf || cap <= 2) return fail(CLIPEBC_EINVAL, "profile_dump: no buffer");
  CUDA_TRY(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  struct Agg { double ms = 0, flops = 0, bytes = 0; int64_t n = 0; };
  std::map<std::string, Agg> agg;
  std::vector<std::string> order;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
    if (!agg.count(r.tag)) order.push_back(r.tag);
    Agg& a = agg[r.tag];
    a.ms += ms; a.flops += r.flops; a.bytes += r.bytes; a.n += 1;
  }
  std::string out = "{";
  for (size_t i = 0; i < order.size(); ++i) {
    const Agg& a = agg[order[i]];
    char line[256];
    std::snprintf(line, sizeof(line), "%s\"%s\": {\"ms\": %.6f, \"launches\": %lld, \"flops\": %.6e, \"bytes\": %.6e}",
                  i ? ", " : "", order[i].c_str(), a.ms, static_cast<long long>(a.n), a.flops, a.bytes);
    out += line;
  }
  out += "}";
  if (static_cast<int>(out.size()) + 1 > cap) return fail(CLIPEBC_EINVAL, "profile_dump: buffer too small");
  std::memcpy(buf, out.c_str(), out.size() + 1);
  return CLIPEBC_OK;
}

int clipebc_model_create(const clipebc_config* cfg, clipebc_model** out) try {
  if (!cfg || !out) return fail(CLIPEBC_EINVAL, "null argument");
  if (cfg->struct_size != sizeof(clipebc_config))
    return fail(CLIPEBC_EINVAL, "clipebc_config.struct_size is " + std::to_string(cfg->struct_size) + " but this library's "
                "clipebc_config has " + std::to_string(sizeof(clipebc_config)) + " bytes (ABI v" +
                std::to_string(CLIPEBC_ABI_VERSION) + "): the caller was built against another header");
  clipebc_config c = *cfg;
  if (c.encoder != 0 && c.encoder != 1) return fail(CLIPEBC_EINVAL, "encoder must be 0 (ViT) or 1 (CLIP-ResNet)");
  if (c.encoder == 1) {  // the ViT-only fields are ignored: normalise them so that the checks below pass
    c.patch = 16; c.width = 768; c.layers = 12; c.input_size = 224; c.num_vpt = 0; c.deep_vpt = 0;
    if (c.embed_dim == 0) c.embed_dim = 1024;
  }
  if (c.patch == 0) c.patch = 16;
  if (c.width == 0) c.width = 768;
  if (c.layers == 0) c.layers = 12;
  if (c.embed_dim == 0) c.embed_dim = 512;
  if (c.reduction != 8 && c.reduction != 16 && c.reduction != 32) return fail(CLIPEBC_EINVAL, "reduction must be 8, 16 or 32");
  if (c.patch != 14 && c.patch != 16 && c.patch != 32)
    return fail(CLIPEBC_EINVAL, "patch must be 16 (ViT-B/16), 32 (ViT-B/32) or 14 (ViT-L/14)");
  if (c.width != 768 && c.width != 1024) return fail(CLIPEBC_EINVAL, "width must be 768 (ViT-B) or 1024 (ViT-L)");
  if (c.layers < 1 || c.layers > 48) return fail(CLIPEBC_EINVAL, "layers must be in 1..48");
  if (c.encoder == 0 && (c.embed_dim <= 0 || c.embed_dim % 256 != 0 || c.embed_dim > 1024))
    return fail(CLIPEBC_EINVAL, "embed_dim must be 256, 512, 768 or 1024");
  if (c.encoder == 1 && (c.embed_dim <= 0 || c.embed_dim % 8 != 0 || c.embed_dim > 2048))
    return fail(CLIPEBC_EINVAL, "embed_dim must be a positive multiple of 8, at most 2048");
  if (c.input_size <= 0 || c.input_size % c.patch != 0) return fail(CLIPEBC_EINVAL, "input_size must be a multiple of the patch size");
  if (c.num_vpt < 0 || c.num_vpt > 64) return fail(CLIPEBC_EINVAL, "num_vpt out of range");
  if (c.num_bins < 1 || c.num_bins > 32) return fail(CLIPEBC_EINVAL, "num_bins must be in 1..32");
  if (c.operand_fp16 != 0 && c.operand_fp16 != 1) return fail(CLIPEBC_EINVAL, "operand_fp16 must be 0 (bf16) or 1 (fp16)");
  if (c.window_chunk < 0) return fail(CLIPEBC_EINVAL, "window_chunk must not be negative");
  clipebc_model* m = new clipebc_model();
  m->cfg = c;
  m->kp_pad = (3 * c.patch * c.patch + 63) / 64 * 64;
  m->split_precision = c.operand_fp16 == 0;
  m->layer.reset(new LayerPack[c.layers]);
  // the device current at creation owns the handle (-1 on a host without a CUDA device: such a handle can be configured
  // and inspected, every compute entry point then fails with CLIPEBC_ECUDA)
  if (cudaGetDevice(&m->device) != cudaSuccess) { cudaGetLastError(); m->device = -1; }
  *out = m;
  return CLIPEBC_OK;
} catch (const std::exception& e) {
  return fail(CLIPEBC_ESTATE, std::string("C++ exception inside the library: ") + e.what());
} catch (...) {
  return fail(CLIPEBC_ESTATE, "unknown C++ exception inside the library");
}

void clipebc_model_destroy(clipebc_model* m) {
  if (!m) return;
  int cur = -1;
  const bool sw = m->device >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != m->device && cudaSetDevice(m->device) == cudaSuccess;
  delete m;  // frees the device buffers on the device that owns them
  if (sw) cudaSetDevice(cur);
  cudaGetLastError();
}

int clipebc_model_set_tensor(clipebc_model* m, const char* name, const float* data, const int64_t* shape, int ndim) try {
  if (!m || !name || !data || ndim < 0 || (ndim > 0 && !shape)) return fail(CLIPEBC_EINVAL, "null argument");
  int rc;
  if ((rc = check_device(m))) return rc;
  int64_t numel = 1;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] <= 0) return fail(CLIPEBC_EINVAL, std::string("empty tensor '") + name + "'");
    numel *= shape[i];
  }
  RawTensor& t = m->raw[name];
  t.shape.assign(shape, shape + ndim);
  t.numel = numel;
  CUDA_TRY(t.buf.reserve(static_cast<size_t>(numel) * 4));
  CUDA_TRY(cudaMemcpy(t.buf.p, data, static_cast<size_t>(numel) * 4, cudaMemcpyDefault));
  m->packed = false;
  return CLIPEBC_OK;
} catch (const std::exception& e) {
  return fail(CLIPEBC_ESTATE, std::string("C++ exception inside the library: ") + e.what());
} catch (...) {
  return fail(CLIPEBC_ESTATE, "unknown C++ exception inside the library");
}

int clipebc_model_pack(clipebc_model* m, void* stream_) try {
  if (!m) return fail(CLIPEBC_EINVAL, "null model");
  int rc;
  if ((rc = check_device(m))) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const clipebc_config& c = m->cfg;
  const int fp16 = c.operand_fp16 != 0;
  if (c.encoder == 1) {
    if ((rc = resnet_pack(m, s))) return rc;
    if (!check_shape_ok(m, "logit_scale")) return fail(CLIPEBC_ESTATE, "pack: tensor 'logit_scale' missing or not a scalar");
    const int e_pad = m->resnet->e_pad;  // text matrix with zero columns up to the padded embedding (see resnet_pack)
    CUDA_TRY(m->tmat.reserve(static_cast<size_t>(c.num_bins) * e_pad * 4));
    CUDA_TRY(cudaMemsetAsync(m->tmat.p, 0, static_cast<size_t>(c.num_bins) * e_pad * 4, s));
    K_TRY(pack_text(s, raw_ptr(m, "text_features"), raw_ptr(m, "logit_scale"), c.num_bins, c.embed_dim, m->tmat.as<float>(), e_pad));
    CUDA_TRY(cudaStreamSynchronize(s));
    m->packed = true;
    return CLIPEBC_OK;
  }
  const int D = c.width, hidden = 4 * D, E = c.embed_dim, nL = c.layers;
  const int kPatch = c.patch, kp = 3 * kPatch * kPatch;
  const int g0 = c.input_size / kPatch;
  std::string err;
  const int n_vpt_layers = c.deep_vpt ? nL : 1;
  bool ok = true;
  if (c.num_vpt > 0)
    for (int l = 0; l < n_vpt_layers && ok; ++l) ok = check_shape(m, "vpt_" + std::to_string(l), {c.num_vpt, D}, &err);
  if (ok) {  // a scalar: 0-dim as in the reference state_dict, or one element
    auto ls = m->raw.find("logit_scale");
    if (ls == m->raw.end()) { err = "missing tensor 'logit_scale'"; ok = false; }
    else if (ls->second.numel != 1) { err = "tensor 'logit_scale' must hold one element"; ok = false; }
  }
  ok = ok && check_shape(m, "image_encoder.class_embedding", {D}, &err) &&
       check_shape(m, "image_encoder.positional_embedding", {1 + g0 * g0, D}, &err) &&
       check_shape(m, "image_encoder.conv1.weight", {D, 3, kPatch, kPatch}, &err) &&
       check_shape(m, "image_encoder.ln_pre.weight", {D}, &err) && check_shape(m, "image_encoder.ln_pre.bias", {D}, &err) &&
       check_shape(m, "image_encoder.ln_post.weight", {D}, &err) && check_shape(m, "image_encoder.ln_post.bias", {D}, &err);
  for (int l = 0; l < nL && ok; ++l) {
    ok = check_shape(m, blk(l, "attn.in_proj_weight"), {3 * D, D}, &err) &&
         check_shape(m, blk(l, "attn.in_proj_bias"), {3 * D}, &err) &&
         check_shape(m, blk(l, "attn.out_proj.weight"), {D, D}, &err) &&
         check_shape(m, blk(l, "attn.out_proj.bias"), {D}, &err) &&
         check_shape(m, blk(l, "ln_1.weight"), {D}, &err) && check_shape(m, blk(l, "ln_1.bias"), {D}, &err) &&
         check_shape(m, blk(l, "ln_2.weight"), {D}, &err) && check_shape(m, blk(l, "ln_2.bias"), {D}, &err) &&
         check_shape(m, blk(l, "mlp.c_fc.weight"), {hidden, D}, &err) && check_shape(m, blk(l, "mlp.c_fc.bias"), {hidden}, &err) &&
         check_shape(m, blk(l, "mlp.c_proj.weight"), {D, hidden}, &err) && check_shape(m, blk(l, "mlp.c_proj.bias"), {D}, &err);
  }
  for (int k = 1; k <= 2 && ok; ++k) {
    const std::string cv = "image_decoder.0.conv" + std::to_string(k) + ".weight", bn = "image_decoder.0.bn" + std::to_string(k);
    ok = check_shape(m, cv, {D, D, 3, 3}, &err) && check_shape(m, bn + ".weight", {D}, &err) &&
         check_shape(m, bn + ".bias", {D}, &err) && check_shape(m, bn + ".running_mean", {D}, &err) &&
         check_shape(m, bn + ".running_var", {D}, &err);
  }
  ok = ok && check_shape(m, "projection.weight", {E, D, 1, 1}, &err) && check_shape(m, "projection.bias", {E}, &err) &&
       check_shape(m, "text_features", {c.num_bins, E}, &err) && check_shape(m, "anchor_points", {c.num_bins}, &err);
  if (!ok) return fail(CLIPEBC_ESTATE, "pack: " + err);

  CUDA_TRY(m->w_patch.reserve(static_cast<size_t>(D) * 3 * m->kp_pad * 2));
  K_TRY(split_weight_hi_hi_lo(s, raw_ptr(m, "image_encoder.conv1.weight"), D, kp, m->kp_pad, m->w_patch.p, fp16));
  for (int l = 0; l < nL; ++l) {
    LayerPack& L = m->layer[l];
    if ((rc = to_16(s, raw_ptr(m, blk(l, "attn.in_proj_weight")), static_cast<int64_t>(3) * D * D, &L.w_qkv, fp16))) return rc;
    if ((rc = to_16(s, raw_ptr(m, blk(l, "attn.out_proj.weight")), static_cast<int64_t>(D) * D, &L.w_out, fp16))) return rc;
    if ((rc = to_16(s, raw_ptr(m, blk(l, "mlp.c_fc.weight")), static_cast<int64_t>(hidden) * D, &L.w_fc, fp16))) return rc;
    if ((rc = to_16(s, raw_ptr(m, blk(l, "mlp.c_proj.weight")), static_cast<int64_t>(D) * hidden, &L.w_proj, fp16))) return rc;
    L.b_qkv = raw_ptr(m, blk(l, "attn.in_proj_bias"));
    L.b_out = raw_ptr(m, blk(l, "attn.out_proj.bias"));
    L.b_fc = raw_ptr(m, blk(l, "mlp.c_fc.bias"));
    L.b_proj = raw_ptr(m, blk(l, "mlp.c_proj.bias"));
    L.ln1_g = raw_ptr(m, blk(l, "ln_1.weight")); L.ln1_b = raw_ptr(m, blk(l, "ln_1.bias"));
    L.ln2_g = raw_ptr(m, blk(l, "ln_2.weight")); L.ln2_b = raw_ptr(m, blk(l, "ln_2.bias"));
  }
  if (c.deep_vpt && c.num_vpt > 0) {
    // constant prompt K/V of every layer: in_proj(LN1_l(vpt_l)) with the kernels of the live path. These GEMMs read
    // weights that the conversion kernels above have just written, so their weight tiles must not be requested ahead of
    // the programmatic-dependency wait (GemmParams::w_prefetch = 0).
    CUDA_TRY(m->pack_tmp_16.reserve(static_cast<size_t>(c.num_vpt) * D * 2));
    for (int l = 0; l < nL; ++l) {
      LayerPack& L = m->layer[l];
      CUDA_TRY(L.const_kv.reserve(static_cast<size_t>(c.num_vpt) * 3 * D * 2));
      K_TRY(layernorm_rows(s, D, raw_ptr(m, "vpt_" + std::to_string(l)), L.ln1_g, L.ln1_b, m->pack_tmp_16.p, fp16 ? 2 : 1, c.num_vpt, 1, 1, 0));
      GemmParams pk = plain(fp16, 0, c.num_vpt, 3 * D, D, L.const_kv.p, 3 * D, L.b_qkv);
      pk.w_prefetch = 0;
      K_TRY(gemm_dispatch(s, EPI_BIAS_BF16, m->pack_tmp_16.as<__nv_bfloat16>(), c.num_vpt, D, D, L.w_qkv.as<__nv_bfloat16>(), D, pk, 0));
    }
  }
  for (int k = 1; k <= 2; ++k) {
    const std::string cv = "image_decoder.0.conv" + std::to_string(k) + ".weight", bn = "image_decoder.0.bn" + std::to_string(k);
    DevBuf& W = (k == 1) ? m->w_c1 : m->w_c2;
    DevBuf& B = (k == 1) ? m->b_c1 : m->b_c2;
    CUDA_TRY(W.reserve(static_cast<size_t>(D) * 9 * D * 2));
    CUDA_TRY(B.reserve(static_cast<size_t>(D) * 4));
    K_TRY(fold_conv3x3_bn(s, raw_ptr(m, cv), raw_ptr(m, bn + ".weight"), raw_ptr(m, bn + ".bias"), raw_ptr(m, bn + ".running_mean"),
                          raw_ptr(m, bn + ".running_var"), 1e-5f, D, D, W.p, B.as<float>(), fp16));
    if (k == 1) {
      CUDA_TRY(m->w_c1z.reserve(static_cast<size_t>(9) * D * D * 2));
      K_TRY(fold_conv3x3_bn_tapout(s, raw_ptr(m, cv), raw_ptr(m, bn + ".weight"), raw_ptr(m, bn + ".running_var"), 1e-5f, D, D,
                                   m->w_c1z.p, fp16));
      CUDA_TRY(m->zero_bias.reserve(static_cast<size_t>(9) * D * 4));
      CUDA_TRY(cudaMemsetAsync(m->zero_bias.p, 0, static_cast<size_t>(9) * D * 4, s));
    }
  }
  CUDA_TRY(m->w_p3.reserve(static_cast<size_t>(E) * 3 * D * 2));
  K_TRY(split_weight_hi_hi_lo(s, raw_ptr(m, "projection.weight"), E, D, D, m->w_p3.p, fp16));
  CUDA_TRY(m->tmat.reserve(static_cast<size_t>(c.num_bins) * E * 4));
  K_TRY(pack_text(s, raw_ptr(m, "text_features"), raw_ptr(m, "logit_scale"), c.num_bins, E, m->tmat.as<float>()));
  CUDA_TRY(cudaStreamSynchronize(s));
  m->pos_cache.clear();
  m->packed = true;
  return CLIPEBC_OK;
} catch (const std::exception& e) {
  return fail(CLIPEBC_ESTATE, std::string("C++ exception inside the library: ") + e.what());
} catch (...) {
  return fail(CLIPEBC_ESTATE, "unknown C++ exception inside the library");
}

int clipebc_forward_windows(clipebc_model* m, const float* x_dev, int B, int h, int w, float* exp_out_dev,
                            float* logits_out_dev, void* stream_) try {
  if (!m || !x_dev || !exp_out_dev) return fail(CLIPEBC_EINVAL, "null argument");
  if (!m->packed) return fail(CLIPEBC_ESTATE, "model is not packed (call clipebc_model_pack after loading tensors)");
  if (B <= 0) return fail(CLIPEBC_EINVAL, "batch must be positive");
  int rc;
  if ((rc = check_device(m))) return rc;
  if ((rc = check_window_geometry(m, h, w))) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (m->cfg.encoder == 1) {
    const int gh = h / m->cfg.reduction, gw = w / m->cfg.reduction;
    const int chunk = resnet_default_chunk(m, h, w);
    for (int b0 = 0; b0 < B; b0 += chunk) {
      const int nw = std::min(chunk, B - b0);
      float* lo = logits_out_dev ? logits_out_dev + static_cast<int64_t>(b0) * m->cfg.num_bins * gh * gw : nullptr;
      if ((rc = resnet_run_windows(m, s, x_dev + static_cast<int64_t>(b0) * 3 * h * w, h, w, nullptr, nw, h, w,
                                   exp_out_dev + static_cast<int64_t>(b0) * gh * gw, lo)))
        return rc;
    }
    return CLIPEBC_OK;
  }
  const int kPatch = m->cfg.patch;
  const int hp = h / kPatch, wp = w / kPatch, npatch = hp * wp;
  const int gh = h / m->cfg.reduction, gw = w / m->cfg.reduction;
  const float* pos;
  if ((rc = get_pos(m, hp, wp, s, &pos))) return rc;

  const int64_t rows = static_cast<int64_t>(B) * npatch;
  const int fp16 = m->cfg.operand_fp16 != 0;
  if ((rc = reserve_patch_rows(m, rows, s))) return rc;
  K_TRY(patchify(s, x_dev, B, h, w, 0, 0, hp, wp, kPatch, m->kp_pad, m->split_precision, m->ws_patch_rows.p, fp16));
  if ((rc = patch_embed(m, s, rows))) return rc;
  // window b reads patch rows [b * npatch, (b+1) * npatch)
  std::vector<int> base(B);
  for (int b = 0; b < B; ++b) base[b] = b * npatch;
  const int* d_win_base;
  if ((rc = cached_table(m, "fw:" + std::to_string(B) + ":" + std::to_string(npatch), base, s, &d_win_base))) return rc;

  const int chunk = default_chunk(m, hp, wp);
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nw = std::min(chunk, B - b0);
    float* lo = logits_out_dev ? logits_out_dev + static_cast<int64_t>(b0) * m->cfg.num_bins * gh * gw : nullptr;
    if ((rc = run_windows(m, s, d_win_base + b0, wp, nw, hp, wp, pos,
                          exp_out_dev + static_cast<int64_t>(b0) * gh * gw, lo)))
      return rc;
  }
  return CLIPEBC_OK;
} catch (const std::exception& e) {
  return fail(CLIPEBC_ESTATE, std::string("C++ exception inside the library: ") + e.what());
} catch (...) {
  return fail(CLIPEBC_ESTATE, "unknown C++ exception inside the library");
}

int clipebc_window_origins(int H, int W, int wh, int ww, int sh, int sw, int* n_rows, int* n_cols, int* row_origins,
                           int* col_origins) {
  if (!n_rows || !n_cols) return fail(CLIPEBC_EINVAL, "null argument");
  if (wh <= 0 || ww <= 0 || sh <= 0 || sw <= 0) return fail(CLIPEBC_EINVAL, "window size and stride must be positive");
  if (sh > wh || sw > ww) return fail(CLIPEBC_EINVAL, "stride must not exceed the window size");
  if (H < wh || W < ww) return fail(CLIPEBC_EINVAL, "image smaller than the window");
  // int(np.ceil((H - h) / s) + 1)   (utils/eval_utils.py:54-55)
  const int nr = (H - wh + sh - 1) / sh + 1, nc = (W - ww + sw - 1) / sw + 1;
  *n_rows = nr; *n_cols = nc;
  if (row_origins)
    for (int i = 0; i < nr; ++i) { int x0 = i * sh; if (x0 + wh > H) x0 = H - wh; row_origins[i] = x0; }
  if (col_origins)
    for (int j = 0; j < nc; ++j) { int y0 = j * sw; if (y0 + ww > W) y0 = W - ww; col_origins[j] = y0; }
  return CLIPEBC_OK;
}

int clipebc_sliding_window_predict(clipebc_model* m, const float* image_dev, int H, int W, int wh, int ww, int sh,
                                   int sw, float* density_out_dev, float* count_out_dev, void* stream_) try {
  if (!m || !image_dev || !density_out_dev) return fail(CLIPEBC_EINVAL, "null argument");
  if (!m->packed) return fail(CLIPEBC_ESTATE, "model is not packed (call clipebc_model_pack after loading tensors)");
  int rc, nr = 0, nc = 0;
  if ((rc = check_device(m))) return rc;
  if ((rc = clipebc_window_origins(H, W, wh, ww, sh, sw, &nr, &nc, nullptr, nullptr))) return rc;
  if ((rc = check_window_geometry(m, wh, ww))) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const int r = m->cfg.reduction;
  std::vector<int> ro(nr), co(nc);
  clipebc_window_origins(H, W, wh, ww, sh, sw, &nr, &nc, ro.data(), co.data());
  const int n_win = nr * nc;
  if (m->cfg.encoder == 1) return sliding_resnet(m, s, image_dev, H, W, wh, ww, sh, sw, ro, co, density_out_dev, count_out_dev);
  const int kPatch = m->cfg.patch;
  const int hp = wh / kPatch, wp = ww / kPatch, npatch = hp * wp;
  const int gh = wh / r, gw = ww / r;
  const float* pos;
  if ((rc = get_pos(m, hp, wp, s, &pos))) return rc;

  // one patch grid per image, shared by all overlapping windows (the unfold never materialises windows), when every
  // window origin is on it; else a per-window unfold
  bool on_grid = H % kPatch == 0 && W % kPatch == 0;
  for (int v : ro) on_grid = on_grid && (v % kPatch == 0);
  for (int v : co) on_grid = on_grid && (v % kPatch == 0);

  // index tables: [0, n_win) win_base | [n_win, 3 n_win) origins (y, x) | row cells | col cells
  const std::string key = "sw:" + std::to_string(H) + ":" + std::to_string(W) + ":" + std::to_string(wh) + ":" +
                          std::to_string(ww) + ":" + std::to_string(sh) + ":" + std::to_string(sw);
  std::vector<int> tab(static_cast<size_t>(3) * n_win + nr + nc);
  int src_pitch;
  int64_t rows;
  if (on_grid) {
    const int GH = H / kPatch, GW = W / kPatch;
    rows = static_cast<int64_t>(GH) * GW;
    src_pitch = GW;
    for (int i = 0; i < nr; ++i)
      for (int j = 0; j < nc; ++j) tab[i * nc + j] = (ro[i] / kPatch) * GW + co[j] / kPatch;
  } else {
    rows = static_cast<int64_t>(n_win) * npatch;
    src_pitch = wp;
    for (int k = 0; k < n_win; ++k) tab[k] = k * npatch;
  }
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j) {
      tab[n_win + 2 * (i * nc + j)] = ro[i];
      tab[n_win + 2 * (i * nc + j) + 1] = co[j];
    }
  for (int i = 0; i < nr; ++i) tab[3 * n_win + i] = ro[i] / r;       // x_start // reduction (eval_utils.py:90)
  for (int j = 0; j < nc; ++j) tab[3 * n_win + nr + j] = co[j] / r;
  const int* d_base;
  if ((rc = cached_table(m, key, tab, s, &d_base))) return rc;
  const int* d_orig = d_base + n_win;
  const int* d_rc = d_base + 3 * n_win;
  const int* d_cc = d_rc + nr;

  const int fp16 = m->cfg.operand_fp16 != 0;
  if ((rc = reserve_patch_rows(m, rows, s))) return rc;
  if (on_grid) K_TRY(patchify(s, image_dev, 1, H, W, 0, 0, H / kPatch, W / kPatch, kPatch, m->kp_pad, m->split_precision, m->ws_patch_rows.p, fp16));
  else K_TRY(patchify_windows(s, image_dev, H, W, d_orig, n_win, hp, wp, kPatch, m->kp_pad, m->split_precision, m->ws_patch_rows.p, fp16));
  if ((rc = patch_embed(m, s, rows))) return rc;

  CUDA_TRY(m->ws_preds.reserve(static_cast<size_t>(n_win) * gh * gw * 4));
  float* preds = m->ws_preds.as<float>();
  const int chunk = default_chunk(m, hp, wp);
  for (int b0 = 0; b0 < n_win; b0 += chunk) {
    const int nw = std::min(chunk, n_win - b0);
    if ((rc = run_windows(m, s, d_base + b0, src_pitch, nw, hp, wp, pos, preds + static_cast<int64_t>(b0) * gh * gw, nullptr)))
      return rc;
  }
  K_TRY(fold_average(s, preds, d_rc, d_cc, nr, nc, gh, gw, H / r, W / r, density_out_dev, count_out_dev));
  return CLIPEBC_OK;
} catch (const std::exception& e) {
  return fail(CLIPEBC_ESTATE, std::string("C++ exception inside the library: ") + e.what());
} catch (...) {
  return fail(CLIPEBC_ESTATE, "unknown C++ exception inside the library");
}


int clipebc_sliding_window_predict_batch(clipebc_model* m, int n_images, const float* const* images_dev, const int* heights,
                                         const int* widths, int wh, int ww, int sh, int sw, float* const* density_out_dev,
                                         float* counts_out_dev, void* stream_) try {
  if (!m || !images_dev || !heights || !widths || !density_out_dev) return fail(CLIPEBC_EINVAL, "null argument");
  if (n_images <= 0) return fail(CLIPEBC_EINVAL, "batch must contain at least one image");
  if (!m->packed) return fail(CLIPEBC_ESTATE, "model is not packed (call clipebc_model_pack after loading tensors)");
  int rc;
  if ((rc = check_device(m))) return rc;
  if ((rc = check_window_geometry(m, wh, ww))) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (m->cfg.encoder == 1) {  // no activations are shared between windows: image by image
    for (int i = 0; i < n_images; ++i) {
      if (!images_dev[i] || !density_out_dev[i]) return fail(CLIPEBC_EINVAL, "null image or output pointer in the batch");
      if ((rc = clipebc_sliding_window_predict(m, images_dev[i], heights[i], widths[i], wh, ww, sh, sw, density_out_dev[i],
                                               counts_out_dev ? counts_out_dev + i : nullptr, stream_)))
        return rc;
    }
    return CLIPEBC_OK;
  }
  const int r = m->cfg.reduction;
  const int kPatch = m->cfg.patch;
  const int hp = wh / kPatch, wp = ww / kPatch, npatch = hp * wp;
  const int gh = wh / r, gw = ww / r;
  const float* pos;
  if ((rc = get_pos(m, hp, wp, s, &pos))) return rc;

  // geometry of every image: windows (utils/eval_utils.py:54-66), patch rows (shared grid when all origins are on it)
  struct Geom { int H, W, nr, nc, n_win, pitch; bool on_grid; int64_t rows, row_off; int win_off, tab_off; std::vector<int> ro, co; };
  std::vector<Geom> gs(n_images);
  int64_t total_rows = 0;
  int total_win = 0;
  std::string key = "swb:" + std::to_string(wh) + ":" + std::to_string(ww) + ":" + std::to_string(sh) + ":" + std::to_string(sw);
  for (int i = 0; i < n_images; ++i) {
    Geom& g = gs[i];
    if (!images_dev[i] || !density_out_dev[i]) return fail(CLIPEBC_EINVAL, "null image or output pointer in the batch");
    g.H = heights[i]; g.W = widths[i];
    if ((rc = clipebc_window_origins(g.H, g.W, wh, ww, sh, sw, &g.nr, &g.nc, nullptr, nullptr))) return rc;
    g.ro.resize(g.nr); g.co.resize(g.nc);
    clipebc_window_origins(g.H, g.W, wh, ww, sh, sw, &g.nr, &g.nc, g.ro.data(), g.co.data());
    g.n_win = g.nr * g.nc;
    g.on_grid = g.H % kPatch == 0 && g.W % kPatch == 0;
    for (int v : g.ro) g.on_grid = g.on_grid && (v % kPatch == 0);
    for (int v : g.co) g.on_grid = g.on_grid && (v % kPatch == 0);
    g.rows = g.on_grid ? static_cast<int64_t>(g.H / kPatch) * (g.W / kPatch) : static_cast<int64_t>(g.n_win) * npatch;
    g.pitch = g.on_grid ? g.W / kPatch : wp;
    g.row_off = total_rows; g.win_off = total_win;
    total_rows += g.rows; total_win += g.n_win;
    key += ":" + std::to_string(g.H) + "x" + std::to_string(g.W);
  }
  if (total_rows > 0x7fffffff) return fail(CLIPEBC_EINVAL, "batch too large");

  // index tables: win_base | win_pitch | origins (y, x) | per image: row cells, col cells
  size_t tab_len = static_cast<size_t>(4) * total_win;
  for (Geom& g : gs) { g.tab_off = static_cast<int>(tab_len); tab_len += g.nr + g.nc; }
  std::vector<int> tab;
  if (m->idx_cache.find(key) == m->idx_cache.end()) {
    tab.resize(tab_len);
    for (const Geom& g : gs) {
      for (int i = 0; i < g.nr; ++i)
        for (int j = 0; j < g.nc; ++j) {
          const int w = g.win_off + i * g.nc + j;
          tab[w] = static_cast<int>(g.row_off) +
                   (g.on_grid ? (g.ro[i] / kPatch) * g.pitch + g.co[j] / kPatch : (i * g.nc + j) * npatch);
          tab[total_win + w] = g.pitch;
          tab[2 * total_win + 2 * w] = g.ro[i];
          tab[2 * total_win + 2 * w + 1] = g.co[j];
        }
      for (int i = 0; i < g.nr; ++i) tab[g.tab_off + i] = g.ro[i] / r;
      for (int j = 0; j < g.nc; ++j) tab[g.tab_off + g.nr + j] = g.co[j] / r;
    }
  }
  const int* d_tab;
  if ((rc = cached_table(m, key, tab, s, &d_tab))) return rc;
  const int* d_base = d_tab;
  const int* d_pitch = d_tab + total_win;
  const int* d_orig = d_tab + 2 * total_win;

  // patch rows of all images, one patch-embedding GEMM over all of them
  const int fp16 = m->cfg.operand_fp16 != 0;
  if ((rc = reserve_patch_rows(m, total_rows, s))) return rc;
  for (int i = 0; i < n_images; ++i) {
    const Geom& g = gs[i];
    uint16_t* dst = m->ws_patch_rows.as<uint16_t>() + g.row_off * (m->split_precision ? 2 : 1) * m->kp_pad;
    if (g.on_grid) K_TRY(patchify(s, images_dev[i], 1, g.H, g.W, 0, 0, g.H / kPatch, g.W / kPatch, kPatch, m->kp_pad, m->split_precision, dst, fp16));
    else K_TRY(patchify_windows(s, images_dev[i], g.H, g.W, d_orig + 2 * g.win_off, g.n_win, hp, wp, kPatch, m->kp_pad, m->split_precision, dst, fp16));
  }
  if ((rc = patch_embed(m, s, total_rows))) return rc;

  // the windows of all images share the passes of the ViT / decoder / head (chunks may span image boundaries)
  CUDA_TRY(m->ws_preds.reserve(static_cast<size_t>(total_win) * gh * gw * 4));
  float* preds = m->ws_preds.as<float>();
  const int chunk = default_chunk(m, hp, wp);
  for (int b0 = 0; b0 < total_win; b0 += chunk) {
    const int nw = std::min(chunk, total_win - b0);
    if ((rc = run_windows(m, s, d_base + b0, 0, nw, hp, wp, pos, preds + static_cast<int64_t>(b0) * gh * gw, nullptr, d_pitch + b0)))
      return rc;
  }
  for (int i = 0; i < n_images; ++i) {
    const Geom& g = gs[i];
    K_TRY(fold_average(s, preds + static_cast<int64_t>(g.win_off) * gh * gw, d_tab + g.tab_off, d_tab + g.tab_off + g.nr, g.nr,
                       g.nc, gh, gw, g.H / r, g.W / r, density_out_dev[i], counts_out_dev ? counts_out_dev + i : nullptr));
  }
  return CLIPEBC_OK;
} catch (const std::exception& e) {
  return fail(CLIPEBC_ESTATE, std::string("C++ exception inside the library: ") + e.what());
} catch (...) {
  return fail(CLIPEBC_ESTATE, "unknown C++ exception inside the library");
}

// ------------------------------------------------------------------------------------------------ single kernels
int clipebc_f32_to_16(const float* in_dev, void* out, int64_t n, int fp16, void* stream) {
  K_TRY(f32_to_16(static_cast<cudaStream_t>(stream), in_dev, out, n, fp16));
  return CLIPEBC_OK;
}

int clipebc_gemm_bf16(int epi, const void* A, int64_t a_rows, int64_t a_cols, int64_t lda, const void* W, int64_t ldw,
                      int M, int N, int K, int n_seg, const int* seg_row_shift, const int* seg_col_start, void* out,
                      int ldo, const float* bias, const float* resid, int ldr, int mask_hp, int mask_wp, int mask_lead, int block_n,
                      int ab_fp16, int out_fp16, void* stream) {
  if (!A || !W || !out) return fail(CLIPEBC_EINVAL, "null argument");
  if (n_seg < 1 || n_seg > kMaxGemmSegs) return fail(CLIPEBC_EINVAL, "n_seg must be in 1..9");
  if (K <= 0 || K % (64 * n_seg) != 0) return fail(CLIPEBC_EINVAL, "K must be a positive multiple of 64 * n_seg");
  GemmParams p = gemm_params_plain(M, N, K);
  p.n_seg = n_seg;
  p.seg_kblocks = K / 64 / n_seg;
  for (int i = 0; i < n_seg; ++i) {
    p.seg_row_shift[i] = seg_row_shift ? seg_row_shift[i] : 0;
    p.seg_col_start[i] = seg_col_start ? seg_col_start[i] : 0;
  }
  p.out = out; p.ldo = ldo; p.bias = bias; p.resid = resid; p.ldr = ldr; p.mask_hp = mask_hp; p.mask_wp = mask_wp; p.mask_lead = mask_lead != 0;
  p.ab_fp16 = ab_fp16 != 0; p.out_fp16 = out_fp16 != 0;
  p.split_lo = 1;  // epilogue 6 as a test hook always writes hi | lo
  if (epi == EPI_BIAS_RESID16_RELU_MASK_BF16) { p.resid16 = resid; p.resid = nullptr; }  // resid_dev is 16-bit for epilogue 9
  const char* e = gemm_dispatch(static_cast<cudaStream_t>(stream), epi, static_cast<const __nv_bfloat16*>(A), a_rows, a_cols,
                               lda, static_cast<const __nv_bfloat16*>(W), ldw, p, block_n);
  if (e) return fail(std::strncmp(e, "gemm:", 5) == 0 ? CLIPEBC_EINVAL : CLIPEBC_ECUDA, e);
  return CLIPEBC_OK;
}

int clipebc_layernorm(const float* in, const float* g, const float* b, int width, void* out, int out_kind, int64_t n_rows_out,
                      int rows_out_per_group, int rows_in_per_group, int in_row_offset, void* stream) {
  if (!in || !g || !b || !out) return fail(CLIPEBC_EINVAL, "null argument");
  if (out_kind < 0 || out_kind > 2) return fail(CLIPEBC_EINVAL, "layernorm: out_kind must be 0 (f32), 1 (bf16) or 2 (fp16)");
  if (width != 768 && width != 1024) return fail(CLIPEBC_EINVAL, "layernorm: width must be 768 or 1024");
  K_TRY(layernorm_rows(static_cast<cudaStream_t>(stream), width, in, g, b, out, out_kind, n_rows_out, rows_out_per_group,
                       rows_in_per_group, in_row_offset));
  return CLIPEBC_OK;
}

int clipebc_attention(const void* qkv, const void* const_kv, int n_const, int n_win, int t_live, int heads, void* out,
                      int out_fp16, void* stream) {
  if (!qkv || !out) return fail(CLIPEBC_EINVAL, "null argument");
  if (heads < 1 || heads > 32) return fail(CLIPEBC_EINVAL, "attention: heads must be in 1..32");
  K_TRY(attention_dispatch(static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(qkv),
                           static_cast<const __nv_bfloat16*>(const_kv), n_const, n_win, t_live, heads, out, out_fp16 != 0));
  return CLIPEBC_OK;
}

int clipebc_patchify(const float* image, int n_img, int H, int W, int y0, int x0, int gh, int gw, int patch, int kp_pad,
                     int split, void* out, int fp16, void* stream) {
  if (!image || !out) return fail(CLIPEBC_EINVAL, "null argument");
  K_TRY(patchify(static_cast<cudaStream_t>(stream), image, n_img, H, W, y0, x0, gh, gw, patch, kp_pad, split, out, fp16 != 0));
  return CLIPEBC_OK;
}

int clipebc_resample_to_padded(const float* Y, int n_win, int hp, int wp, int gh, int gw, int width, void* Ub, float* Uf,
                               int fp16, void* stream) {
  if (!Y || !Uf) return fail(CLIPEBC_EINVAL, "null argument");
  K_TRY(resample_to_padded(static_cast<cudaStream_t>(stream), width, Y, n_win, hp, wp, gh, gw, Ub, Uf, fp16 != 0));
  return CLIPEBC_OK;
}

int clipebc_fold_average(const float* preds, const int* row_cells_host, const int* col_cells_host, int n_rows, int n_cols,
                         int gh, int gw, int Ho, int Wo, float* density, float* count, void* stream_) {
  if (!preds || !row_cells_host || !col_cells_host || !density) return fail(CLIPEBC_EINVAL, "null argument");
  if (n_rows <= 0 || n_cols <= 0) return fail(CLIPEBC_EINVAL, "fold: no windows");
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  int* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, static_cast<size_t>(n_rows + n_cols) * 4));
  cudaError_t e = cudaMemcpyAsync(d, row_cells_host, static_cast<size_t>(n_rows) * 4, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + n_rows, col_cells_host, static_cast<size_t>(n_cols) * 4, cudaMemcpyHostToDevice, s);
  const char* msg = nullptr;
  if (e == cudaSuccess) msg = fold_average(s, preds, d, d + n_rows, n_rows, n_cols, gh, gw, Ho, Wo, density, count);
  cudaError_t e2 = cudaStreamSynchronize(s);
  cudaFree(d);
  if (e != cudaSuccess) return fail_cuda(e, "fold upload");
  if (msg) return fail(CLIPEBC_ECUDA, msg);
  if (e2 != cudaSuccess) return fail_cuda(e2, "fold sync");
  return CLIPEBC_OK;
}

// ------------------------------------------------------------------------------------------------ pre / post steps
int clipebc_resize_bicubic_aa(const void* in_dev, int in_is_u8, int C, int h, int w, float* tmp_dev, float* out_dev, int H,
                              int W, const float* mean_host, const float* std_host, void* stream) {
  if (!in_dev || !tmp_dev || !out_dev) return fail(CLIPEBC_EINVAL, "null argument");
  if ((mean_host == nullptr) != (std_host == nullptr)) return fail(CLIPEBC_EINVAL, "mean and std must be given together");
  const char* e = resize_bicubic_aa(static_cast<cudaStream_t>(stream), in_dev, in_is_u8 != 0, C, h, w, tmp_dev, out_dev, H, W,
                                    mean_host, std_host);
  if (e) return fail(std::strncmp(e, "resize:", 7) == 0 ? CLIPEBC_EINVAL : CLIPEBC_ECUDA, e);
  return CLIPEBC_OK;
}

int clipebc_pad_normalize(const void* in_dev, int in_is_u8, int C, int h, int w, float* out_dev, int H, int W,
                          const float* mean_host, const float* std_host, void* stream) {
  if (!in_dev || !out_dev) return fail(CLIPEBC_EINVAL, "null argument");
  if ((mean_host == nullptr) != (std_host == nullptr)) return fail(CLIPEBC_EINVAL, "mean and std must be given together");
  const char* e = pad_normalize(static_cast<cudaStream_t>(stream), in_dev, in_is_u8 != 0, C, h, w, out_dev, H, W, mean_host,
                                std_host);
  if (e) return fail(std::strncmp(e, "pad:", 4) == 0 ? CLIPEBC_EINVAL : CLIPEBC_ECUDA, e);
  return CLIPEBC_OK;
}

int clipebc_resize_density_workspace_floats(void) { return resize_density_workspace_floats(); }

int clipebc_resize_density_map(const float* x_dev, int h, int w, int H, int W, float* out_dev, float* workspace_dev,
                               float* sums_out_dev, void* stream) {
  if (!x_dev || !out_dev || !workspace_dev) return fail(CLIPEBC_EINVAL, "null argument");
  const char* e = resize_density_map(static_cast<cudaStream_t>(stream), x_dev, h, w, H, W, out_dev, workspace_dev, sums_out_dev);
  if (e) return fail(std::strncmp(e, "resize_density_map:", 19) == 0 ? CLIPEBC_EINVAL : CLIPEBC_ECUDA, e);
  return CLIPEBC_OK;
}

}  // extern "C"
