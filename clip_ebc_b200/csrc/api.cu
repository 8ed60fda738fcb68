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
#include <mutex>
#include <string>
#include <vector>

#include "../../include/clipebc_b200.h"
#include "kernels.h"

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

// which GEMM kernel serves the hot path: 1 = one CTA per 128-row tile, 2 = CTA pair (cta_group::2, 256-row tiles)
static std::atomic<int> g_gemm_impl{2};

const char* gemm_dispatch(cudaStream_t stream, int epi, const __nv_bfloat16* A, int64_t a_rows, int64_t a_cols, int64_t lda,
                          const __nv_bfloat16* W, int64_t ldw, GemmParams p, int block_n) {
  return g_gemm_impl.load() == 2 ? gemm2_bf16_tn(stream, epi, A, a_rows, a_cols, lda, W, ldw, p, block_n)
                                 : gemm_bf16_tn(stream, epi, A, a_rows, a_cols, lda, W, ldw, p, block_n);
}

// which attention kernel serves the hot path: 1 = mma.sync (legacy tensor path), 2 = tcgen05 one CTA per query tile,
// 3 = tcgen05 persistent warp-specialised (P in TMEM, the softmax groups split the keys of a tile),
// 4 = tcgen05 persistent, two independent chains (a thread owns a query row)
static std::atomic<int> g_attn_impl{4};
// LayerNorm folded into the GEMMs either side of it (run_windows): off by default -- measured on B200 it removes the 24
// LayerNorm launches of a pass (-0.33 ms per 64 windows) but the heavier GEMM epilogues give 0.30 ms back
static std::atomic<int> g_ln_fold{0};
// conv1 of the decoder computed from the coarse patch grid (one GEMM with hp*wp rows per window + a gather kernel) when
// the decoder grid is finer than the patch grid (reduction 8 with ViT-B/16, reductions 8 / 16 with ViT-B/32); 0 = the
// implicit GEMM on the fine grid
static std::atomic<int> g_conv1_coarse{std::getenv("CLIPEBC_CONV1_FINE") ? 0 : 1};  // env: A/B knob for bench runs

const char* attention_dispatch(cudaStream_t stream, const __nv_bfloat16* qkv, const __nv_bfloat16* const_kv, int n_const,
                               int n_win, int t_live, void* out, int out_fp16) {
  const int impl = g_attn_impl.load();
  // windows with more than 256 tokens (e.g. 448 x 448): streamed-K/V kernel, the tcgen05 kernels hold one 256-key tile
  if (t_live + n_const > 256) return attention_h64_long(stream, qkv, const_kv, n_const, n_win, t_live, out, out_fp16);
  if (impl == 4 && n_const % 8 == 0) return attention_h64_pp(stream, qkv, const_kv, n_const, n_win, t_live, out, out_fp16);
  if (impl == 3 && n_const % 8 == 0) return attention_h64_fa(stream, qkv, const_kv, n_const, n_win, t_live, out, out_fp16);
  if (impl == 2 && n_const % 8 == 0) return attention_h64_tc(stream, qkv, const_kv, n_const, n_win, t_live, out, out_fp16);
  return attention_h64(stream, qkv, const_kv, n_const, n_win, t_live, out, out_fp16);
}

namespace {

thread_local std::string g_err;

}  // namespace
namespace { thread_local const cudaAccessPolicyWindow* g_l2_window_tls = nullptr; }
const cudaAccessPolicyWindow* current_l2_window() { return g_l2_window_tls; }
static void set_l2_window(const cudaAccessPolicyWindow* w) { g_l2_window_tls = w; }
namespace {

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int fail_cuda(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return CLIPEBC_ECUDA;
}
#define CUDA_TRY(expr)                                         \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return fail_cuda(_e, #expr);        \
  } while (0)
// kernel launchers return nullptr or a message
#define K_TRY(expr)                                                        \
  do {                                                                     \
    const char* _m = (expr);                                               \
    if (_m != nullptr) return fail(CLIPEBC_ECUDA, std::string(_m));        \
  } while (0)

constexpr int kWidth = 768, kLayers = 12, kHeads = 12, kEmbed = 512, kHidden = 3072;

// Bumped by every clipebc_set_* switch AND by every (re)allocation or release of a device buffer of this library: host
// layers that cache captured CUDA graphs of the library's launches key them on it (a graph holds the kernels chosen and
// the buffer addresses used at capture time -- replaying it after a workspace has moved would write to freed memory).
std::atomic<int64_t> g_config_epoch{0};

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
  DevBuf w_qkv, w_out, w_fc, w_proj;  // bf16
  DevBuf const_kv;                    // bf16 [num_vpt, 2304] (deep VPT)
  // LayerNorm folded into the Linear behind it (rewrite 7): W * diag(gamma) in 16 bits, the column
  // sums of the rounded weights and b + W beta
  DevBuf wf_qkv, wf_fc;               // 16-bit [2304, 768], [3072, 768]
  DevBuf ln_aux;                      // f32: colsum_qkv [2304] | bias_qkv [2304] | colsum_fc [3072] | bias_fc [3072]
  const float *b_qkv, *b_out, *b_fc, *b_proj, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
};

}  // namespace
}  // namespace cebc

using namespace cebc;

struct clipebc_model {
  clipebc_config cfg;
  std::map<std::string, RawTensor> raw;
  bool packed = false;
  // packed
  LayerPack layer[kLayers];
  DevBuf w_patch;           // bf16 [768, 768]
  DevBuf w_c1z;             // 16-bit [9*768, 768]: conv1 with the tap on the output side (coarse-grid form)
  DevBuf zero_bias;         // f32 [9*768] zeros
  DevBuf ws_Y16, ws_Z;      // coarse-grid conv1: 16-bit ln_post rows, per-tap products [n * hp * wp, 9*768]
  DevBuf w_c1, w_c2;        // bf16 [768, 9*768]
  DevBuf b_c1, b_c2;        // f32 [768]
  DevBuf w_p3;              // bf16 [512, 3*768] = hi|hi|lo
  DevBuf tmat;              // f32 [N, 512]
  DevBuf pack_tmp_f32, pack_tmp_bf16;
  std::map<int, DevBuf> pos_cache;  // key hp * 4096 + wp -> f32 [1 + hp*wp, 768]
  // workspace
  DevBuf ws_stats;          // float2 [M, kLnStatSlots]: per-row LayerNorm partials between a residual GEMM and its consumer
  DevBuf ws_patch_rows, ws_patch_embed, ws_X, ws_Xn, ws_QKV, ws_AO, ws_Hid, ws_Y, ws_Ub, ws_Uf, ws_D1, ws_D2, ws_F,
      ws_preds;
  // device-resident index tables (window -> patch-grid row, window origins, fold cell origins), cached per geometry so
  // the steady state has no host->device upload and no host synchronisation
  std::map<std::string, DevBuf> idx_cache;
};

namespace {

const float* raw_ptr(clipebc_model* m, const std::string& name) { return m->raw.at(name).buf.as<float>(); }

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
  const int g0 = m->cfg.input_size / m->cfg.patch;
  if (hp == g0 && wp == g0) { *out = raw_ptr(m, "image_encoder.positional_embedding"); return CLIPEBC_OK; }
  const int key = hp * 4096 + wp;
  auto it = m->pos_cache.find(key);
  if (it != m->pos_cache.end()) { *out = it->second.as<float>(); return CLIPEBC_OK; }
  std::vector<float> src(static_cast<size_t>(1 + g0 * g0) * kWidth);
  CUDA_TRY(cudaMemcpy(src.data(), raw_ptr(m, "image_encoder.positional_embedding"), src.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<float> dst(static_cast<size_t>(1 + hp * wp) * kWidth);
  std::memcpy(dst.data(), src.data(), kWidth * 4);
  const float sy = static_cast<float>(g0) / hp, sx = static_cast<float>(g0) / wp;
  for (int oy = 0; oy < hp; ++oy) {
    const float fy = (oy + 0.5f) * sy - 0.5f;
    const int iy = static_cast<int>(std::floor(fy));
    float wy[4]; cubic_coeffs(fy - iy, wy);
    for (int ox = 0; ox < wp; ++ox) {
      const float fx = (ox + 0.5f) * sx - 0.5f;
      const int ix = static_cast<int>(std::floor(fx));
      float wx[4]; cubic_coeffs(fx - ix, wx);
      float* o = &dst[static_cast<size_t>(1 + oy * wp + ox) * kWidth];
      for (int c = 0; c < kWidth; ++c) o[c] = 0.f;
      for (int a = 0; a < 4; ++a) {
        const int yy = std::min(std::max(iy - 1 + a, 0), g0 - 1);
        for (int b = 0; b < 4; ++b) {
          const int xx = std::min(std::max(ix - 1 + b, 0), g0 - 1);
          const float wgt = wy[a] * wx[b];
          const float* s = &src[static_cast<size_t>(1 + yy * g0 + xx) * kWidth];
          for (int c = 0; c < kWidth; ++c) o[c] += wgt * s[c];
        }
      }
    }
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

// patch embedding as a split-precision GEMM: rows [hi | lo] (2*KP) x W3 = [Whi | Whi | Wlo] (3*KP), segments hi, lo, hi;
// KP = 3 * patch^2 (768 for ViT-B/16, 3072 for ViT-B/32)
GemmParams patch_embed_params(int fp16, int rows, void* out, int kp) {
  GemmParams p = gemm_params_plain(rows, 768, 3 * kp);
  p.n_seg = 3; p.seg_kblocks = kp / 64;
  p.seg_col_start[0] = 0; p.seg_col_start[1] = kp; p.seg_col_start[2] = 0;
  p.out = out; p.ldo = 768; p.bias = nullptr; p.ab_fp16 = fp16; p.out_fp16 = fp16;
  return p;
}

// RAII: persisting-L2 access-policy window over [ptr, ptr + bytes), attached as a LAUNCH attribute to every kernel this
// thread launches while it lives (the caller's stream state is not touched). The device-wide carve-out
// (cudaLimitPersistingL2CacheSize) only ever grows, up to kL2PersistCapMB / the device maximum.
constexpr int kL2PersistCapMB = 60;
struct L2Window {
  cudaAccessPolicyWindow win_ = {};
  bool active_ = false;
  L2Window(const void* ptr, size_t bytes) {
    static const int env_mb = std::getenv("CLIPEBC_L2_PERSIST") ? std::atoi(std::getenv("CLIPEBC_L2_PERSIST")) : -1;
    if (env_mb == 0 || bytes == 0 || cebc::current_l2_window() != nullptr) return;
    static size_t max_persist = [] {
      int dev = 0, v = 0;
      if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, dev) != cudaSuccess) v = 0;
      cudaGetLastError();
      return static_cast<size_t>(v > 0 ? v : 0);
    }();
    size_t carve = std::min(bytes, static_cast<size_t>(env_mb > 0 ? env_mb : kL2PersistCapMB) << 20);
    carve = std::min(carve, max_persist);
    if (carve == 0) return;
    static std::mutex mu;
    static size_t limit_now = 0;
    {
      std::lock_guard<std::mutex> lock(mu);
      if (carve > limit_now) {
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) { cudaGetLastError(); return; }
        limit_now = carve;
      }
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
  const bool deep = c.deep_vpt != 0;
  const int fp16 = c.operand_fp16 != 0;   // 16-bit operand format of every GEMM of the path
  const int ln16 = fp16 ? 2 : 1;          // layernorm768 out_kind
  const int n_prompt_live = deep ? 0 : c.num_vpt;
  const int n_const = deep ? c.num_vpt : 0;
  const int npatch = hp * wp;
  const int T = 1 + n_prompt_live + npatch;
  const int M = nw * T;
  const int gh = hp * c.patch / c.reduction, gw = wp * c.patch / c.reduction;
  const int Hp = gh + 1, Wp = gw + 1;  // shared-border decoder grid (kernels.h: resample_to_padded): 841 rows per r8
  const int Mp = nw * Hp * Wp;         // window instead of 900 on a grid bordered on all four sides

  CUDA_TRY(m->ws_X.reserve(static_cast<size_t>(M) * kWidth * 4));
  CUDA_TRY(m->ws_Xn.reserve(static_cast<size_t>(M) * kWidth * 2));
  CUDA_TRY(m->ws_QKV.reserve(static_cast<size_t>(M) * 3 * kWidth * 2));
  CUDA_TRY(m->ws_AO.reserve(static_cast<size_t>(M) * kWidth * 2));
  CUDA_TRY(m->ws_Hid.reserve(static_cast<size_t>(M) * kHidden * 2));
  CUDA_TRY(m->ws_Y.reserve(static_cast<size_t>(nw) * npatch * kWidth * 4));
  CUDA_TRY(m->ws_Ub.reserve(static_cast<size_t>(Mp) * kWidth * 2));
  CUDA_TRY(m->ws_Uf.reserve(static_cast<size_t>(Mp) * kWidth * 4));
  CUDA_TRY(m->ws_D1.reserve(static_cast<size_t>(Mp) * kWidth * 2));
  CUDA_TRY(m->ws_D2.reserve(static_cast<size_t>(Mp) * 2 * kWidth * 2));
  CUDA_TRY(m->ws_F.reserve(static_cast<size_t>(Mp) * kEmbed * 4));

  float* X = m->ws_X.as<float>();
  // The fp32 residual stream is read / updated four times per block while QKV (58 MB at 64 windows) and Hid (77 MB)
  // stream through L2 once: an access-policy window on the launching stream keeps X in the persisting carve-out of L2
  // for the duration of the pass (measured on B200: +4 % on 2048x1536 images, +0.3..1.3 % at 64 windows;
  // profiles/r01i_l2_persist.txt). CLIPEBC_L2_PERSIST=<MB> overrides the carve-out (0 = off).
  const L2Window l2win(X, static_cast<size_t>(M) * kWidth * 4);
  __nv_bfloat16* Xn = m->ws_Xn.as<__nv_bfloat16>();
  __nv_bfloat16* QKV = m->ws_QKV.as<__nv_bfloat16>();
  __nv_bfloat16* AO = m->ws_AO.as<__nv_bfloat16>();
  __nv_bfloat16* Hid = m->ws_Hid.as<__nv_bfloat16>();

  // Optional (clipebc_set_ln_fold): LayerNorm folded into the GEMMs either side of it (CTA-pair GEMM only): the residual
  // GEMMs leave a 16-bit copy of the new rows and per-row (mean, M2) partials, QKV / c_fc read the raw rows and normalise
  // in their epilogue -- no LayerNorm launch and no second pass over X inside the blocks.
  const bool ln_fold = g_gemm_impl.load() == 2 && g_ln_fold.load() != 0;
  float2* stats = nullptr;
  if (ln_fold) {
    CUDA_TRY(m->ws_stats.reserve(static_cast<size_t>(M) * kLnStatSlots * sizeof(float2)));
    stats = m->ws_stats.as<float2>();
  }

  K_TRY(assemble_tokens(s, m->ws_patch_embed.as<float>(), win_base_dev, src_pitch, win_pitch_dev,
                        raw_ptr(m, "image_encoder.class_embedding"), pos, raw_ptr(m, "image_encoder.ln_pre.weight"),
                        raw_ptr(m, "image_encoder.ln_pre.bias"), deep ? nullptr : raw_ptr(m, "vpt_0"), n_prompt_live, nw,
                        hp, wp, X, ln_fold ? Xn : nullptr, stats, fp16));

  for (int l = 0; l < kLayers && ln_fold; ++l) {
    const LayerPack& L = m->layer[l];
    const float* aux = L.ln_aux.as<float>();
    // ln_1 + in_proj: the rows come from assemble_tokens (whole-row statistics) or from the previous c_proj (8 partials)
    GemmParams pq = plain(fp16, 0, M, 3 * kWidth, kWidth, QKV, 3 * kWidth, aux + 3 * kWidth);
    pq.ln_stats = stats; pq.ln_colsum = aux; pq.ln_parts = (l == 0) ? 1 : kLnStatSlots;
    set_launch_tag("qkv");
    K_TRY(gemm_dispatch(s, EPI_LN_BIAS_BF16, Xn, M, kWidth, kWidth, L.wf_qkv.as<__nv_bfloat16>(), kWidth, pq, 0));
    set_launch_tag(nullptr);
    K_TRY(attention_dispatch(s, QKV, deep ? L.const_kv.as<__nv_bfloat16>() : nullptr, n_const, nw, T, AO, fp16));
    GemmParams po = plain(fp16, fp16, M, kWidth, kWidth, X, kWidth, L.b_out, X, kWidth);
    po.x16_out = Xn; po.stats_out = stats;
    set_launch_tag("out_proj");
    K_TRY(gemm_dispatch(s, EPI_BIAS_RESID_STATS, AO, M, kWidth, kWidth, L.w_out.as<__nv_bfloat16>(), kWidth, po, 192));
    // ln_2 + c_fc + QuickGELU
    GemmParams pf = plain(fp16, fp16, M, kHidden, kWidth, Hid, kHidden, aux + 6 * kWidth + kHidden);
    pf.ln_stats = stats; pf.ln_colsum = aux + 6 * kWidth; pf.ln_parts = kLnStatSlots;
    set_launch_tag("c_fc");
    K_TRY(gemm_dispatch(s, EPI_LN_BIAS_GELU_BF16, Xn, M, kWidth, kWidth, L.wf_fc.as<__nv_bfloat16>(), kWidth, pf, 0));
    GemmParams pj = plain(fp16, fp16, M, kWidth, kHidden, X, kWidth, L.b_proj, X, kWidth);
    pj.x16_out = Xn; pj.stats_out = stats;
    set_launch_tag("c_proj");
    // the last block feeds ln_post (fp32 rows only)
    K_TRY(gemm_dispatch(s, l + 1 < kLayers ? EPI_BIAS_RESID_STATS : EPI_BIAS_RESID_F32, Hid, M, kHidden, kHidden,
                        L.w_proj.as<__nv_bfloat16>(), kHidden, pj, 192));
    set_launch_tag(nullptr);
  }

  for (int l = 0; l < kLayers && !ln_fold; ++l) {
    const LayerPack& L = m->layer[l];
    K_TRY(layernorm768(s, X, L.ln1_g, L.ln1_b, Xn, ln16, M, 1, 1, 0));
    set_launch_tag("qkv");
    K_TRY(gemm_dispatch(s, EPI_BIAS_BF16, Xn, M, kWidth, kWidth, L.w_qkv.as<__nv_bfloat16>(), kWidth,
                       plain(fp16, 0, M, 3 * kWidth, kWidth, QKV, 3 * kWidth, L.b_qkv), 0));
    set_launch_tag(nullptr);
    K_TRY(attention_dispatch(s, QKV, deep ? L.const_kv.as<__nv_bfloat16>() : nullptr, n_const, nw, T, AO, fp16));
    set_launch_tag("out_proj");
    K_TRY(gemm_dispatch(s, EPI_BIAS_RESID_F32, AO, M, kWidth, kWidth, L.w_out.as<__nv_bfloat16>(), kWidth,
                       plain(fp16, fp16, M, kWidth, kWidth, X, kWidth, L.b_out, X, kWidth), 0));
    set_launch_tag(nullptr);
    K_TRY(layernorm768(s, X, L.ln2_g, L.ln2_b, Xn, ln16, M, 1, 1, 0));
    set_launch_tag("c_fc");
    K_TRY(gemm_dispatch(s, EPI_BIAS_GELU_BF16, Xn, M, kWidth, kWidth, L.w_fc.as<__nv_bfloat16>(), kWidth,
                       plain(fp16, fp16, M, kHidden, kWidth, Hid, kHidden, L.b_fc), 0));
    set_launch_tag("c_proj");
    K_TRY(gemm_dispatch(s, EPI_BIAS_RESID_F32, Hid, M, kHidden, kHidden, L.w_proj.as<__nv_bfloat16>(), kHidden,
                       plain(fp16, fp16, M, kWidth, kHidden, X, kWidth, L.b_proj, X, kWidth), 0));
    set_launch_tag(nullptr);
  }

  // ln_post on the patch rows only (cls / prompt rows are dropped, model.py:185-188), fp32 out
  float* Y = m->ws_Y.as<float>();
  // conv1 from the coarse grid (elementwise.cu: conv1_from_coarse_kernel) whenever the decoder grid is at least twice as
  // fine as the patch grid: ln_post then also leaves the 16-bit rows its GEMM reads
  const bool coarse1 = g_conv1_coarse.load() != 0 && g_gemm_impl.load() == 2 && gh >= 2 * hp && gw >= 2 * wp;
  if (coarse1) CUDA_TRY(m->ws_Y16.reserve(static_cast<size_t>(nw) * npatch * kWidth * 2));
  K_TRY(layernorm768(s, X, raw_ptr(m, "image_encoder.ln_post.weight"), raw_ptr(m, "image_encoder.ln_post.bias"), Y, 0,
                     static_cast<int64_t>(nw) * npatch, npatch, T, T - npatch, coarse1 ? m->ws_Y16.p : nullptr, fp16));
  __nv_bfloat16* Ub = m->ws_Ub.as<__nv_bfloat16>();
  float* Uf = m->ws_Uf.as<float>();
  // coarse-grid conv1: the fine-grid map is never materialised -- conv1 reads the per-tap products on the patch grid and
  // conv2's epilogue evaluates the BasicBlock skip (bilinear_up(Y)) on the fly (EPI_BIAS_UPSKIP_RELU_SPLIT)
  if (!coarse1) K_TRY(resample_to_padded(s, Y, nw, hp, wp, gh, gw, Ub, Uf, fp16));

  // decoder BasicBlock as two implicit GEMMs over the zero-bordered grid: 9 taps = 9 row-shifted K-segments
  GemmParams pc = gemm_params_plain(Mp, kWidth, 9 * kWidth);
  pc.ab_fp16 = fp16; pc.out_fp16 = fp16;
  pc.n_seg = 9; pc.seg_kblocks = kWidth / 64;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      pc.seg_row_shift[ky * 3 + kx] = (ky - 1) * Wp + (kx - 1);
      pc.seg_col_start[ky * 3 + kx] = 0;
    }
  __nv_bfloat16* D1 = m->ws_D1.as<__nv_bfloat16>();
  __nv_bfloat16* D2 = m->ws_D2.as<__nv_bfloat16>();
  GemmParams p1 = pc;
  p1.out = D1; p1.ldo = kWidth; p1.bias = m->b_c1.as<float>(); p1.mask_hp = Hp; p1.mask_wp = Wp; p1.mask_lead = 0;
  set_launch_tag("dec_conv1");
  if (coarse1) {
    const int64_t rows_c = static_cast<int64_t>(nw) * npatch;
    CUDA_TRY(m->ws_Z.reserve(static_cast<size_t>(rows_c) * 9 * kWidth * 2));
    K_TRY(gemm_dispatch(s, EPI_BIAS_BF16, m->ws_Y16.as<__nv_bfloat16>(), rows_c, kWidth, kWidth, m->w_c1z.as<__nv_bfloat16>(), kWidth,
                        plain(fp16, fp16, static_cast<int>(rows_c), 9 * kWidth, kWidth, m->ws_Z.p, 9 * kWidth,
                              m->zero_bias.as<float>()), 0));
    set_launch_tag(nullptr);
    K_TRY(conv1_from_coarse(s, m->ws_Z.p, m->b_c1.as<float>(), nw, hp, wp, gh, gw, D1, fp16));
  } else
  K_TRY(gemm_dispatch(s, EPI_BIAS_RELU_MASK_BF16, Ub, Mp, kWidth, kWidth, m->w_c1.as<__nv_bfloat16>(), 9 * kWidth, p1, 0));
  GemmParams p2 = pc;
  p2.out = D2; p2.ldo = 2 * kWidth; p2.bias = m->b_c2.as<float>(); p2.resid = Uf; p2.ldr = kWidth;
  if (coarse1) { p2.resid = Y; p2.mask_hp = Hp; p2.mask_wp = Wp; p2.up_hp = hp; p2.up_wp = wp; }
  set_launch_tag("dec_conv2");
  K_TRY(gemm_dispatch(s, coarse1 ? EPI_BIAS_UPSKIP_RELU_SPLIT : EPI_BIAS_RESID_RELU_SPLIT, D1, Mp, kWidth, kWidth,
                      m->w_c2.as<__nv_bfloat16>(), 9 * kWidth, p2, 0));

  // projection 1x1 in split precision: [hi | lo | hi] x [Whi | Whi | Wlo]  (A segments re-use the hi columns)
  GemmParams pp = gemm_params_plain(Mp, kEmbed, 3 * kWidth);
  pp.ab_fp16 = fp16; pp.out_fp16 = fp16;
  pp.n_seg = 3; pp.seg_kblocks = kWidth / 64;
  pp.seg_col_start[0] = 0; pp.seg_col_start[1] = kWidth; pp.seg_col_start[2] = 0;
  float* F = m->ws_F.as<float>();
  pp.bias = raw_ptr(m, "projection.bias");
  static const bool unfused_head = std::getenv("CLIPEBC_HEAD_UNFUSED") != nullptr;  // A/B knob (profiles/)
  if (g_gemm_impl.load() == 2 && !unfused_head) {
    // projection fused with the head: the 512 projected features of a cell are never written; the GEMM epilogue
    // leaves ||f||^2 and the N bin dot products per half tile (4 partials per row), ebc_head_finish does the rest
    constexpr int kParts = 2 * (kEmbed / 256);
    pp.out = F; pp.ldo = kParts * (1 + c.num_bins);
    pp.head_tmat = m->tmat.as<float>(); pp.head_bins = c.num_bins;
    set_launch_tag("projection+head");
    K_TRY(gemm_dispatch(s, EPI_BIAS_HEAD_PARTIAL, D2, Mp, 2 * kWidth, 2 * kWidth, m->w_p3.as<__nv_bfloat16>(), 3 * kWidth, pp, 256));
    set_launch_tag(nullptr);
    K_TRY(ebc_head_finish(s, F, kParts, raw_ptr(m, "anchor_points"), c.num_bins, nw, gh, gw, exp_out, logits_out));
    return CLIPEBC_OK;
  }
  pp.out = F; pp.ldo = kEmbed;
  set_launch_tag("projection");
  K_TRY(gemm_dispatch(s, EPI_BIAS_F32, D2, Mp, 2 * kWidth, 2 * kWidth, m->w_p3.as<__nv_bfloat16>(), 3 * kWidth, pp, 0));

  set_launch_tag(nullptr);
  K_TRY(ebc_head(s, F, m->tmat.as<float>(), raw_ptr(m, "anchor_points"), c.num_bins, nw, gh, gw, exp_out, logits_out));
  return CLIPEBC_OK;
}

// windows per internal pass: 96 windows of 229 tokens (148 x 128 rows) by default; windows with more tokens get
// proportionally fewer per pass so that the workspaces (4 MB per 229-token window) stay the same size
int default_chunk(const clipebc_model* m, int hp = 14, int wp = 14) {
  if (m->cfg.window_chunk > 0) return m->cfg.window_chunk;
  const int64_t tokens = 1 + m->cfg.num_vpt + static_cast<int64_t>(hp) * wp;
  if (tokens <= 128) return 256;  // ViT-B/32 windows (82 tokens): 256 windows give the GEMMs as many rows as 96 x 197
  if (tokens <= 256) return 96;
  return static_cast<int>(std::max<int64_t>(1, 96 * 229 / tokens));
}

int check_window_geometry(clipebc_model* m, int h, int w) {
  const int kPatch = m->cfg.patch;
  if (h <= 0 || w <= 0 || h % kPatch != 0 || w % kPatch != 0)
    return fail(CLIPEBC_EINVAL, "window height/width must be positive multiples of the patch size");
  if ((h % m->cfg.reduction) != 0 || (w % m->cfg.reduction) != 0)
    return fail(CLIPEBC_EINVAL, "window height/width must be multiples of the reduction");
  const int64_t T = 1 + m->cfg.num_vpt + static_cast<int64_t>(h / kPatch) * (w / kPatch);
  if (T > 16384) return fail(CLIPEBC_EINVAL, "window too large: 1 + num_vpt + patches must be <= 16384 tokens");
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

int clipebc_set_gemm_impl(int impl) {
  g_config_epoch.fetch_add(1);
  if (impl != 1 && impl != 2) return fail(CLIPEBC_EINVAL, "gemm impl must be 1 (single CTA) or 2 (CTA pair)");
  g_gemm_impl.store(impl);
  return CLIPEBC_OK;
}

int clipebc_set_conv1_coarse(int on) {
  g_config_epoch.fetch_add(1);
  g_conv1_coarse.store(on != 0);
  return CLIPEBC_OK;
}

int clipebc_set_ln_fold(int on) {
  g_config_epoch.fetch_add(1);
  g_ln_fold.store(on != 0);
  return CLIPEBC_OK;
}

int clipebc_set_attention_impl(int impl) {
  g_config_epoch.fetch_add(1);
  if (impl < 1 || impl > 4) return fail(CLIPEBC_EINVAL, "attention impl must be 1 (mma.sync), 2, 3 or 4 (tcgen05)");
  g_attn_impl.store(impl);
  return CLIPEBC_OK;
}

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

int clipebc_model_create(const clipebc_config* cfg, clipebc_model** out) {
  if (!cfg || !out) return fail(CLIPEBC_EINVAL, "null argument");
  if (cfg->reduction != 8 && cfg->reduction != 16 && cfg->reduction != 32)
    return fail(CLIPEBC_EINVAL, "reduction must be 8, 16 or 32");
  if (cfg->patch != 0 && cfg->patch != 16 && cfg->patch != 32) return fail(CLIPEBC_EINVAL, "patch must be 16 (ViT-B/16) or 32 (ViT-B/32)");
  const int kPatch = cfg->patch ? cfg->patch : 16;
  if (cfg->input_size <= 0 || cfg->input_size % kPatch != 0) return fail(CLIPEBC_EINVAL, "input_size must be a multiple of the patch size");
  if (cfg->num_vpt < 0 || cfg->num_vpt > 64) return fail(CLIPEBC_EINVAL, "num_vpt out of range");
  if (cfg->num_bins < 1 || cfg->num_bins > 32) return fail(CLIPEBC_EINVAL, "num_bins must be in 1..32");
  if (cfg->operand_fp16 != 0 && cfg->operand_fp16 != 1) return fail(CLIPEBC_EINVAL, "operand_fp16 must be 0 (bf16) or 1 (fp16)");
  clipebc_model* m = new clipebc_model();
  m->cfg = *cfg;
  m->cfg.patch = kPatch;
  *out = m;
  return CLIPEBC_OK;
}

void clipebc_model_destroy(clipebc_model* m) { delete m; }

int clipebc_model_set_tensor(clipebc_model* m, const char* name, const float* data, const int64_t* shape, int ndim) {
  if (!m || !name || !data || ndim < 0 || (ndim > 0 && !shape)) return fail(CLIPEBC_EINVAL, "null argument");
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
}

int clipebc_model_pack(clipebc_model* m, void* stream_) {
  if (!m) return fail(CLIPEBC_EINVAL, "null model");
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const clipebc_config& c = m->cfg;
  const int fp16 = c.operand_fp16 != 0;
  const int kPatch = c.patch, kp = 3 * kPatch * kPatch;
  const int g0 = c.input_size / kPatch;
  std::string err;
  const int n_vpt_layers = c.deep_vpt ? kLayers : 1;
  bool ok = true;
  if (c.num_vpt > 0)
    for (int l = 0; l < n_vpt_layers && ok; ++l) ok = check_shape(m, "vpt_" + std::to_string(l), {c.num_vpt, kWidth}, &err);
  ok = ok && check_shape(m, "logit_scale", {}, &err) &&
       check_shape(m, "image_encoder.class_embedding", {kWidth}, &err) &&
       check_shape(m, "image_encoder.positional_embedding", {1 + g0 * g0, kWidth}, &err) &&
       check_shape(m, "image_encoder.conv1.weight", {kWidth, 3, kPatch, kPatch}, &err) &&
       check_shape(m, "image_encoder.ln_pre.weight", {kWidth}, &err) && check_shape(m, "image_encoder.ln_pre.bias", {kWidth}, &err) &&
       check_shape(m, "image_encoder.ln_post.weight", {kWidth}, &err) && check_shape(m, "image_encoder.ln_post.bias", {kWidth}, &err);
  for (int l = 0; l < kLayers && ok; ++l) {
    ok = check_shape(m, blk(l, "attn.in_proj_weight"), {3 * kWidth, kWidth}, &err) &&
         check_shape(m, blk(l, "attn.in_proj_bias"), {3 * kWidth}, &err) &&
         check_shape(m, blk(l, "attn.out_proj.weight"), {kWidth, kWidth}, &err) &&
         check_shape(m, blk(l, "attn.out_proj.bias"), {kWidth}, &err) &&
         check_shape(m, blk(l, "ln_1.weight"), {kWidth}, &err) && check_shape(m, blk(l, "ln_1.bias"), {kWidth}, &err) &&
         check_shape(m, blk(l, "ln_2.weight"), {kWidth}, &err) && check_shape(m, blk(l, "ln_2.bias"), {kWidth}, &err) &&
         check_shape(m, blk(l, "mlp.c_fc.weight"), {kHidden, kWidth}, &err) && check_shape(m, blk(l, "mlp.c_fc.bias"), {kHidden}, &err) &&
         check_shape(m, blk(l, "mlp.c_proj.weight"), {kWidth, kHidden}, &err) && check_shape(m, blk(l, "mlp.c_proj.bias"), {kWidth}, &err);
  }
  for (int k = 1; k <= 2 && ok; ++k) {
    const std::string cv = "image_decoder.0.conv" + std::to_string(k) + ".weight", bn = "image_decoder.0.bn" + std::to_string(k);
    ok = check_shape(m, cv, {kWidth, kWidth, 3, 3}, &err) && check_shape(m, bn + ".weight", {kWidth}, &err) &&
         check_shape(m, bn + ".bias", {kWidth}, &err) && check_shape(m, bn + ".running_mean", {kWidth}, &err) &&
         check_shape(m, bn + ".running_var", {kWidth}, &err);
  }
  ok = ok && check_shape(m, "projection.weight", {kEmbed, kWidth, 1, 1}, &err) && check_shape(m, "projection.bias", {kEmbed}, &err) &&
       check_shape(m, "text_features", {c.num_bins, kEmbed}, &err) && check_shape(m, "anchor_points", {c.num_bins}, &err);
  if (!ok) return fail(CLIPEBC_ESTATE, "pack: " + err);

  int rc;
  CUDA_TRY(m->w_patch.reserve(static_cast<size_t>(kWidth) * 3 * kp * 2));
  K_TRY(split_weight_hi_hi_lo(s, raw_ptr(m, "image_encoder.conv1.weight"), kWidth, kp, m->w_patch.p, fp16));
  for (int l = 0; l < kLayers; ++l) {
    LayerPack& L = m->layer[l];
    if ((rc = to_16(s, raw_ptr(m, blk(l, "attn.in_proj_weight")), static_cast<int64_t>(3) * kWidth * kWidth, &L.w_qkv, fp16))) return rc;
    if ((rc = to_16(s, raw_ptr(m, blk(l, "attn.out_proj.weight")), static_cast<int64_t>(kWidth) * kWidth, &L.w_out, fp16))) return rc;
    if ((rc = to_16(s, raw_ptr(m, blk(l, "mlp.c_fc.weight")), static_cast<int64_t>(kHidden) * kWidth, &L.w_fc, fp16))) return rc;
    if ((rc = to_16(s, raw_ptr(m, blk(l, "mlp.c_proj.weight")), static_cast<int64_t>(kWidth) * kHidden, &L.w_proj, fp16))) return rc;
    L.b_qkv = raw_ptr(m, blk(l, "attn.in_proj_bias"));
    L.b_out = raw_ptr(m, blk(l, "attn.out_proj.bias"));
    L.b_fc = raw_ptr(m, blk(l, "mlp.c_fc.bias"));
    L.b_proj = raw_ptr(m, blk(l, "mlp.c_proj.bias"));
    L.ln1_g = raw_ptr(m, blk(l, "ln_1.weight")); L.ln1_b = raw_ptr(m, blk(l, "ln_1.bias"));
    L.ln2_g = raw_ptr(m, blk(l, "ln_2.weight")); L.ln2_b = raw_ptr(m, blk(l, "ln_2.bias"));
    CUDA_TRY(L.wf_qkv.reserve(static_cast<size_t>(3) * kWidth * kWidth * 2));
    CUDA_TRY(L.wf_fc.reserve(static_cast<size_t>(kHidden) * kWidth * 2));
    CUDA_TRY(L.ln_aux.reserve(static_cast<size_t>(2) * (3 * kWidth + kHidden) * 4));
    float* aux = L.ln_aux.as<float>();
    K_TRY(fold_ln_linear(s, raw_ptr(m, blk(l, "attn.in_proj_weight")), L.b_qkv, L.ln1_g, L.ln1_b, 3 * kWidth, L.wf_qkv.p, aux,
                         aux + 3 * kWidth, fp16));
    K_TRY(fold_ln_linear(s, raw_ptr(m, blk(l, "mlp.c_fc.weight")), L.b_fc, L.ln2_g, L.ln2_b, kHidden, L.wf_fc.p,
                         aux + 6 * kWidth, aux + 6 * kWidth + kHidden, fp16));
    if (c.deep_vpt && c.num_vpt > 0) {
      // constant prompt K/V of layer l: in_proj(LN1_l(vpt_l)) -- same kernels as the live path
      CUDA_TRY(m->pack_tmp_bf16.reserve(static_cast<size_t>(c.num_vpt) * kWidth * 2));
      CUDA_TRY(L.const_kv.reserve(static_cast<size_t>(c.num_vpt) * 3 * kWidth * 2));
      K_TRY(layernorm768(s, raw_ptr(m, "vpt_" + std::to_string(l)), L.ln1_g, L.ln1_b, m->pack_tmp_bf16.p, fp16 ? 2 : 1, c.num_vpt, 1, 1, 0));
      K_TRY(gemm_dispatch(s, EPI_BIAS_BF16, m->pack_tmp_bf16.as<__nv_bfloat16>(), c.num_vpt, kWidth, kWidth,
                         L.w_qkv.as<__nv_bfloat16>(), kWidth,
                         plain(fp16, 0, c.num_vpt, 3 * kWidth, kWidth, L.const_kv.p, 3 * kWidth, L.b_qkv), 0));
    }
  }
  for (int k = 1; k <= 2; ++k) {
    const std::string cv = "image_decoder.0.conv" + std::to_string(k) + ".weight", bn = "image_decoder.0.bn" + std::to_string(k);
    DevBuf& W = (k == 1) ? m->w_c1 : m->w_c2;
    DevBuf& B = (k == 1) ? m->b_c1 : m->b_c2;
    CUDA_TRY(W.reserve(static_cast<size_t>(kWidth) * 9 * kWidth * 2));
    CUDA_TRY(B.reserve(kWidth * 4));
    K_TRY(fold_conv3x3_bn(s, raw_ptr(m, cv), raw_ptr(m, bn + ".weight"), raw_ptr(m, bn + ".bias"), raw_ptr(m, bn + ".running_mean"),
                          raw_ptr(m, bn + ".running_var"), 1e-5f, kWidth, kWidth, W.p, B.as<float>(), fp16));
    if (k == 1) {
      CUDA_TRY(m->w_c1z.reserve(static_cast<size_t>(9) * kWidth * kWidth * 2));
      K_TRY(fold_conv3x3_bn_tapout(s, raw_ptr(m, cv), raw_ptr(m, bn + ".weight"), raw_ptr(m, bn + ".running_var"), 1e-5f, kWidth,
                                   kWidth, m->w_c1z.p, fp16));
      CUDA_TRY(m->zero_bias.reserve(static_cast<size_t>(9) * kWidth * 4));
      CUDA_TRY(cudaMemsetAsync(m->zero_bias.p, 0, static_cast<size_t>(9) * kWidth * 4, s));
    }
  }
  CUDA_TRY(m->w_p3.reserve(static_cast<size_t>(kEmbed) * 3 * kWidth * 2));
  K_TRY(split_weight_hi_hi_lo(s, raw_ptr(m, "projection.weight"), kEmbed, kWidth, m->w_p3.p, fp16));
  CUDA_TRY(m->tmat.reserve(static_cast<size_t>(c.num_bins) * kEmbed * 4));
  K_TRY(pack_text(s, raw_ptr(m, "text_features"), raw_ptr(m, "logit_scale"), c.num_bins, kEmbed, m->tmat.as<float>()));
  CUDA_TRY(cudaStreamSynchronize(s));
  m->pos_cache.clear();
  m->packed = true;
  return CLIPEBC_OK;
}

int clipebc_forward_windows(clipebc_model* m, const float* x_dev, int B, int h, int w, float* exp_out_dev,
                            float* logits_out_dev, void* stream_) {
  if (!m || !x_dev || !exp_out_dev) return fail(CLIPEBC_EINVAL, "null argument");
  if (!m->packed) return fail(CLIPEBC_ESTATE, "model is not packed (call clipebc_model_pack after loading tensors)");
  if (B <= 0) return fail(CLIPEBC_EINVAL, "batch must be positive");
  int rc;
  if ((rc = check_window_geometry(m, h, w))) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const int kPatch = m->cfg.patch, kp = 3 * kPatch * kPatch;
  const int hp = h / kPatch, wp = w / kPatch, npatch = hp * wp;
  const int gh = h / m->cfg.reduction, gw = w / m->cfg.reduction;
  const float* pos;
  if ((rc = get_pos(m, hp, wp, s, &pos))) return rc;

  const int64_t rows = static_cast<int64_t>(B) * npatch;
  const int fp16 = m->cfg.operand_fp16 != 0;
  CUDA_TRY(m->ws_patch_rows.reserve(static_cast<size_t>(rows) * 2 * kp * 2));
  CUDA_TRY(m->ws_patch_embed.reserve(static_cast<size_t>(rows) * kWidth * 4));
  K_TRY(patchify(s, x_dev, B, h, w, 0, 0, hp, wp, kPatch, m->ws_patch_rows.p, fp16));
  set_launch_tag("patch_embed");
  K_TRY(gemm_dispatch(s, EPI_F32, m->ws_patch_rows.as<__nv_bfloat16>(), rows, 2 * kp, 2 * kp,
                      m->w_patch.as<__nv_bfloat16>(), 3 * kp,
                      patch_embed_params(fp16, static_cast<int>(rows), m->ws_patch_embed.p, kp), 0));
  set_launch_tag(nullptr);
  // window b reads patch rows [b * npatch, (b+1) * npatch)
  const std::string key = "fw:" + std::to_string(B) + ":" + std::to_string(npatch);
  auto cached = m->idx_cache.find(key);
  if (cached == m->idx_cache.end()) {
    std::vector<int> base(B);
    for (int b = 0; b < B; ++b) base[b] = b * npatch;
    DevBuf& buf = m->idx_cache[key];
    CUDA_TRY(buf.reserve(static_cast<size_t>(B) * 4));
    CUDA_TRY(cudaMemcpy(buf.p, base.data(), static_cast<size_t>(B) * 4, cudaMemcpyHostToDevice));
    cached = m->idx_cache.find(key);
  }
  const int* d_win_base = cached->second.as<int>();

  const int chunk = default_chunk(m, hp, wp);
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nw = std::min(chunk, B - b0);
    float* lo = logits_out_dev ? logits_out_dev + static_cast<int64_t>(b0) * m->cfg.num_bins * gh * gw : nullptr;
    if ((rc = run_windows(m, s, d_win_base + b0, wp, nw, hp, wp, pos,
                          exp_out_dev + static_cast<int64_t>(b0) * gh * gw, lo)))
      return rc;
  }
  return CLIPEBC_OK;
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
                                   int sw, float* density_out_dev, float* count_out_dev, void* stream_) {
  if (!m || !image_dev || !density_out_dev) return fail(CLIPEBC_EINVAL, "null argument");
  if (!m->packed) return fail(CLIPEBC_ESTATE, "model is not packed (call clipebc_model_pack after loading tensors)");
  int rc, nr = 0, nc = 0;
  if ((rc = clipebc_window_origins(H, W, wh, ww, sh, sw, &nr, &nc, nullptr, nullptr))) return rc;
  if ((rc = check_window_geometry(m, wh, ww))) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const int r = m->cfg.reduction;
  std::vector<int> ro(nr), co(nc);
  clipebc_window_origins(H, W, wh, ww, sh, sw, &nr, &nc, ro.data(), co.data());
  const int n_win = nr * nc;
  const int kPatch = m->cfg.patch, kp = 3 * kPatch * kPatch;
  const int hp = wh / kPatch, wp = ww / kPatch, npatch = hp * wp;
  const int gh = wh / r, gw = ww / r;
  const float* pos;
  if ((rc = get_pos(m, hp, wp, s, &pos))) return rc;

  bool on_grid = true;
  for (int v : ro) on_grid = on_grid && (v % kPatch == 0);
  for (int v : co) on_grid = on_grid && (v % kPatch == 0);

  // host-side index tables: [0, n_win) win_base | [n_win, 3 n_win) origins (y, x) | row cells | col cells
  const std::string key = "sw:" + std::to_string(H) + ":" + std::to_string(W) + ":" + std::to_string(wh) + ":" +
                          std::to_string(ww) + ":" + std::to_string(sh) + ":" + std::to_string(sw);
  auto cached = m->idx_cache.find(key);
  const bool need_upload = cached == m->idx_cache.end();
  std::vector<int> tab(static_cast<size_t>(3) * n_win + nr + nc);
  int src_pitch;
  int64_t rows;
  if (on_grid) {
    // one patch grid per image, shared by all overlapping windows (the unfold never materialises windows)
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
  if (need_upload) {
    if (m->idx_cache.size() > 64) m->idx_cache.clear();  // bound the cache for streams of differently sized images
    DevBuf& buf = m->idx_cache[key];
    CUDA_TRY(buf.reserve(tab.size() * 4));
    CUDA_TRY(cudaMemcpy(buf.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    cached = m->idx_cache.find(key);
  }
  const int* d_base = cached->second.as<int>();
  const int* d_orig = d_base + n_win;
  const int* d_rc = d_base + 3 * n_win;
  const int* d_cc = d_rc + nr;

  const int fp16 = m->cfg.operand_fp16 != 0;
  CUDA_TRY(m->ws_patch_rows.reserve(static_cast<size_t>(rows) * 2 * kp * 2));
  CUDA_TRY(m->ws_patch_embed.reserve(static_cast<size_t>(rows) * kWidth * 4));
  if (on_grid) K_TRY(patchify(s, image_dev, 1, H, W, 0, 0, H / kPatch, W / kPatch, kPatch, m->ws_patch_rows.p, fp16));
  else K_TRY(patchify_windows(s, image_dev, H, W, d_orig, n_win, hp, wp, kPatch, m->ws_patch_rows.p, fp16));
  set_launch_tag("patch_embed");
  K_TRY(gemm_dispatch(s, EPI_F32, m->ws_patch_rows.as<__nv_bfloat16>(), rows, 2 * kp, 2 * kp,
                      m->w_patch.as<__nv_bfloat16>(), 3 * kp,
                      patch_embed_params(fp16, static_cast<int>(rows), m->ws_patch_embed.p, kp), 0));
  set_launch_tag(nullptr);

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
}


int clipebc_sliding_window_predict_batch(clipebc_model* m, int n_images, const float* const* images_dev, const int* heights,
                                         const int* widths, int wh, int ww, int sh, int sw, float* const* density_out_dev,
                                         float* counts_out_dev, void* stream_) {
  if (!m || !images_dev || !heights || !widths || !density_out_dev) return fail(CLIPEBC_EINVAL, "null argument");
  if (n_images <= 0) return fail(CLIPEBC_EINVAL, "batch must contain at least one image");
  if (!m->packed) return fail(CLIPEBC_ESTATE, "model is not packed (call clipebc_model_pack after loading tensors)");
  int rc;
  if ((rc = check_window_geometry(m, wh, ww))) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const int r = m->cfg.reduction;
  const int kPatch = m->cfg.patch, kp = 3 * kPatch * kPatch;
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
    g.on_grid = true;
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
  auto cached = m->idx_cache.find(key);
  if (cached == m->idx_cache.end()) {
    std::vector<int> tab(tab_len);
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
    if (m->idx_cache.size() > 64) m->idx_cache.clear();
    DevBuf& buf = m->idx_cache[key];
    CUDA_TRY(buf.reserve(tab.size() * 4));
    CUDA_TRY(cudaMemcpy(buf.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    cached = m->idx_cache.find(key);
  }
  const int* d_tab = cached->second.as<int>();
  const int* d_base = d_tab;
  const int* d_pitch = d_tab + total_win;
  const int* d_orig = d_tab + 2 * total_win;

  // patch rows of all images, one patch-embedding GEMM over all of them
  const int fp16 = m->cfg.operand_fp16 != 0;
  CUDA_TRY(m->ws_patch_rows.reserve(static_cast<size_t>(total_rows) * 2 * kp * 2));
  CUDA_TRY(m->ws_patch_embed.reserve(static_cast<size_t>(total_rows) * kWidth * 4));
  for (int i = 0; i < n_images; ++i) {
    const Geom& g = gs[i];
    uint16_t* dst = m->ws_patch_rows.as<uint16_t>() + g.row_off * 2 * kp;
    if (g.on_grid) K_TRY(patchify(s, images_dev[i], 1, g.H, g.W, 0, 0, g.H / kPatch, g.W / kPatch, kPatch, dst, fp16));
    else K_TRY(patchify_windows(s, images_dev[i], g.H, g.W, d_orig + 2 * g.win_off, g.n_win, hp, wp, kPatch, dst, fp16));
  }
  set_launch_tag("patch_embed");
  K_TRY(gemm_dispatch(s, EPI_F32, m->ws_patch_rows.as<__nv_bfloat16>(), total_rows, 2 * kp, 2 * kp,
                      m->w_patch.as<__nv_bfloat16>(), 3 * kp,
                      patch_embed_params(fp16, static_cast<int>(total_rows), m->ws_patch_embed.p, kp), 0));
  set_launch_tag(nullptr);

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
  const char* e = gemm_dispatch(static_cast<cudaStream_t>(stream), epi, static_cast<const __nv_bfloat16*>(A), a_rows, a_cols,
                               lda, static_cast<const __nv_bfloat16*>(W), ldw, p, block_n);
  if (e) return fail(std::strncmp(e, "gemm:", 5) == 0 ? CLIPEBC_EINVAL : CLIPEBC_ECUDA, e);
  return CLIPEBC_OK;
}

int clipebc_gemm_resid_stats(const void* A, int64_t a_rows, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                             float* X, const float* bias, void* x16_out, void* stats_out, int block_n, int ab_fp16,
                             int out_fp16, void* stream) {
  if (!A || !W || !X || !bias || !x16_out || !stats_out) return fail(CLIPEBC_EINVAL, "null argument");
  if (g_gemm_impl.load() != 2) return fail(CLIPEBC_ESTATE, "the statistics epilogue exists in the CTA-pair GEMM only");
  GemmParams p = gemm_params_plain(M, N, K);
  p.out = X; p.ldo = N; p.bias = bias; p.resid = X; p.ldr = N; p.ab_fp16 = ab_fp16; p.out_fp16 = out_fp16;
  p.x16_out = x16_out; p.stats_out = static_cast<float2*>(stats_out);
  K_TRY(gemm2_bf16_tn(static_cast<cudaStream_t>(stream), EPI_BIAS_RESID_STATS, static_cast<const __nv_bfloat16*>(A), a_rows, K,
                      lda, static_cast<const __nv_bfloat16*>(W), ldw, p, block_n));
  return CLIPEBC_OK;
}

int clipebc_gemm_ln(int gelu, const void* A, int64_t a_rows, int64_t lda, const void* Wf, int64_t ldw, int M, int N, int K,
                    void* out, int ldo, const float* bias_f, const void* ln_stats, int ln_parts, const float* ln_colsum,
                    int block_n, int ab_fp16, int out_fp16, void* stream) {
  if (!A || !Wf || !out || !bias_f || !ln_stats || !ln_colsum) return fail(CLIPEBC_EINVAL, "null argument");
  if (g_gemm_impl.load() != 2) return fail(CLIPEBC_ESTATE, "the LayerNorm epilogue exists in the CTA-pair GEMM only");
  GemmParams p = gemm_params_plain(M, N, K);
  p.out = out; p.ldo = ldo; p.bias = bias_f; p.ab_fp16 = ab_fp16; p.out_fp16 = out_fp16;
  p.ln_stats = static_cast<const float2*>(ln_stats); p.ln_colsum = ln_colsum; p.ln_parts = ln_parts;
  K_TRY(gemm2_bf16_tn(static_cast<cudaStream_t>(stream), gelu ? EPI_LN_BIAS_GELU_BF16 : EPI_LN_BIAS_BF16,
                      static_cast<const __nv_bfloat16*>(A), a_rows, K, lda, static_cast<const __nv_bfloat16*>(Wf), ldw, p, block_n));
  return CLIPEBC_OK;
}

int clipebc_rowstats768(const float* in, int64_t n_rows, void* x16_out, void* stats_out, int fp16, void* stream) {
  if (!in || !x16_out || !stats_out) return fail(CLIPEBC_EINVAL, "null argument");
  K_TRY(rowstats768(static_cast<cudaStream_t>(stream), in, n_rows, x16_out, static_cast<float2*>(stats_out), fp16));
  return CLIPEBC_OK;
}

int clipebc_fold_ln_linear(const float* W, const float* b, const float* gamma, const float* beta, int O, void* Wf, float* colsum,
                           float* bias_f, int fp16, void* stream) {
  if (!W || !b || !gamma || !beta || !Wf || !colsum || !bias_f) return fail(CLIPEBC_EINVAL, "null argument");
  K_TRY(fold_ln_linear(static_cast<cudaStream_t>(stream), W, b, gamma, beta, O, Wf, colsum, bias_f, fp16));
  return CLIPEBC_OK;
}

int clipebc_layernorm768(const float* in, const float* g, const float* b, void* out, int out_kind, int64_t n_rows_out,
                         int rows_out_per_group, int rows_in_per_group, int in_row_offset, void* stream) {
  if (out_kind < 0 || out_kind > 2) return fail(CLIPEBC_EINVAL, "layernorm: out_kind must be 0 (f32), 1 (bf16) or 2 (fp16)");
  K_TRY(layernorm768(static_cast<cudaStream_t>(stream), in, g, b, out, out_kind, n_rows_out, rows_out_per_group,
                     rows_in_per_group, in_row_offset));
  return CLIPEBC_OK;
}

int clipebc_attention(const void* qkv, const void* const_kv, int n_const, int n_win, int t_live, void* out, int out_fp16,
                      void* stream) {
  K_TRY(attention_dispatch(static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(qkv),
                           static_cast<const __nv_bfloat16*>(const_kv), n_const, n_win, t_live, out, out_fp16 != 0));
  return CLIPEBC_OK;
}

int clipebc_patchify(const float* image, int n_img, int H, int W, int y0, int x0, int gh, int gw, int patch, void* out,
                     int fp16, void* stream) {
  if (!image || !out) return fail(CLIPEBC_EINVAL, "null argument");
  K_TRY(patchify(static_cast<cudaStream_t>(stream), image, n_img, H, W, y0, x0, gh, gw, patch, out, fp16 != 0));
  return CLIPEBC_OK;
}

int clipebc_patchify16(const float* image, int n_img, int H, int W, int y0, int x0, int gh, int gw, void* out, int fp16,
                       void* stream) {
  K_TRY(patchify(static_cast<cudaStream_t>(stream), image, n_img, H, W, y0, x0, gh, gw, 16, out, fp16 != 0));
  return CLIPEBC_OK;
}

int clipebc_resample_to_padded(const float* Y, int n_win, int hp, int wp, int gh, int gw, void* Ub, float* Uf, int fp16,
                               void* stream) {
  K_TRY(resample_to_padded(static_cast<cudaStream_t>(stream), Y, n_win, hp, wp, gh, gw, Ub, Uf, fp16 != 0));
  return CLIPEBC_OK;
}

int clipebc_ebc_head(const float* F, const float* tmat, const float* anchors, int n_bins, int n_win, int gh, int gw,
                     float* exp_out, float* logits_out, void* stream) {
  K_TRY(ebc_head(static_cast<cudaStream_t>(stream), F, tmat, anchors, n_bins, n_win, gh, gw, exp_out, logits_out));
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
