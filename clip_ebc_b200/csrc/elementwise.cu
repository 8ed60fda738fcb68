// HBM-bound kernels of the CLIP-EBC hot path: LayerNorm, window unfold / patchify, token assembly (cls + pos + ln_pre
// + VPT splice), bilinear resample onto the zero-bordered decoder grid, and the pack-time weight transforms.
// All are one-warp-per-row (D = 768 or 1024 channels = 6 or 8 x 128-bit per lane) or one-thread-per-vector kernels with
// coalesced, vectorised global accesses and warp-shuffle reductions. D is a template parameter: 768 for the ViT-B
// backbones, 1024 for ViT-L/14 (models/clip/model.py:16-24).
#include "common.cuh"
#include "kernels.h"

namespace cebc {

namespace {

template <int D>
struct Row {
  static constexpr int kVec = D / 4 / 32;  // float4 per lane: 6 (768) or 8 (1024)
  float4 v[kVec];
};

template <int D>
__device__ __forceinline__ Row<D> load_row(const float* p, int lane) {
  Row<D> r;
  const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < Row<D>::kVec; ++i) r.v[i] = p4[i * 32 + lane];
  return r;
}
template <int D>
__device__ __forceinline__ Row<D> load_row_ldg(const float* p, int lane) {
  Row<D> r;
  const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < Row<D>::kVec; ++i) r.v[i] = __ldg(p4 + i * 32 + lane);
  return r;
}
template <int D>
__device__ __forceinline__ void add_row(Row<D>& a, const Row<D>& b) {
#pragma unroll
  for (int i = 0; i < Row<D>::kVec; ++i) {
    a.v[i].x += b.v[i].x; a.v[i].y += b.v[i].y; a.v[i].z += b.v[i].z; a.v[i].w += b.v[i].w;
  }
}
// (mean, rstd) of one row held by a warp: nn.LayerNorm(D, eps=1e-5) statistics, two-pass in fp32 (blocks.py:8-14)
template <int D>
__device__ __forceinline__ void row_stats(const Row<D>& r, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < Row<D>::kVec; ++i) s += (r.v[i].x + r.v[i].y) + (r.v[i].z + r.v[i].w);
  mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < Row<D>::kVec; ++i) {
    const float a = r.v[i].x - mean, b = r.v[i].y - mean, c = r.v[i].z - mean, d = r.v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
}
template <int D>
__device__ __forceinline__ void layernorm_row(Row<D>& r, const float* gamma, const float* beta, int lane) {
  float mean, rstd;
  row_stats<D>(r, mean, rstd);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < Row<D>::kVec; ++i) {
    const float4 g = __ldg(g4 + i * 32 + lane), b = __ldg(b4 + i * 32 + lane);
    r.v[i].x = (r.v[i].x - mean) * rstd * g.x + b.x;
    r.v[i].y = (r.v[i].y - mean) * rstd * g.y + b.y;
    r.v[i].z = (r.v[i].z - mean) * rstd * g.z + b.z;
    r.v[i].w = (r.v[i].w - mean) * rstd * g.w + b.w;
  }
}
template <int D>
__device__ __forceinline__ void store_row_f32(float* p, const Row<D>& r, int lane) {
  float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < Row<D>::kVec; ++i) p4[i * 32 + lane] = r.v[i];
}
template <int D>
__device__ __forceinline__ void store_row_16(void* p, const Row<D>& r, int lane, int fp16) {
  uint2* p2 = reinterpret_cast<uint2*>(p);
#pragma unroll
  for (int i = 0; i < Row<D>::kVec; ++i)
    p2[i * 32 + lane] = make_uint2(pack16x2(r.v[i].x, r.v[i].y, fp16), pack16x2(r.v[i].z, r.v[i].w, fp16));
}

// ------------------------------------------------------------------ LayerNorm ----------------------------------
template <int D, bool OUT_16>
__global__ void __launch_bounds__(256, D == 768 ? 4 : 3) layernorm_kernel(const float* __restrict__ in, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, void* __restrict__ out,
                                                         int64_t n_rows_out, int rows_out_per_group,
                                                         int rows_in_per_group, int in_row_offset, int fp16,
                                                         uint16_t* __restrict__ out16_extra) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows_out;
       r += warps_total) {
    const int64_t g = r / rows_out_per_group;
    const int64_t in_row = g * rows_in_per_group + in_row_offset + (r - g * rows_out_per_group);
    Row<D> x = load_row<D>(in + in_row * D, lane);
    layernorm_row<D>(x, gamma, beta, lane);
    if constexpr (OUT_16) {
      store_row_16<D>(static_cast<uint16_t*>(out) + r * D, x, lane, fp16);
    } else {
      store_row_f32<D>(static_cast<float*>(out) + r * D, x, lane);
      if (out16_extra != nullptr) store_row_16<D>(out16_extra + r * D, x, lane, fp16);  // ln_post: + the 16-bit GEMM operand
    }
  }
}

// LayerNorm over contiguous rows with the loads taken off the warps: the warp-per-row kernel above is bound by the latency
// of its row loads (ncu: long-scoreboard stalls, 7 of 16 warps per scheduler active, no pipe above 35 %) and can only
// keep 32 warps x 3 KB in flight per SM. Here every CTA streams blocks of 8 rows (24 KB at D = 768) through a 3-stage
// shared-memory ring with bulk async copies (2 CTAs per SM: 144 KB in flight whatever the warps are doing), warp w
// normalises row w of a block from shared memory, and gamma / beta live in registers for the whole kernel (they are
// loaded before the dependency wait -- they do not depend on the previous kernel).
constexpr int kLnRows = 8, kLnStages = 3, kLnCtasPerSm = 2;
template <int D, int STAGES = kLnStages>
struct LnStream {
  static constexpr int kStageBytes = kLnRows * D * 4;                           // 24 KB / 32 KB
  static constexpr int kSmem = STAGES * kStageBytes + 2 * STAGES * 8 + 128;     // ring + barriers + alignment slack
};

// STAGES x CTAS: ring depth per CTA and resident CTAs per SM (profiles/probes/ln_probe.cu sweeps them)
template <int D, int STAGES = kLnStages, int CTAS = kLnCtasPerSm>
__global__ void __launch_bounds__(256, CTAS) layernorm_stream_kernel(const float* __restrict__ in, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, uint16_t* __restrict__ out,
                                                                int64_t n_rows, int fp16) {
  constexpr int kVec = Row<D>::kVec;
  constexpr int kStageBytes = LnStream<D, STAGES>::kStageBytes;
  constexpr int kLnStages = STAGES;  // shadows the default: this instance's ring depth
  extern __shared__ uint8_t ln_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ln_smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kLnStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kLnStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kLnStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kLnRows); }
    fence_mbar_init();
  }
  // gamma / beta in registers (constants of the model)
  float4 g[kVec], b[kVec];
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    b[i] = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();

  const int64_t n_blocks = (n_rows + kLnRows - 1) / kLnRows;
  const int64_t first = blockIdx.x, step = gridDim.x;
  auto request = [&](int64_t blk, int s) {  // thread 0: rows [blk * 8, ...) -> stage s
    const int64_t r0 = blk * kLnRows;
    const int nr = static_cast<int>(n_rows - r0 < kLnRows ? n_rows - r0 : kLnRows);
    const uint32_t bytes = static_cast<uint32_t>(nr) * D * 4;
    mbar_arrive_expect_tx(&full_bar[s], bytes);
    bulk_load_1d(smem + s * kStageBytes, in + r0 * D, bytes, &full_bar[s]);
  };
  if (threadIdx.x == 0) {
    int s = 0;
    for (int64_t blk = first; blk < n_blocks && s < kLnStages; blk += step, ++s) request(blk, s);
  }
  int s = 0;
  uint32_t phase = 0;
  for (int64_t blk = first; blk < n_blocks; blk += step) {
    mbar_wait(&full_bar[s], phase);
    const int64_t r = blk * kLnRows + warp;
    if (r < n_rows) {
      Row<D> x;
      const float4* p4 = reinterpret_cast<const float4*>(smem + s * kStageBytes + warp * D * 4);
#pragma unroll
      for (int i = 0; i < kVec; ++i) x.v[i] = p4[i * 32 + lane];
      float mean, rstd;
      row_stats<D>(x, mean, rstd);
#pragma unroll
      for (int i = 0; i < kVec; ++i) {
        x.v[i].x = (x.v[i].x - mean) * rstd * g[i].x + b[i].x;
        x.v[i].y = (x.v[i].y - mean) * rstd * g[i].y + b[i].y;
        x.v[i].z = (x.v[i].z - mean) * rstd * g[i].z + b[i].z;
        x.v[i].w = (x.v[i].w - mean) * rstd * g[i].w + b[i].w;
      }
      store_row_16<D>(out + r * D, x, lane, fp16);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);  // this warp is done with the stage (its row is in registers / stored)
    // thread 0 refills the stage with the block kLnStages ahead once all eight warps have released it
    if (threadIdx.x == 0) {
      const int64_t nxt = blk + static_cast<int64_t>(kLnStages) * step;
      if (nxt < n_blocks) {
        mbar_wait(&empty_bar[s], phase);
        request(nxt, s);
      }
    }
    if (++s == kLnStages) { s = 0; phase ^= 1; }
  }
}

// ------------------------------------------------------------------ patchify -----------------------------------
// One thread per VEC horizontally adjacent pixels of a patch row (VEC = 4 for patch 16 / 32, 2 for the 14-pixel patches of
// ViT-L/14); consecutive threads walk along an image row (coalesced reads), each writes VEC 16-bit values into its patch
// row. P = patch size, KP = columns per half (>= 3 * P * P; the columns beyond 3 * P * P are zeroed once by the caller).
// split = 1 (bf16 operands): rows are written as [hi(KP) | lo(KP)] with hi = round16(x), lo = round16(x - hi): the patch-embed
// GEMM multiplies [hi | lo | hi] x [Whi | Whi | Wlo] and so sees the fp32 pixels to ~2^-17 instead of 2^-9. split = 0 (fp16
// operands, 11-bit mantissa): hi only -- one rounding of the pixels, like every later activation of the path.
template <int VEC>
__device__ __forceinline__ void load_store_hi_lo(const float* src, uint16_t* dst, int kp, int split, int fp16) {
  if constexpr (VEC == 4) {
    float4 v;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) v = *reinterpret_cast<const float4*>(src);
    else v = make_float4(src[0], src[1], src[2], src[3]);
    const float hx = round16(v.x, fp16), hy = round16(v.y, fp16), hz = round16(v.z, fp16), hw = round16(v.w, fp16);
    *reinterpret_cast<uint2*>(dst) = make_uint2(pack16x2(hx, hy, fp16), pack16x2(hz, hw, fp16));
    if (split)
      *reinterpret_cast<uint2*>(dst + kp) = make_uint2(pack16x2(v.x - hx, v.y - hy, fp16), pack16x2(v.z - hz, v.w - hw, fp16));
  } else {
    const float vx = src[0], vy = src[1];
    const float hx = round16(vx, fp16), hy = round16(vy, fp16);
    *reinterpret_cast<uint32_t*>(dst) = pack16x2(hx, hy, fp16);
    if (split) *reinterpret_cast<uint32_t*>(dst + kp) = pack16x2(vx - hx, vy - hy, fp16);
  }
}

// origins_yx == nullptr: n_units images [3, H, W], patch grid gh x gw starting at pixel (y0, x0) of every image;
// else: n_units windows of ONE image, window u has its (0, 0) patch at pixel origins_yx[2u], origins_yx[2u + 1].
// Grid: x = 64-vector segments of a pixel row of the grid, y = groups of 4 pixel rows, z = (unit, channel) planes -- the
// position of a thread needs no run-time division (P is a template constant; the first version decoded a flat 64-bit index with
// four 64-bit divisions: 231 instructions per thread, issue-bound at a quarter of the HBM rate).
template <int VEC, int P>
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ image, int n_units, int H, int W, int y0,
                                                       int x0, const int* __restrict__ origins_yx, int gh, int gw, int kp,
                                                       int split, uint16_t* __restrict__ out, int fp16) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int vpp = P / VEC;  // vectors per patch row
  const int qx = blockIdx.x * 64 + threadIdx.x;
  const int yy = blockIdx.y * 4 + threadIdx.y;
  if (qx >= gw * vpp || yy >= gh * P) return;
  const int gx = qx / vpp, pxv = qx - gx * vpp;
  const int gy = yy / P, py = yy - gy * P;
  const int row_stride = (1 + split) * kp;
  for (int z = blockIdx.z; z < 3 * n_units; z += gridDim.z) {
    const int unit = z / 3, c = z - 3 * unit;
    const float* src;
    if (origins_yx != nullptr)
      src = image + (static_cast<int64_t>(c) * H + (origins_yx[2 * unit] + yy)) * W + origins_yx[2 * unit + 1] + qx * VEC;
    else
      src = image + (static_cast<int64_t>(z) * H + (y0 + yy)) * W + x0 + qx * VEC;
    const int64_t patch = (static_cast<int64_t>(unit) * gh + gy) * gw + gx;
    load_store_hi_lo<VEC>(src, out + patch * row_stride + c * P * P + py * P + pxv * VEC, kp, split, fp16);
  }
}

// ------------------------------------------------------------------ token assembly -----------------------------
template <int D>
__global__ void __launch_bounds__(256) assemble_tokens_kernel(const float* __restrict__ patch_embed,
                                                              const int* __restrict__ win_base, int src_pitch,
                                                              const int* __restrict__ win_pitch,
                                                              const float* __restrict__ class_emb,
                                                              const float* __restrict__ pos,
                                                              const float* __restrict__ ln_g,
                                                              const float* __restrict__ ln_b,
                                                              const float* __restrict__ vpt0, int n_prompt, int n_win,
                                                              int hp, int wp, float* __restrict__ X) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int t_live = 1 + n_prompt + hp * wp;
  const int64_t n_rows = static_cast<int64_t>(n_win) * t_live;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows;
       r += warps_total) {
    const int win = static_cast<int>(r / t_live);
    const int t = static_cast<int>(r - static_cast<int64_t>(win) * t_live);
    Row<D> x;
    if (t == 0) {
      x = load_row_ldg<D>(class_emb, lane);
      add_row<D>(x, load_row_ldg<D>(pos, lane));
      layernorm_row<D>(x, ln_g, ln_b, lane);
    } else if (t <= n_prompt) {
      x = load_row_ldg<D>(vpt0 + static_cast<int64_t>(t - 1) * D, lane);  // prompts join after ln_pre (model.py:161-168)
    } else {
      const int pidx = t - 1 - n_prompt;
      const int py = pidx / wp, px = pidx - py * wp;
      const int pitch = win_pitch != nullptr ? win_pitch[win] : src_pitch;  // windows of different images in one pass
      const int64_t src = static_cast<int64_t>(win_base[win]) + static_cast<int64_t>(py) * pitch + px;
      x = load_row<D>(patch_embed + src * D, lane);
      add_row<D>(x, load_row_ldg<D>(pos + static_cast<int64_t>(1 + pidx) * D, lane));
      layernorm_row<D>(x, ln_g, ln_b, lane);
    }
    store_row_f32<D>(X + r * D, x, lane);
  }
}

// ------------------------------------------------------------------ resample -----------------------------------
template <int D>
__global__ void __launch_bounds__(256) resample_to_padded_kernel(const float* __restrict__ Y, int n_win, int hp, int wp,
                                                                 int gh, int gw, uint16_t* __restrict__ U_16,
                                                                 float* __restrict__ U_f32, int fp16) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int Hp = gh + 1, Wp = gw + 1;  // shared-border grid: one trailing zero column per line, one zero row per window
  const int64_t n_rows = static_cast<int64_t>(n_win) * Hp * Wp;
  const float inv_sy = static_cast<float>(hp) / static_cast<float>(gh);
  const float inv_sx = static_cast<float>(wp) / static_cast<float>(gw);
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows;
       r += warps_total) {
    const int win = static_cast<int>(r / (Hp * Wp));
    const int q = static_cast<int>(r - static_cast<int64_t>(win) * Hp * Wp);
    const int py = q / Wp, px = q - py * Wp;
    Row<D> o;
    if (py == gh || px == gw) {
#pragma unroll
      for (int i = 0; i < Row<D>::kVec; ++i) o.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (gh == hp && gw == wp) {
      o = load_row<D>(Y + (static_cast<int64_t>(win) * hp * wp + static_cast<int64_t>(py) * wp + px) * D, lane);
    } else {
      int y0, y1, x0, x1;
      float ly, lx;
      bilinear_src(py, inv_sy, hp, y0, y1, ly);
      bilinear_src(px, inv_sx, wp, x0, x1, lx);
      const float* base = Y + static_cast<int64_t>(win) * hp * wp * D;
      const Row<D> a = load_row<D>(base + (static_cast<int64_t>(y0) * wp + x0) * D, lane);
      const Row<D> b = load_row<D>(base + (static_cast<int64_t>(y0) * wp + x1) * D, lane);
      const Row<D> c = load_row<D>(base + (static_cast<int64_t>(y1) * wp + x0) * D, lane);
      const Row<D> d = load_row<D>(base + (static_cast<int64_t>(y1) * wp + x1) * D, lane);
      const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#pragma unroll
      for (int i = 0; i < Row<D>::kVec; ++i) {
        o.v[i].x = w00 * a.v[i].x + w01 * b.v[i].x + w10 * c.v[i].x + w11 * d.v[i].x;
        o.v[i].y = w00 * a.v[i].y + w01 * b.v[i].y + w10 * c.v[i].y + w11 * d.v[i].y;
        o.v[i].z = w00 * a.v[i].z + w01 * b.v[i].z + w10 * c.v[i].z + w11 * d.v[i].z;
        o.v[i].w = w00 * a.v[i].w + w01 * b.v[i].w + w10 * c.v[i].w + w11 * d.v[i].w;
      }
    }
    if (U_16 != nullptr) store_row_16<D>(U_16 + r * D, o, lane, fp16);
    store_row_f32<D>(U_f32 + r * D, o, lane);
  }
}

// ------------------------------------------------------------------ conv1 of the decoder from the coarse grid ------
// conv3x3(bilinear_up(Y)) = sum over the 9 taps of bilinear_up(W_tap Y) shifted by the tap: the channel contraction
// commutes with the (linear, per-channel) resampling, so it runs ONCE per tap on the coarse patch grid -- Z[src, tap, o] =
// sum_c W'[o, c, tap] Y[src, c], a GEMM with hp*wp rows per window instead of (g+1)^2 (196 instead of 841 at reduction 8:
// 2.08 instead of 8.9 GFLOP per window) -- and this kernel gathers, per output cell and tap, the four bilinear neighbours
// of the shifted position: 36 multiply-adds per output value instead of 6912. Zero padding of the conv applies to the
// FINE grid (taps that leave the window are skipped), the bilinear source clamps at the window edge, exactly as
// F.interpolate followed by conv2d(padding=1) (models/clip/model.py:195-197, models/utils.py:290-296).
// Thread = (cell of the shared-border grid, 8 output channels); ReLU and the zero border rows as in the GEMM epilogue it
// replaces. Z: 16-bit [n_win * hp * wp, 9 * 768] (column = tap * 768 + o), bias f32 [768], D1: 16-bit [.., 768].
template <bool FP16>
__device__ __forceinline__ void fma8_16(float2 (&acc)[4], const uint4& z, float w) {
  const uint32_t u[4] = {z.x, z.y, z.z, z.w};
  const float2 w2 = make_float2(w, w);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 f;
    if constexpr (FP16) f = __half22float2(*reinterpret_cast<const __half2*>(&u[k]));
    else f = make_float2(__uint_as_float(u[k] << 16), __uint_as_float(u[k] & 0xFFFF0000u));
    acc[k] = __ffma2_rn(w2, f, acc[k]);  // packed fp32 FMA (sm_100): two multiply-adds per instruction
  }
}

// Tiling: one CTA = (window, band of `bh` fine rows, tile of `bw` fine columns, slice of CS output channels); the coarse
// positions the tile can touch are staged once in shared memory for all 9 taps (cp.async), thread item = (cell, 8
// channels of the slice). Two shapes are used: the whole window with 16-channel slices when its patch grid fits
// (hp * wp * 9 * 32 B <= 56 KB: no position is fetched twice), else bands of 4 x 32 cells with 32-channel slices (<= 5 x
// 19 coarse positions, 54 KB). What bounds it: 36 16-byte shared-memory reads per 8 outputs (no reuse between cells:
// neighbouring cells read the same positions under different taps) -- 3 GB per 64 windows.
constexpr int kC1Smem = 5 * 19 * 9 * 32 * 2;  // 54720 B: the band shape; the whole-window shape is checked against it too

template <int D, int CS, bool FP16>
__global__ void __launch_bounds__(256) conv1_from_coarse_kernel(const uint16_t* __restrict__ Z, const float* __restrict__ bias,
                                                                int n_win, int hp, int wp, int gh, int gw, int bh, int bw,
                                                                int n_bands, int n_ctiles, uint16_t* __restrict__ D1) {
  constexpr int fp16 = FP16 ? 1 : 0;
  extern __shared__ __align__(16) uint8_t c1_smem[];
  pdl_launch_dependents();
  pdl_wait();
  constexpr int kSlices = D / CS;
  constexpr int kG = CS / 8;          // 8-channel groups per cell
  constexpr int kRowB = CS * 2;       // bytes per (position, tap)
  constexpr int kPosB = 9 * kRowB;    // bytes per position
  const int Hp = gh + 1, Wp = gw + 1;
  const float inv_sy = static_cast<float>(hp) / static_cast<float>(gh), inv_sx = static_cast<float>(wp) / static_cast<float>(gw);
  int b = blockIdx.x;
  const int slice = b % kSlices; b /= kSlices;
  const int ct = b % n_ctiles; b /= n_ctiles;
  const int band = b % n_bands;
  const int win = b / n_bands;
  const int y_lo = band * bh, x_lo = ct * bw;
  const int y_hi = min(y_lo + bh, Hp), x_hi = min(x_lo + bw, Wp);  // cells [y_lo, y_hi) x [x_lo, x_hi) incl. border
  // coarse range touched by the fine positions [y_lo - 1, y_hi] x [x_lo - 1, x_hi] clipped to the window
  int i_lo, i_hi, j_lo, j_hi, t0, t1;
  float lam;
  bilinear_src(max(y_lo - 1, 0), inv_sy, hp, i_lo, t1, lam);
  bilinear_src(min(y_hi, gh - 1), inv_sy, hp, t0, i_hi, lam);
  bilinear_src(max(x_lo - 1, 0), inv_sx, wp, j_lo, t1, lam);
  bilinear_src(min(x_hi, gw - 1), inv_sx, wp, t0, j_hi, lam);
  const int nr = i_hi - i_lo + 1, nc = j_hi - j_lo + 1;
  // ---- stage Z[i_lo..i_hi, j_lo..j_hi, all taps, slice] ----
  const int64_t ldz = 9 * D;
  const uint16_t* zw = Z + static_cast<int64_t>(win) * hp * wp * ldz + slice * CS;
  const int n_chunks = nr * nc * 9 * kG;
  for (int i = threadIdx.x; i < n_chunks; i += blockDim.x) {
    const int ch = i % kG, pt = i / kG;
    const int tap = pt % 9, pos = pt / 9;
    const int rr = pos / nc, cc = pos - rr * nc;
    // asynchronous copies: all of a thread's chunks are in flight at once (a load -> store loop waits for every load)
    cp_async_16(smem_u32(c1_smem + pt * kRowB + (ch << 4)),
                zw + (static_cast<int64_t>(i_lo + rr) * wp + (j_lo + cc)) * ldz + tap * D + ch * 8, true);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  // ---- compute ----
  const int tw = x_hi - x_lo, th = y_hi - y_lo;
  const int n_items = th * tw * kG;
  for (int it = threadIdx.x; it < n_items; it += blockDim.x) {
    const int g = it % kG, cell = it / kG;
    const int cy = cell / tw, cx = cell - cy * tw;
    const int py = y_lo + cy, px = x_lo + cx;
    uint4 outv = make_uint4(0u, 0u, 0u, 0u);
    if (py < gh && px < gw) {
      float2 acc[4];
      const float* bp = bias + slice * CS + g * 8;
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bp)), b1 = __ldg(reinterpret_cast<const float4*>(bp) + 1);
      acc[0] = make_float2(b0.x, b0.y); acc[1] = make_float2(b0.z, b0.w); acc[2] = make_float2(b1.x, b1.y); acc[3] = make_float2(b1.z, b1.w);
      // column sources of the three horizontal taps, once per cell
      int xo0[3], xo1[3];
      float lxs[3];
      bool xok[3];
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = px + dx;
        xok[dx + 1] = xx >= 0 && xx < gw;
        int x0, x1;
        bilinear_src(xok[dx + 1] ? xx : px, inv_sx, wp, x0, x1, lxs[dx + 1]);
        xo0[dx + 1] = (x0 - j_lo) * kPosB; xo1[dx + 1] = (x1 - j_lo) * kPosB;
      }
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = py + dy;
        if (yy < 0 || yy >= gh) continue;
        int y0, y1;
        float ly;
        bilinear_src(yy, inv_sy, hp, y0, y1, ly);
        const int yo0 = (y0 - i_lo) * nc * kPosB, yo1 = (y1 - i_lo) * nc * kPosB;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          if (!xok[dx + 1]) continue;
          const float lx = lxs[dx + 1];
          const uint8_t* zt = c1_smem + ((dy + 1) * 3 + (dx + 1)) * kRowB + (g << 4);
          const uint4 a = *reinterpret_cast<const uint4*>(zt + yo0 + xo0[dx + 1]);
          const uint4 bq = *reinterpret_cast<const uint4*>(zt + yo0 + xo1[dx + 1]);
          const uint4 c = *reinterpret_cast<const uint4*>(zt + yo1 + xo0[dx + 1]);
          const uint4 d = *reinterpret_cast<const uint4*>(zt + yo1 + xo1[dx + 1]);
          fma8_16<FP16>(acc, a, (1.f - ly) * (1.f - lx));
          fma8_16<FP16>(acc, bq, (1.f - ly) * lx);
          fma8_16<FP16>(acc, c, ly * (1.f - lx));
          fma8_16<FP16>(acc, d, ly * lx);
        }
      }
      outv = make_uint4(pack16x2(fmaxf(acc[0].x, 0.f), fmaxf(acc[0].y, 0.f), fp16), pack16x2(fmaxf(acc[1].x, 0.f), fmaxf(acc[1].y, 0.f), fp16),
                        pack16x2(fmaxf(acc[2].x, 0.f), fmaxf(acc[2].y, 0.f), fp16), pack16x2(fmaxf(acc[3].x, 0.f), fmaxf(acc[3].y, 0.f), fp16));
    }
    const int64_t r = (static_cast<int64_t>(win) * Hp + py) * Wp + px;
    *reinterpret_cast<uint4*>(D1 + r * D + slice * CS + g * 8) = outv;
  }
}

// The same gather for a decoder grid exactly twice the patch grid in height (reduction 8 on patch 16, reduction 16 on patch
// 32: the benchmark shapes), with the horizontal and the vertical half of the bilinear weights applied one after the other:
//   H[i, px, dy] = sum_dx sum_{j in {x0, x1}(px + dx)} wx * Z[i, j, tap(dy, dx)]                       (6 terms, coarse row i)
//   out[py, px]  = bias + sum_dy ( wa * H[ia(py + dy), px, dy] + wb * H[ib(py + dy), px, dy] )        (6 terms)
// At ratio 2 the fine row f reads the coarse rows floor((f - 1) / 2) and the one after it (clamped to the window) with
// weights (1/4, 3/4) for even f and (3/4, 1/4) for odd f, so the pair of output rows (2k, 2k + 1) needs H of the coarse rows
// k - 1, k, k + 1 only and every H row serves four output rows: a thread = (output column, 8 channels) walks down its column
// with the three H rows in registers -- 9 shared-memory terms + 6 register terms per output value instead of 36
// shared-memory terms, and the column geometry is computed once per thread instead of once per cell (the cell-per-thread
// kernel above is issue-bound on exactly that: 151 M instructions per 64 windows). `rs` threads share a column (each a
// range of coarse rows, recomputing one H row either side). Rows are staged per band of `nb` coarse rows (+ one either side).
template <int D, bool FP16>
__global__ void __launch_bounds__(256, 2) conv1_from_coarse_x2_kernel(const uint16_t* __restrict__ Z, const float* __restrict__ bias,
                                                                      int n_win, int hp, int wp, int gw, int nb, int n_bands,
                                                                      int rs, uint16_t* __restrict__ D1) {
  constexpr int fp16 = FP16 ? 1 : 0;
  constexpr int CS = 16, kSlices = D / CS, kG = CS / 8, kRowB = CS * 2, kPosB = 9 * kRowB;
  extern __shared__ __align__(16) uint8_t c1_smem[];
  pdl_launch_dependents();
  const int gh = 2 * hp, Hp = gh + 1, Wp = gw + 1;
  int b = blockIdx.x;
  const int slice = b % kSlices; b /= kSlices;
  const int band = b % n_bands;
  const int win = b / n_bands;
  const int kb_lo = band * nb, kb_hi = min(kb_lo + nb, hp);  // output row pairs [kb_lo, kb_hi)
  const int i_lo = max(kb_lo - 1, 0), i_hi = min(kb_hi, hp - 1);
  // ---- this thread's column: sources and weights of the three horizontal taps ----
  const int per_seg = Wp * kG;
  const int seg = threadIdx.x / per_seg, rem = threadIdx.x - seg * per_seg;
  const int px = rem / kG, g = rem - px * kG;
  const float inv_sx = static_cast<float>(wp) / static_cast<float>(gw);
  int xo0[3], xo1[3];
  float wl[3], wr[3];
#pragma unroll
  for (int dx = -1; dx <= 1; ++dx) {
    const int xx = px + dx;
    const bool ok = seg < rs && px < gw && xx >= 0 && xx < gw;  // zero padding applies to the FINE grid; the border column is zero
    int x0, x1;
    float lx;
    bilinear_src(ok ? xx : 0, inv_sx, wp, x0, x1, lx);
    xo0[dx + 1] = x0 * kPosB + (dx + 1) * kRowB + (g << 4);
    xo1[dx + 1] = x1 * kPosB + (dx + 1) * kRowB + (g << 4);
    wl[dx + 1] = ok ? 1.f - lx : 0.f;
    wr[dx + 1] = ok ? lx : 0.f;
  }
  float2 bia[4];
  {
    const float* bp = bias + slice * CS + g * 8;
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bp)), b1 = __ldg(reinterpret_cast<const float4*>(bp) + 1);
    bia[0] = make_float2(b0.x, b0.y); bia[1] = make_float2(b0.z, b0.w); bia[2] = make_float2(b1.x, b1.y); bia[3] = make_float2(b1.z, b1.w);
  }
  pdl_wait();
  // ---- stage Z[i_lo..i_hi, all columns, all taps, slice] ----
  {
    const int64_t ldz = 9 * D;
    const uint16_t* zw = Z + (static_cast<int64_t>(win) * hp + i_lo) * wp * ldz + slice * CS;
    const int n_chunks = (i_hi - i_lo + 1) * wp * 9 * kG;
    for (int i = threadIdx.x; i < n_chunks; i += blockDim.x) {
      const int ch = i % kG, pt = i / kG;
      const int tap = pt % 9, pos = pt / 9;
      cp_async_16(smem_u32(c1_smem + pt * kRowB + (ch << 4)), zw + pos * ldz + tap * D + ch * 8, true);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
  }
  const int nk = kb_hi - kb_lo, per = (nk + rs - 1) / rs;
  const int k_lo = kb_lo + seg * per, k_hi = min(k_lo + per, kb_hi);
  if (seg >= rs || px >= Wp || k_lo >= k_hi) return;
  // H of one coarse row for the vertical taps [d_lo, d_hi]
  auto hrow = [&](int i, float2 (&h)[3][4], int d_lo, int d_hi) {
    const uint8_t* zr = c1_smem + (i - i_lo) * wp * kPosB;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      if (dy < d_lo || dy > d_hi) continue;
#pragma unroll
      for (int q = 0; q < 4; ++q) h[dy][q] = make_float2(0.f, 0.f);
      const uint8_t* zt = zr + dy * 3 * kRowB;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const uint4 a = *reinterpret_cast<const uint4*>(zt + xo0[dx]);
        const uint4 c = *reinterpret_cast<const uint4*>(zt + xo1[dx]);
        fma8_16<FP16>(h[dy], a, wl[dx]);
        fma8_16<FP16>(h[dy], c, wr[dx]);
      }
    }
  };
  auto axpy2 = [](float2 (&acc)[4], const float2 (&u)[4], float wu, const float2 (&v)[4], float wv) {
    const float2 a2 = make_float2(wu, wu), b2 = make_float2(wv, wv);
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = __ffma2_rn(b2, v[q], __ffma2_rn(a2, u[q], acc[q]));
  };
  const bool border_col = px >= gw;
  auto store_row = [&](int py, const float2 (&acc)[4]) {
    uint4 outv = make_uint4(0u, 0u, 0u, 0u);
    if (!border_col)
      outv = make_uint4(pack16x2(fmaxf(acc[0].x, 0.f), fmaxf(acc[0].y, 0.f), fp16), pack16x2(fmaxf(acc[1].x, 0.f), fmaxf(acc[1].y, 0.f), fp16),
                        pack16x2(fmaxf(acc[2].x, 0.f), fmaxf(acc[2].y, 0.f), fp16), pack16x2(fmaxf(acc[3].x, 0.f), fmaxf(acc[3].y, 0.f), fp16));
    const int64_t r = (static_cast<int64_t>(win) * Hp + py) * Wp + px;
    *reinterpret_cast<uint4*>(D1 + r * D + slice * CS + g * 8) = outv;
  };
  float2 Hm[3][4], Hc[3][4], Hn[3][4];
  hrow(k_lo, Hc, 0, 2);
  if (k_lo > 0) hrow(k_lo - 1, Hm, 0, 1);
  else {
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int q = 0; q < 4; ++q) Hm[dy][q] = Hc[dy][q];
  }
  for (int k = k_lo; k < k_hi; ++k) {
    if (k + 1 < hp) hrow(k + 1, Hn, k + 1 < k_hi ? 0 : 1, 2);
    else {
#pragma unroll
      for (int dy = 1; dy < 3; ++dy)
#pragma unroll
        for (int q = 0; q < 4; ++q) Hn[dy][q] = Hc[dy][q];
    }
    float2 acc[4];
    // fine row 2k: taps read the fine rows 2k - 1 (k-1: 3/4, k: 1/4), 2k (k-1: 1/4, k: 3/4), 2k + 1 (k: 3/4, k+1: 1/4)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = bia[q];
    if (k > 0) axpy2(acc, Hm[0], 0.75f, Hc[0], 0.25f);
    axpy2(acc, Hm[1], 0.25f, Hc[1], 0.75f);
    axpy2(acc, Hc[2], 0.75f, Hn[2], 0.25f);
    store_row(2 * k, acc);
    // fine row 2k + 1: 2k (k-1: 1/4, k: 3/4), 2k + 1 (k: 3/4, k+1: 1/4), 2k + 2 (k: 1/4, k+1: 3/4)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = bia[q];
    axpy2(acc, Hm[0], 0.25f, Hc[0], 0.75f);
    axpy2(acc, Hc[1], 0.75f, Hn[1], 0.25f);
    if (k + 1 < hp) axpy2(acc, Hc[2], 0.25f, Hn[2], 0.75f);
    store_row(2 * k + 1, acc);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int q = 0; q < 4; ++q) { Hm[dy][q] = Hc[dy][q]; Hc[dy][q] = Hn[dy][q]; }
  }
  if (k_hi == hp) {  // the zero border row below the window
    const int64_t r = (static_cast<int64_t>(win) * Hp + gh) * Wp + px;
    *reinterpret_cast<uint4*>(D1 + r * D + slice * CS + g * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// ------------------------------------------------------------------ pack-time ----------------------------------
__global__ void f32_to_16_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, int64_t n, int fp16) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = cvt16(in[i], fp16);
}

// BN(eval) folds exactly into the bias-free conv: W' = W * g / sqrt(var + eps), b' = beta - mean * g / sqrt(var + eps)
// (reference: models/utils.py:290-303 with nn.BatchNorm2d in eval mode).
__global__ void fold_conv3x3_bn_kernel(const float* __restrict__ W, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ mean,
                                       const float* __restrict__ var, float eps, int O, int I,
                                       uint16_t* __restrict__ Wp, float* __restrict__ bias, int fp16) {
  const int64_t total = static_cast<int64_t>(O) * I * 9;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // idx enumerates the destination [O][tap][I]
    const int i = static_cast<int>(idx % I);
    const int tap = static_cast<int>((idx / I) % 9);
    const int o = static_cast<int>(idx / (static_cast<int64_t>(I) * 9));
    const float s = gamma[o] / sqrtf(var[o] + eps);
    Wp[idx] = cvt16(W[(static_cast<int64_t>(o) * I + i) * 9 + tap] * s, fp16);
    if (i == 0 && tap == 0) bias[o] = beta[o] - mean[o] * s;
  }
}

// The same fold with the TAP on the output side: Wz[(tap * O + o), i] = W[o, i, tap] * g / sqrt(var + eps) -- the weight
// of the coarse-grid form of conv1 (conv1_from_coarse_kernel): one [O, I] matrix per tap, stacked along N
__global__ void fold_conv3x3_bn_tapout_kernel(const float* __restrict__ W, const float* __restrict__ gamma,
                                              const float* __restrict__ var, float eps, int O, int I,
                                              uint16_t* __restrict__ Wz, int fp16) {
  const int64_t total = static_cast<int64_t>(O) * I * 9;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // idx enumerates the destination [tap][O][I]
    const int i = static_cast<int>(idx % I);
    const int o = static_cast<int>((idx / I) % O);
    const int tap = static_cast<int>(idx / (static_cast<int64_t>(I) * O));
    Wz[idx] = cvt16(W[(static_cast<int64_t>(o) * I + i) * 9 + tap] * (gamma[o] / sqrtf(var[o] + eps)), fp16);
  }
}

__global__ void split_weight_kernel(const float* __restrict__ W, int O, int I, int Ip, uint16_t* __restrict__ out, int fp16) {
  const int64_t total = static_cast<int64_t>(O) * Ip;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx % Ip);
    const int64_t o = idx / Ip;
    const float w = i < I ? W[o * I + i] : 0.0f;  // columns I .. Ip - 1: zero padding up to the GEMM's K granularity
    const float hf = round16(w, fp16);
    const uint16_t hi = cvt16(hf, fp16);
    const uint16_t lo = cvt16(w - hf, fp16);
    uint16_t* row = out + o * 3 * Ip;
    row[i] = hi;
    row[Ip + i] = hi;
    row[2 * Ip + i] = lo;
  }
}

// F.normalize(text, p=2, dim=-1) (eps 1e-12) scaled by exp(logit_scale)  (model.py:204,207-208); one warp per bin
__global__ void pack_text_kernel(const float* __restrict__ text, const float* __restrict__ logit_scale, int n, int d,
                                 float* __restrict__ tmat, int ld_out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) { const float v = text[static_cast<int64_t>(row) * d + i]; s += v * v; }
  s = warp_sum(s);
  const float scale = expf(logit_scale[0]) / fmaxf(sqrtf(s), 1e-12f);
  for (int i = lane; i < d; i += 32) tmat[static_cast<int64_t>(row) * ld_out + i] = text[static_cast<int64_t>(row) * d + i] * scale;
}

inline int grid_for(int64_t work_items, int per_block, int max_blocks) {
  int64_t b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return static_cast<int>(b);
}

inline const char* last_err() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace

const char* layernorm_rows(cudaStream_t stream, int width, const float* in, const float* gamma, const float* beta, void* out,
                           int out_kind, int64_t n_rows_out, int rows_out_per_group, int rows_in_per_group,
                           int in_row_offset, void* out16_extra, int fp16_extra) {
  if (n_rows_out <= 0) return nullptr;
  if (width != 768 && width != 1024) return "layernorm: width must be 768 or 1024";
  if (rows_out_per_group <= 0 || rows_in_per_group <= 0) return "layernorm: bad row map";
  LaunchScope scope(stream, "layernorm", 0.0, static_cast<double>(n_rows_out) * width * (4.0 + (out_kind ? 2.0 : 4.0)));
  // contiguous rows to a 16-bit output (the LayerNorms inside the blocks): streaming kernel
  if (out_kind != 0 && rows_out_per_group == rows_in_per_group && in_row_offset == 0 && n_rows_out >= 64 &&
      (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    static unsigned long long attr_done_768 = 0, attr_done_1024 = 0;
    cudaError_t ea = width == 768 ? ensure_dyn_smem(layernorm_stream_kernel<768>, LnStream<768>::kSmem, &attr_done_768)
                                  : ensure_dyn_smem(layernorm_stream_kernel<1024>, LnStream<1024>::kSmem, &attr_done_1024);
    if (ea != cudaSuccess) return cudaGetErrorString(ea);
    const int blocks_s = grid_for(n_rows_out, kLnRows, device_num_sms() * 2);
    cudaError_t es = width == 768
        ? launch_pdl(layernorm_stream_kernel<768>, dim3(blocks_s), dim3(256), LnStream<768>::kSmem, stream, 1, in, gamma, beta,
                     static_cast<uint16_t*>(out), n_rows_out, out_kind == 2)
        : launch_pdl(layernorm_stream_kernel<1024>, dim3(blocks_s), dim3(256), LnStream<1024>::kSmem, stream, 1, in, gamma, beta,
                     static_cast<uint16_t*>(out), n_rows_out, out_kind == 2);
    return es != cudaSuccess ? cudaGetErrorString(es) : last_err();
  }
  const int blocks = grid_for(n_rows_out, 8, device_num_sms() * 8);
  auto go = [&](auto kern, int fp16, uint16_t* extra) {
    return launch_pdl(kern, dim3(blocks), dim3(256), 0, stream, 1, in, gamma, beta, out, n_rows_out, rows_out_per_group,
                      rows_in_per_group, in_row_offset, fp16, extra);
  };
  cudaError_t e;
  if (out_kind) e = width == 768 ? go(layernorm_kernel<768, true>, out_kind == 2, static_cast<uint16_t*>(nullptr))
                                 : go(layernorm_kernel<1024, true>, out_kind == 2, static_cast<uint16_t*>(nullptr));
  else e = width == 768 ? go(layernorm_kernel<768, false>, fp16_extra, static_cast<uint16_t*>(out16_extra))
                        : go(layernorm_kernel<1024, false>, fp16_extra, static_cast<uint16_t*>(out16_extra));
  return e != cudaSuccess ? cudaGetErrorString(e) : last_err();
}

namespace {
const char* patchify_launch(cudaStream_t stream, const float* image, int n_units, int H, int W, int y0, int x0,
                            const int* origins, int gh, int gw, int patch, int kp_pad, int split, void* out, int fp16) {
  split = split != 0;
  if (patch != 14 && patch != 16 && patch != 32) return "patchify: patch size must be 14, 16 or 32";
  if (kp_pad < 3 * patch * patch || kp_pad % 8 != 0) return "patchify: bad padded patch row length";
  const int vec = patch % 4 == 0 ? 4 : 2;
  const int64_t total = static_cast<int64_t>(n_units) * 3 * gh * patch * gw * (patch / vec);
  LaunchScope scope(stream, "patchify", 0.0, static_cast<double>(total) * vec * (4.0 + 2.0 * (1 + split)));
  const int64_t planes = static_cast<int64_t>(n_units) * 3;
  const dim3 grid((gw * (patch / vec) + 63) / 64, (gh * patch + 3) / 4, static_cast<unsigned>(planes < 65535 ? planes : 65535));
  if (grid.y > 65535u) return "patchify: patch grid too tall";
  auto go = [&](auto kern) {
    return launch_pdl(kern, grid, dim3(64, 4), 0, stream, 1, image, n_units, H, W, y0, x0, origins, gh, gw, kp_pad, split,
                      static_cast<uint16_t*>(out), fp16);
  };
  cudaError_t e = patch == 16 ? go(patchify_kernel<4, 16>) : patch == 32 ? go(patchify_kernel<4, 32>) : go(patchify_kernel<2, 14>);
  return e != cudaSuccess ? cudaGetErrorString(e) : last_err();
}
}  // namespace

const char* patchify(cudaStream_t stream, const float* image, int n_img, int H, int W, int y0, int x0, int gh, int gw,
                     int patch, int kp_pad, int split, void* out, int fp16) {
  if (n_img <= 0 || gh <= 0 || gw <= 0) return "patchify: empty grid";
  if (y0 < 0 || x0 < 0 || y0 + gh * patch > H || x0 + gw * patch > W) return "patchify: grid exceeds image";
  return patchify_launch(stream, image, n_img, H, W, y0, x0, nullptr, gh, gw, patch, kp_pad, split, out, fp16);
}

const char* patchify_windows(cudaStream_t stream, const float* image, int H, int W, const int* origins_yx_dev,
                             int n_win, int hp, int wp, int patch, int kp_pad, int split, void* out, int fp16) {
  if (n_win <= 0) return "patchify: no windows";
  if (origins_yx_dev == nullptr) return "patchify: window origins missing";
  return patchify_launch(stream, image, n_win, H, W, 0, 0, origins_yx_dev, hp, wp, patch, kp_pad, split, out, fp16);
}

const char* assemble_tokens(cudaStream_t stream, int width, const float* patch_embed, const int* win_base_dev, int src_pitch,
                            const int* win_pitch_dev, const float* class_emb, const float* pos, const float* ln_g, const float* ln_b,
                            const float* vpt0, int n_prompt, int n_win, int hp, int wp, float* X) {
  if (n_win <= 0) return "assemble_tokens: no windows";
  if (width != 768 && width != 1024) return "assemble_tokens: width must be 768 or 1024";
  if (n_prompt > 0 && vpt0 == nullptr) return "assemble_tokens: prompts missing";
  const int64_t rows = static_cast<int64_t>(n_win) * (1 + n_prompt + hp * wp);
  LaunchScope scope(stream, "assemble_tokens", 0.0, static_cast<double>(rows) * width * 8.0);
  const dim3 grid(grid_for(rows, 8, device_num_sms() * 8));
  cudaError_t e = width == 768
      ? launch_pdl(assemble_tokens_kernel<768>, grid, dim3(256), 0, stream, 1, patch_embed, win_base_dev, src_pitch, win_pitch_dev,
                   class_emb, pos, ln_g, ln_b, vpt0, n_prompt, n_win, hp, wp, X)
      : launch_pdl(assemble_tokens_kernel<1024>, grid, dim3(256), 0, stream, 1, patch_embed, win_base_dev, src_pitch, win_pitch_dev,
                   class_emb, pos, ln_g, ln_b, vpt0, n_prompt, n_win, hp, wp, X);
  return e != cudaSuccess ? cudaGetErrorString(e) : last_err();
}

const char* resample_to_padded(cudaStream_t stream, int width, const float* Y, int n_win, int hp, int wp, int gh, int gw,
                               void* U_16, float* U_f32, int fp16) {
  if (n_win <= 0) return "resample: no windows";
  if (width != 768 && width != 1024) return "resample: width must be 768 or 1024";
  const int64_t rows = static_cast<int64_t>(n_win) * (gh + 1) * (gw + 1);
  LaunchScope scope(stream, "resample", 0.0, static_cast<double>(n_win) * hp * wp * width * 4.0 + static_cast<double>(rows) * width * 6.0);
  const dim3 grid(grid_for(rows, 8, device_num_sms() * 8));
  cudaError_t e = width == 768 ? launch_pdl(resample_to_padded_kernel<768>, grid, dim3(256), 0, stream, 1, Y, n_win, hp, wp, gh, gw,
                                            static_cast<uint16_t*>(U_16), U_f32, fp16)
                               : launch_pdl(resample_to_padded_kernel<1024>, grid, dim3(256), 0, stream, 1, Y, n_win, hp, wp, gh, gw,
                                            static_cast<uint16_t*>(U_16), U_f32, fp16);
  return e != cudaSuccess ? cudaGetErrorString(e) : last_err();
}

namespace {
template <int D>
const char* conv1_from_coarse_t(cudaStream_t stream, const void* Z, const float* bias, int n_win, int hp, int wp, int gh, int gw,
                                void* D1, int fp16) {
  static unsigned long long attr_done = 0;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr_done & (1ull << (dev & 63)))) {
      unsigned long long m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0, m5 = 0;
      cudaError_t ea = ensure_dyn_smem(conv1_from_coarse_kernel<D, 16, true>, 57344, &m0);
      if (ea == cudaSuccess) ea = ensure_dyn_smem(conv1_from_coarse_kernel<D, 16, false>, 57344, &m1);
      if (ea == cudaSuccess) ea = ensure_dyn_smem(conv1_from_coarse_kernel<D, 32, true>, kC1Smem, &m2);
      if (ea == cudaSuccess) ea = ensure_dyn_smem(conv1_from_coarse_kernel<D, 32, false>, kC1Smem, &m3);
      if (ea == cudaSuccess) ea = ensure_dyn_smem(conv1_from_coarse_x2_kernel<D, true>, 57344, &m4);
      if (ea == cudaSuccess) ea = ensure_dyn_smem(conv1_from_coarse_x2_kernel<D, false>, 57344, &m5);
      if (ea != cudaSuccess) return cudaGetErrorString(ea);
      attr_done |= 1ull << (dev & 63);
    }
  }
  LaunchScope scope(stream, "conv1_interp", 0.0,
                    static_cast<double>(n_win) * hp * wp * 9 * D * 2.0 + static_cast<double>(n_win) * (gh + 1) * (gw + 1) * D * 2.0);
  cudaError_t e;
  const int whole = hp * wp * 9 * 16 * 2;  // the window's patch grid, all taps, 16 channels
  constexpr int c1_rs = 2;  // threads per column (profiles/r02/conv1_x2_ab.txt: 1 -> 106 us, 2 -> 95, 3 / 4 -> 116 at 64 windows)
  const int row_b = wp * 9 * 16 * 2;                  // one coarse row, all taps, 16 channels
  const int nb_fit = 57344 / row_b - 2;               // band height that fits with a row either side
  if (gh == 2 * hp && gw >= 2 * wp && (whole <= 57344 || nb_fit >= 2) && (gw + 1) * 2 <= 256) {
    const int nb = whole <= 57344 ? hp : nb_fit;
    const int n_bands = (hp + nb - 1) / nb;
    int rs = std::max(1, std::min(c1_rs, std::min(nb, 256 / ((gw + 1) * 2))));
    const int threads = ((gw + 1) * 2 * rs + 31) / 32 * 32;
    const int smem = std::min(hp, nb + 2) * row_b;
    const int64_t blocks = static_cast<int64_t>(n_win) * n_bands * (D / 16);
    if (blocks > 0x7fffffff) return "conv1_from_coarse: grid too large";
    e = launch_pdl(fp16 ? conv1_from_coarse_x2_kernel<D, true> : conv1_from_coarse_x2_kernel<D, false>,
                   dim3(static_cast<unsigned>(blocks)), dim3(threads), static_cast<size_t>(smem), stream, 1,
                   static_cast<const uint16_t*>(Z), bias, n_win, hp, wp, gw, nb, n_bands, rs, static_cast<uint16_t*>(D1));
  } else if (whole <= 57344) {
    const int64_t blocks = static_cast<int64_t>(n_win) * (D / 16);
    e = launch_pdl(fp16 ? conv1_from_coarse_kernel<D, 16, true> : conv1_from_coarse_kernel<D, 16, false>,
                   dim3(static_cast<unsigned>(blocks)), dim3(256), static_cast<size_t>(whole), stream, 1,
                   static_cast<const uint16_t*>(Z), bias, n_win, hp, wp, gh, gw, gh + 1, gw + 1, 1, 1, static_cast<uint16_t*>(D1));
  } else {
    const int bh = 4, bw = 32;
    const int n_bands = (gh + 1 + bh - 1) / bh, n_ctiles = (gw + 1 + bw - 1) / bw;
    const int64_t blocks = static_cast<int64_t>(n_win) * n_bands * n_ctiles * (D / 32);
    if (blocks > 0x7fffffff) return "conv1_from_coarse: grid too large";
    e = launch_pdl(fp16 ? conv1_from_coarse_kernel<D, 32, true> : conv1_from_coarse_kernel<D, 32, false>,
                   dim3(static_cast<unsigned>(blocks)), dim3(256), static_cast<size_t>(kC1Smem), stream, 1,
                   static_cast<const uint16_t*>(Z), bias, n_win, hp, wp, gh, gw, bh, bw, n_bands, n_ctiles, static_cast<uint16_t*>(D1));
  }
  return e != cudaSuccess ? cudaGetErrorString(e) : last_err();
}
}  // namespace

const char* conv1_from_coarse(cudaStream_t stream, int width, const void* Z, const float* bias, int n_win, int hp, int wp,
                              int gh, int gw, void* D1, int fp16) {
  if (n_win <= 0) return "conv1_from_coarse: no windows";
  if (gh < 2 * hp || gw < 2 * wp) return "conv1_from_coarse: the decoder grid must be at least twice as fine as the patch grid";
  if (width == 768) return conv1_from_coarse_t<768>(stream, Z, bias, n_win, hp, wp, gh, gw, D1, fp16);
  if (width == 1024) return conv1_from_coarse_t<1024>(stream, Z, bias, n_win, hp, wp, gh, gw, D1, fp16);
  return "conv1_from_coarse: width must be 768 or 1024";
}

const char* fold_conv3x3_bn_tapout(cudaStream_t stream, const float* W, const float* gamma, const float* var, float eps, int O,
                                   int I, void* Wz, int fp16) {
  LaunchScope scope(stream, "pack");
  fold_conv3x3_bn_tapout_kernel<<<grid_for(static_cast<int64_t>(O) * I * 9, 256, 4096), 256, 0, stream>>>(
      W, gamma, var, eps, O, I, static_cast<uint16_t*>(Wz), fp16);
  return last_err();
}

const char* f32_to_16(cudaStream_t stream, const float* in, void* out, int64_t n, int fp16) {
  if (n <= 0) return nullptr;
  LaunchScope scope(stream, "pack");
  f32_to_16_kernel<<<grid_for(n, 256, 4096), 256, 0, stream>>>(in, static_cast<uint16_t*>(out), n, fp16);
  return last_err();
}

const char* fold_conv3x3_bn(cudaStream_t stream, const float* W, const float* gamma, const float* beta, const float* mean,
                            const float* var, float eps, int O, int I, void* Wp, float* bias, int fp16) {
  LaunchScope scope(stream, "pack");
  fold_conv3x3_bn_kernel<<<grid_for(static_cast<int64_t>(O) * I * 9, 256, 4096), 256, 0, stream>>>(
      W, gamma, beta, mean, var, eps, O, I, static_cast<uint16_t*>(Wp), bias, fp16);
  return last_err();
}

const char* split_weight_hi_hi_lo(cudaStream_t stream, const float* W, int O, int I, int Ip, void* out, int fp16) {
  if (Ip < I) return "split_weight: padded length below the row length";
  LaunchScope scope(stream, "pack");
  split_weight_kernel<<<grid_for(static_cast<int64_t>(O) * Ip, 256, 4096), 256, 0, stream>>>(
      W, O, I, Ip, static_cast<uint16_t*>(out), fp16);
  return last_err();
}

const char* pack_text(cudaStream_t stream, const float* text, const float* logit_scale, int n, int d, float* tmat, int ld_out) {
  if (n <= 0) return "pack_text: no bins";
  if (ld_out == 0) ld_out = d;
  if (ld_out < d) return "pack_text: output pitch below the row length";
  LaunchScope scope(stream, "pack");
  pack_text_kernel<<<(n + 3) / 4, 128, 0, stream>>>(text, logit_scale, n, d, tmat, ld_out);
  return last_err();
}

}  // namespace cebc
