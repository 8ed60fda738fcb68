// Element-wise kernels of the CLIP-ResNet encoder path (reference
// /root/reference/models/clip/_clip/image_encoder.py:10-115, _clip/blocks.py:56-101). Every convolution of that path runs on
// the tcgen05 GEMM (gemm2_tcgen05.cu) over 16-bit NHWC activations laid out on shared-border grids (kernels.h:
// resample_to_padded); what is left for this file is data movement: the im2col of the stride-2 stem convolution, average
// pooling / copies between grids, bilinear resampling of 16-bit maps, and the pack-time BatchNorm fold.
#include "common.cuh"
#include "kernels.h"

namespace cebc {

namespace {

__device__ __forceinline__ float2 unpack16x2(uint32_t u, int fp16) {
  if (fp16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u));
}

// conv1 of the stem: 3 -> 32 channels, 3x3, stride 2, padding 1 (image_encoder.py:37) as a GEMM over im2col rows. One row
// per cell of the (h/2 + 1) x (w/2 + 1) shared-border grid of a unit: 64 columns, k = c * 9 + ky * 3 + kx < 27 holds pixel
// (2y - 1 + ky, 2x - 1 + kx) of the unit (zero outside it: the reference slices the window first, then pads), the rest zero.
// Unit u = image u of a batch [n, 3, H, W] (origins == nullptr) or the window of ONE image at origins[2u], origins[2u + 1].
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ image, int n_units, int H, int W,
                                                          const int* __restrict__ origins_yx, int h, int w,
                                                          uint16_t* __restrict__ out, int fp16) {
  pdl_launch_dependents();
  pdl_wait();
  const int gh = h / 2, gw = w / 2, Hp = gh + 1, Wp = gw + 1;
  // grid: y = unit, x = (cell, 8-column group) of the unit -- a thread's position is two 32-bit divisions (the flat 64-bit index
  // of the first version cost two 64-bit ones per thread: these kernels are issue-bound, not bandwidth-bound)
  const unsigned per_unit = static_cast<unsigned>(Hp) * Wp * 8;
  for (int unit = blockIdx.y; unit < n_units; unit += gridDim.y)
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < per_unit; i += gridDim.x * blockDim.x) {
    const int g8 = static_cast<int>(i & 7);
    const int q = static_cast<int>(i >> 3);
    const int64_t row = static_cast<int64_t>(unit) * Hp * Wp + q;
    const int y = static_cast<int>(static_cast<unsigned>(q) / static_cast<unsigned>(Wp)), x = q - y * Wp;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (y < gh && x < gw && g8 < 4) {
      const float* base;
      int oy = 0, ox = 0;
      if (origins_yx != nullptr) { oy = origins_yx[2 * unit]; ox = origins_yx[2 * unit + 1]; base = image; }
      else base = image + static_cast<int64_t>(unit) * 3 * H * W;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = g8 * 8 + j;
        if (k < 27) {
          const int c = k / 9, t = k - c * 9, ky = t / 3, kx = t - ky * 3;
          const int iy = 2 * y - 1 + ky, ix = 2 * x - 1 + kx;
          if (iy >= 0 && iy < h && ix >= 0 && ix < w) v[j] = __ldg(base + (static_cast<int64_t>(c) * H + oy + iy) * W + ox + ix);
        }
      }
    }
    *reinterpret_cast<uint4*>(out + row * 64 + g8 * 8) =
        make_uint4(pack16x2(v[0], v[1], fp16), pack16x2(v[2], v[3], fp16), pack16x2(v[4], v[5], fp16), pack16x2(v[6], v[7], fp16));
  }
}

// dst[unit, y, x, dst_col + c] = mean over the S x S block of src[unit, S y + dy, S x + dx, src_col + c] (S = 1: a copy; S = 2:
// nn.AvgPool2d(2)), c < C, both maps 16-bit NHWC on shared-border grids ((gi + 1)^2 rows in, (go + 1)^2 rows out per unit,
// gi = S * go); the border row / column of dst is written as zero. Thread = (dst cell, 8 channels).
__global__ void __launch_bounds__(256) pool_copy_kernel(const uint16_t* __restrict__ src, int ld_src, int src_col,
                                                        uint16_t* __restrict__ dst, int ld_dst, int dst_col, int C, int n_units,
                                                        int go_h, int go_w, int S, int fp16) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = go_h + 1, Wo = go_w + 1, Wi = S * go_w + 1, Hi = S * go_h + 1;
  const int groups = C >> 3;
  const float inv = 1.0f / static_cast<float>(S * S);
  const unsigned per_unit = static_cast<unsigned>(Ho) * Wo * groups;  // grid: y = unit, x = (cell, channel group) of the unit
  for (int unit = blockIdx.y; unit < n_units; unit += gridDim.y)
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < per_unit; i += gridDim.x * blockDim.x) {
    const int q = static_cast<int>(i / static_cast<unsigned>(groups));
    const int g = static_cast<int>(i) - q * groups;
    const int64_t row = static_cast<int64_t>(unit) * Ho * Wo + q;
    const int y = static_cast<int>(static_cast<unsigned>(q) / static_cast<unsigned>(Wo)), x = q - y * Wo;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (y < go_h && x < go_w) {
      const uint16_t* s0 = src + (static_cast<int64_t>(unit) * Hi * Wi) * ld_src + src_col + g * 8;
      if (S == 1) {
        o = *reinterpret_cast<const uint4*>(s0 + (static_cast<int64_t>(y) * Wi + x) * ld_src);
      } else {
        float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        for (int dy = 0; dy < S; ++dy)
          for (int dx = 0; dx < S; ++dx) {
            const uint4 u = *reinterpret_cast<const uint4*>(s0 + (static_cast<int64_t>(S * y + dy) * Wi + S * x + dx) * ld_src);
            const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = unpack16x2(uu[k], fp16);
              acc[k].x += f.x; acc[k].y += f.y;
            }
          }
        o = make_uint4(pack16x2(acc[0].x * inv, acc[0].y * inv, fp16), pack16x2(acc[1].x * inv, acc[1].y * inv, fp16),
                       pack16x2(acc[2].x * inv, acc[2].y * inv, fp16), pack16x2(acc[3].x * inv, acc[3].y * inv, fp16));
      }
    }
    *reinterpret_cast<uint4*>(dst + row * ld_dst + dst_col + g * 8) = o;
  }
}

// F.interpolate(mode="bilinear", align_corners=False) of a 16-bit NHWC map between shared-border grids (gi -> go per axis,
// any ratio; models/clip/model.py:195-196 for the ResNet encoders: x2 at reduction 8). Thread = (dst cell, 8 channels).
__global__ void __launch_bounds__(256) resample16_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int C,
                                                         int n_units, int gi_h, int gi_w, int go_h, int go_w, int fp16) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = go_h + 1, Wo = go_w + 1, Hi = gi_h + 1, Wi = gi_w + 1;
  const int groups = C >> 3;
  const float inv_sy = static_cast<float>(gi_h) / static_cast<float>(go_h), inv_sx = static_cast<float>(gi_w) / static_cast<float>(go_w);
  const unsigned per_unit = static_cast<unsigned>(Ho) * Wo * groups;  // grid: y = unit, x = (cell, channel group) of the unit
  for (int unit = blockIdx.y; unit < n_units; unit += gridDim.y)
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < per_unit; i += gridDim.x * blockDim.x) {
    const int q = static_cast<int>(i / static_cast<unsigned>(groups));
    const int g = static_cast<int>(i) - q * groups;
    const int64_t row = static_cast<int64_t>(unit) * Ho * Wo + q;
    const int y = static_cast<int>(static_cast<unsigned>(q) / static_cast<unsigned>(Wo)), x = q - y * Wo;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (y < go_h && x < go_w) {
      int y0, y1, x0, x1;
      float ly, lx;
      bilinear_src(y, inv_sy, gi_h, y0, y1, ly);
      bilinear_src(x, inv_sx, gi_w, x0, x1, lx);
      const uint16_t* s0 = src + (static_cast<int64_t>(unit) * Hi * Wi) * C + g * 8;
      const uint4 a = *reinterpret_cast<const uint4*>(s0 + (static_cast<int64_t>(y0) * Wi + x0) * C);
      const uint4 b = *reinterpret_cast<const uint4*>(s0 + (static_cast<int64_t>(y0) * Wi + x1) * C);
      const uint4 c = *reinterpret_cast<const uint4*>(s0 + (static_cast<int64_t>(y1) * Wi + x0) * C);
      const uint4 d = *reinterpret_cast<const uint4*>(s0 + (static_cast<int64_t>(y1) * Wi + x1) * C);
      const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
      const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w}, uc[4] = {c.x, c.y, c.z, c.w}, ud[4] = {d.x, d.y, d.z, d.w};
      uint32_t r[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 fa = unpack16x2(ua[k], fp16), fb = unpack16x2(ub[k], fp16), fc = unpack16x2(uc[k], fp16), fd = unpack16x2(ud[k], fp16);
        r[k] = pack16x2(w00 * fa.x + w01 * fb.x + w10 * fc.x + w11 * fd.x, w00 * fa.y + w01 * fb.y + w10 * fc.y + w11 * fd.y, fp16);
      }
      o = make_uint4(r[0], r[1], r[2], r[3]);
    }
    *reinterpret_cast<uint4*>(dst + row * C + g * 8) = o;
  }
}

// Pack time. W f32 [O, I, taps] (a conv weight [O, I, kh, kw] with taps = kh * kw) and the BatchNorm behind it ->
// Wp[o * ldw + col_off + tap * i_pad + i] = 16-bit(W[o, i, tap] * gamma[o] / sqrt(var[o] + eps)), bias[o] (+)= beta[o] -
// mean[o] * gamma[o] / sqrt(var[o] + eps). Pad columns / rows of Wp are not touched (the caller zero-fills the buffer).
// tap_major_src = 1: W is [O, taps * I] with k = c * taps + tap handled by the caller (im2col ordering) -> taps = 1, I = 27.
__global__ void fold_conv_bn_general_kernel(const float* __restrict__ W, const float* __restrict__ gamma,
                                            const float* __restrict__ beta, const float* __restrict__ mean,
                                            const float* __restrict__ var, float eps, int O, int I, int taps, int i_pad,
                                            uint16_t* __restrict__ Wp, int ldw, int col_off, float* __restrict__ bias,
                                            int accumulate_bias, int fp16) {
  const int64_t total = static_cast<int64_t>(O) * I * taps;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx % I);
    const int tap = static_cast<int>((idx / I) % taps);
    const int o = static_cast<int>(idx / (static_cast<int64_t>(I) * taps));
    const float sc = gamma != nullptr ? gamma[o] / sqrtf(var[o] + eps) : 1.0f;  // gamma == nullptr: no BatchNorm, no bias
    Wp[static_cast<int64_t>(o) * ldw + col_off + tap * i_pad + i] = cvt16(W[(static_cast<int64_t>(o) * I + i) * taps + tap] * sc, fp16);
    if (gamma != nullptr && i == 0 && tap == 0) {
      const float b = beta[o] - mean[o] * sc;
      bias[o] = accumulate_bias ? bias[o] + b : b;
    }
  }
}

inline int grid_1d(int64_t items, int cap) {
  int64_t b = (items + 255) / 256;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return static_cast<int>(b);
}
inline const char* last_err() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace

// grid of the (unit, flat index inside the unit) kernels: y = units (grid-strided beyond 65535), x = enough 256-thread blocks to
// cover a unit, capped so that the whole grid stays near 16 blocks per SM
dim3 grid_units(int n_units, int64_t per_unit) {
  const int64_t need = (per_unit + 255) / 256;
  const int gy = n_units < 65535 ? n_units : 65535;
  int64_t gx = (static_cast<int64_t>(device_num_sms()) * 16 + gy - 1) / gy;
  gx = gx < 1 ? 1 : (gx > need ? need : gx);
  return dim3(static_cast<unsigned>(gx), static_cast<unsigned>(gy));
}

const char* stem_im2col(cudaStream_t stream, const float* image, int n_units, int H, int W, const int* origins_yx_dev, int h,
                        int w, void* out, int fp16) {
  if (n_units <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1)) return "stem_im2col: bad geometry";
  const int64_t rows = static_cast<int64_t>(n_units) * (h / 2 + 1) * (w / 2 + 1);
  LaunchScope scope(stream, "stem_im2col", 0.0, static_cast<double>(n_units) * 3 * h * w * 4.0 + static_cast<double>(rows) * 128.0);
  cudaError_t e = launch_pdl(stem_im2col_kernel, grid_units(n_units, static_cast<int64_t>(h / 2 + 1) * (w / 2 + 1) * 8), dim3(256), 0, stream, 1, image,
                             n_units, H, W, origins_yx_dev, h, w, static_cast<uint16_t*>(out), fp16);
  return e != cudaSuccess ? cudaGetErrorString(e) : last_err();
}

const char* pool_copy(cudaStream_t stream, const void* src, int ld_src, int src_col, void* dst, int ld_dst, int dst_col, int C,
                      int n_units, int go_h, int go_w, int S, int fp16) {
  if (n_units <= 0 || C <= 0 || (C & 7) || (S != 1 && S != 2)) return "pool_copy: bad arguments";
  if ((ld_src & 7) || (ld_dst & 7) || (src_col & 7) || (dst_col & 7)) return "pool_copy: pitches / offsets must be multiples of 8";
  const int64_t items = static_cast<int64_t>(n_units) * (go_h + 1) * (go_w + 1) * (C / 8);
  LaunchScope scope(stream, S == 1 ? "copy" : "avgpool", 0.0, static_cast<double>(items) * 16.0 * (1 + S * S));
  cudaError_t e = launch_pdl(pool_copy_kernel, grid_units(n_units, items / n_units), dim3(256), 0, stream, 1,
                             static_cast<const uint16_t*>(src), ld_src, src_col, static_cast<uint16_t*>(dst), ld_dst, dst_col, C,
                             n_units, go_h, go_w, S, fp16);
  return e != cudaSuccess ? cudaGetErrorString(e) : last_err();
}

const char* resample16(cudaStream_t stream, const void* src, void* dst, int C, int n_units, int gi_h, int gi_w, int go_h, int go_w,
                       int fp16) {
  if (n_units <= 0 || C <= 0 || (C & 7)) return "resample16: bad arguments";
  const int64_t items = static_cast<int64_t>(n_units) * (go_h + 1) * (go_w + 1) * (C / 8);
  LaunchScope scope(stream, "resample", 0.0, static_cast<double>(items) * 16.0 * 2);
  cudaError_t e = launch_pdl(resample16_kernel, grid_units(n_units, items / n_units), dim3(256), 0, stream, 1,
                             static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst), C, n_units, gi_h, gi_w, go_h, go_w, fp16);
  return e != cudaSuccess ? cudaGetErrorString(e) : last_err();
}

const char* fold_conv_bn_general(cudaStream_t stream, const float* W, const float* gamma, const float* beta, const float* mean,
                                 const float* var, float eps, int O, int I, int taps, int i_pad, void* Wp, int ldw, int col_off,
                                 float* bias, int accumulate_bias, int fp16) {
  if (O <= 0 || I <= 0 || taps <= 0 || i_pad < I) return "fold_conv_bn: bad arguments";
  LaunchScope scope(stream, "pack");
  fold_conv_bn_general_kernel<<<grid_1d(static_cast<int64_t>(O) * I * taps, 4096), 256, 0, stream>>>(
      W, gamma, beta, mean, var, eps, O, I, taps, i_pad, static_cast<uint16_t*>(Wp), ldw, col_off, bias, accumulate_bias, fp16);
  return last_err();
}

}  // namespace cebc
