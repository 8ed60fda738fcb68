// The steps either side of the hot path, all HBM-bound one-thread-per-output-element kernels
// with coalesced accesses along the image row:
//   * bicubic antialiased resize (Resize2Multiple -> TF.resize(BICUBIC, antialias=True), datasets/transforms.py:27-35,
//     69-105) as two separable passes (width, then height), with the uint8 -> [0,1] conversion fused into the first
//     pass and the ImageNet normalisation (datasets/crowd.py:64) fused into the second;
//   * zero padding to window + k * stride (ZeroPad2Multiple, datasets/transforms.py:108-140) fused with the same
//     conversion and normalisation;
//   * resize_density_map (utils/eval_utils.py:19-23): bilinear resize + sum-preserving rescale, with deterministic
//     two-stage sums (fixed chunking, fixed tree order) instead of atomics.
#include "common.cuh"
#include "kernels.h"

namespace cebc {

namespace {

struct ChanStats {
  float mean[4];
  float std[4];
  int normalize;
};

__device__ __forceinline__ float load_unit(const void* in, int is_u8, int64_t idx) {
  // uint8 pixels become floats in [0,1] exactly as `torch.from_numpy(image).float() / 255.` (datasets/crowd.py:218)
  return is_u8 ? static_cast<float>(static_cast<const uint8_t*>(in)[idx]) / 255.0f : static_cast<const float*>(in)[idx];
}

// PyTorch's antialias bicubic filter (a = -0.5; aten/src/ATen/native/UpSample.h cubic_convolution1/2)
__device__ __forceinline__ float aa_cubic(float x) {
  const float a = -0.5f;
  x = fabsf(x);
  if (x < 1.0f) return ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
  if (x < 2.0f) return ((a * x - 5.0f * a) * x + 8.0f * a) * x - 4.0f * a;
  return 0.0f;
}

// One output sample of a 1-D antialiased resize along a line of `in_size` samples with stride `in_stride`
// (aten/src/ATen/native/cpu/UpSampleKernel.cpp: _compute_indices_min_size_weights_aa + interpolate_aa_single_dim):
// weights are evaluated on the fly, normalised by their sum, and applied in ascending tap order.
template <class Load>
__device__ __forceinline__ float aa_sample(int o, int in_size, int out_size, Load load) {
  // float variables with the double-typed 0.5 literals of the ATen code, so that the integer tap ranges agree
  const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  const float support = (scale >= 1.0f) ? 2.0f * scale : 2.0f;
  const float invscale = (scale >= 1.0f) ? 1.0f / scale : 1.0f;
  const float center = static_cast<float>(static_cast<double>(scale) * (static_cast<double>(o) + 0.5));
  int xmin = static_cast<int>(static_cast<double>(center - support) + 0.5);
  xmin = xmin < 0 ? 0 : xmin;
  int xmax = static_cast<int>(static_cast<double>(center + support) + 0.5);
  xmax = xmax > in_size ? in_size : xmax;
  const int xsize = xmax - xmin;
  auto weight = [&](int j) {
    return aa_cubic(static_cast<float>((static_cast<double>(static_cast<float>(j + xmin) - center) + 0.5) * static_cast<double>(invscale)));
  };
  float total = 0.f;
  for (int j = 0; j < xsize; ++j) total += weight(j);
  const float wscale = total != 0.f ? 1.0f / total : 0.f;
  float acc = 0.f;
  for (int j = 0; j < xsize; ++j) {
    const float wgt = weight(j) * wscale;
    acc = (j == 0) ? load(xmin) * wgt : acc + load(xmin + j) * wgt;
  }
  return acc;
}

// pass 1: in [C, h, w] (u8 or f32) -> tmp f32 [C, h, W]
__global__ void __launch_bounds__(256) resize_aa_width_kernel(const void* __restrict__ in, int is_u8, int C, int h, int w,
                                                              int W, float* __restrict__ tmp) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = static_cast<int64_t>(C) * h * W;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(idx % W);
    const int64_t line = idx / W;  // c * h + y
    const int64_t base = line * w;
    tmp[idx] = aa_sample(ox, w, W, [&](int x) { return load_unit(in, is_u8, base + x); });
  }
}

// pass 2: tmp f32 [C, h, W] -> out f32 [C, H, W], optional (v - mean[c]) / std[c]
__global__ void __launch_bounds__(256) resize_aa_height_kernel(const float* __restrict__ tmp, int C, int h, int W, int H,
                                                               float* __restrict__ out, ChanStats st) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = static_cast<int64_t>(C) * H * W;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(idx % W);
    const int oy = static_cast<int>((idx / W) % H);
    const int c = static_cast<int>(idx / (static_cast<int64_t>(W) * H));
    const float* col = tmp + static_cast<int64_t>(c) * h * W + ox;
    float v = aa_sample(oy, h, H, [&](int y) { return col[static_cast<int64_t>(y) * W]; });
    if (st.normalize) v = (v - st.mean[c]) / st.std[c];
    out[idx] = v;
  }
}

// in [C, h, w] (u8 or f32) -> out f32 [C, H, W]: right/bottom zero padding (of the [0,1] image), then normalisation
__global__ void __launch_bounds__(256) pad_normalize_kernel(const void* __restrict__ in, int is_u8, int C, int h, int w,
                                                            int H, int W, float* __restrict__ out, ChanStats st) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = static_cast<int64_t>(C) * H * W;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(idx % W);
    const int oy = static_cast<int>((idx / W) % H);
    const int c = static_cast<int>(idx / (static_cast<int64_t>(W) * H));
    float v = 0.f;
    if (oy < h && ox < w) v = load_unit(in, is_u8, (static_cast<int64_t>(c) * h + oy) * w + ox);
    if (st.normalize) v = (v - st.mean[c]) / st.std[c];
    out[idx] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// resize_density_map
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRdmMaxBlocks = 1024;

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (static_cast<int>(threadIdx.x) < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  const float r = sh[0];
  __syncthreads();
  return r;
}

// PyTorch upsample_bilinear2d, align_corners=False: src = scale * (dst + 0.5) - 0.5 clamped at 0, scale = in / out
__device__ __forceinline__ void bilinear_src(int o, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = scale * (static_cast<float>(o) + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = static_cast<int>(src);
  i0 = i0 > in_size - 1 ? in_size - 1 : i0;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - static_cast<float>(i0);
  l0 = 1.0f - l1;
}

// stage 1: out = bilinear(x), partial sums of out and of x per block (block b owns a fixed contiguous chunk)
__global__ void __launch_bounds__(256) rdm_resize_kernel(const float* __restrict__ x, int h, int w, int H, int W,
                                                         float* __restrict__ out, float* __restrict__ part_out,
                                                         float* __restrict__ part_in) {
  __shared__ float sh[256];
  pdl_launch_dependents();
  pdl_wait();
  const float sy = static_cast<float>(h) / static_cast<float>(H), sx = static_cast<float>(w) / static_cast<float>(W);
  const int64_t n_out = static_cast<int64_t>(H) * W, n_in = static_cast<int64_t>(h) * w;
  const int64_t chunk_out = (n_out + gridDim.x - 1) / gridDim.x, chunk_in = (n_in + gridDim.x - 1) / gridDim.x;
  float s_out = 0.f, s_in = 0.f;
  const int64_t o_end = min(n_out, (static_cast<int64_t>(blockIdx.x) + 1) * chunk_out);
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * chunk_out + threadIdx.x; idx < o_end; idx += 256) {
    const int oy = static_cast<int>(idx / W), ox = static_cast<int>(idx - static_cast<int64_t>(oy) * W);
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    bilinear_src(oy, sy, h, y0, y1, ly0, ly1);
    bilinear_src(ox, sx, w, x0, x1, lx0, lx1);
    const float* r0 = x + static_cast<int64_t>(y0) * w;
    const float* r1 = x + static_cast<int64_t>(y1) * w;
    const float v = ly0 * (lx0 * r0[x0] + lx1 * r0[x1]) + ly1 * (lx0 * r1[x0] + lx1 * r1[x1]);
    out[idx] = v;
    s_out += v;
  }
  const int64_t i_end = min(n_in, (static_cast<int64_t>(blockIdx.x) + 1) * chunk_in);
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * chunk_in + threadIdx.x; idx < i_end; idx += 256) s_in += x[idx];
  s_out = block_sum_256(s_out, sh);
  s_in = block_sum_256(s_in, sh);
  if (threadIdx.x == 0) { part_out[blockIdx.x] = s_out; part_in[blockIdx.x] = s_in; }
}

// stage 2: every block reduces the partials in the same fixed order, then scales its slice:
// scale = nan_to_num(sum(out) / sum(x), nan=0, posinf=0, neginf=0)
__global__ void __launch_bounds__(256) rdm_scale_kernel(float* __restrict__ out, int64_t n_out, const float* __restrict__ part_out,
                                                        const float* __restrict__ part_in, int n_part,
                                                        float* __restrict__ sums_out /* nullable: [sum_in, sum_resized] */) {
  __shared__ float sh[256];
  pdl_launch_dependents();
  pdl_wait();
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < n_part; i += 256) { a += part_out[i]; b += part_in[i]; }
  const float tot_out = block_sum_256(a, sh);
  const float tot_in = block_sum_256(b, sh);
  float scale = tot_out / tot_in;
  if (isnan(scale) || isinf(scale)) scale = 0.f;
  if (blockIdx.x == 0 && threadIdx.x == 0 && sums_out != nullptr) { sums_out[0] = tot_in; sums_out[1] = tot_out; }
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < n_out;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[idx] *= scale;
}

inline int grid_1d(int64_t items, int max_blocks) {
  int64_t b = (items + 255) / 256;
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

ChanStats make_stats(int C, const float* mean, const float* std) {
  ChanStats st{};
  st.normalize = (mean != nullptr && std != nullptr) ? 1 : 0;
  for (int c = 0; c < 4; ++c) {
    st.mean[c] = (st.normalize && c < C) ? mean[c] : 0.f;
    st.std[c] = (st.normalize && c < C) ? std[c] : 1.f;
  }
  return st;
}

}  // namespace

const char* resize_bicubic_aa(cudaStream_t stream, const void* in, int in_is_u8, int C, int h, int w, float* tmp, float* out,
                              int H, int W, const float* mean_host, const float* std_host) {
  if (C < 1 || C > 4) return "resize: 1..4 channels supported";
  if (h <= 0 || w <= 0 || H <= 0 || W <= 0) return "resize: empty image";
  const int cap = device_num_sms() * 16;
  {
    LaunchScope scope(stream, "resize_aa_w", 0.0, static_cast<double>(C) * h * (static_cast<double>(w) * (in_is_u8 ? 1 : 4) + 4.0 * W));
    cudaError_t e = launch_pdl(resize_aa_width_kernel, dim3(grid_1d(static_cast<int64_t>(C) * h * W, cap)), dim3(256), 0, stream,
                               1, in, in_is_u8, C, h, w, W, tmp);
    if (e != cudaSuccess) return cudaGetErrorString(e);
  }
  {
    LaunchScope scope(stream, "resize_aa_h", 0.0, 4.0 * C * W * (static_cast<double>(h) + H));
    cudaError_t e = launch_pdl(resize_aa_height_kernel, dim3(grid_1d(static_cast<int64_t>(C) * H * W, cap)), dim3(256), 0, stream,
                               1, static_cast<const float*>(tmp), C, h, W, H, out, make_stats(C, mean_host, std_host));
    if (e != cudaSuccess) return cudaGetErrorString(e);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* pad_normalize(cudaStream_t stream, const void* in, int in_is_u8, int C, int h, int w, float* out, int H, int W,
                          const float* mean_host, const float* std_host) {
  if (C < 1 || C > 4) return "pad: 1..4 channels supported";
  if (h <= 0 || w <= 0) return "pad: empty image";
  if (H < h || W < w) return "pad: output smaller than input";
  LaunchScope scope(stream, "pad_normalize", 0.0, static_cast<double>(C) * (static_cast<double>(h) * w * (in_is_u8 ? 1 : 4) + 4.0 * H * W));
  cudaError_t e = launch_pdl(pad_normalize_kernel, dim3(grid_1d(static_cast<int64_t>(C) * H * W, device_num_sms() * 16)), dim3(256),
                             0, stream, 1, in, in_is_u8, C, h, w, H, W, out, make_stats(C, mean_host, std_host));
  if (e != cudaSuccess) return cudaGetErrorString(e);
  e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

int resize_density_workspace_floats() { return 2 * kRdmMaxBlocks; }

const char* resize_density_map(cudaStream_t stream, const float* x, int h, int w, int H, int W, float* out, float* workspace,
                               float* sums_out) {
  if (h <= 0 || w <= 0 || H <= 0 || W <= 0) return "resize_density_map: empty map";
  const int64_t n_out = static_cast<int64_t>(H) * W;
  const int nblk = grid_1d(n_out, kRdmMaxBlocks);
  float* part_out = workspace;
  float* part_in = workspace + kRdmMaxBlocks;
  {
    LaunchScope scope(stream, "resize_density", 0.0, 4.0 * (static_cast<double>(h) * w + static_cast<double>(n_out)));
    cudaError_t e = launch_pdl(rdm_resize_kernel, dim3(nblk), dim3(256), 0, stream, 1, x, h, w, H, W, out, part_out, part_in);
    if (e != cudaSuccess) return cudaGetErrorString(e);
  }
  {
    LaunchScope scope(stream, "resize_density_scale", 0.0, 8.0 * static_cast<double>(n_out));
    cudaError_t e = launch_pdl(rdm_scale_kernel, dim3(grid_1d(n_out, device_num_sms() * 8)), dim3(256), 0, stream, 1, out, n_out,
                               static_cast<const float*>(part_out), static_cast<const float*>(part_in), nblk, sums_out);
    if (e != cudaSuccess) return cudaGetErrorString(e);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace cebc
