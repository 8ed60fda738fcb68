// Second half of the EBC head (the first half -- ||f||^2 and the bin dot products -- is the epilogue of the projection GEMM,
// gemm2_tcgen05.cu EPI_BIAS_HEAD_PARTIAL): L2-normalise -> cosine logits -> softmax -> anchor expectation; the
// overlapping-window fold/average as an atomic-free gather; the per-image count reduction.
#include "common.cuh"
#include "kernels.h"

namespace cebc {

namespace {

// Second half of the fused head: one thread per interior cell sums the per-half-tile partials of the projection GEMM
// (EPI_BIAS_HEAD_PARTIAL) in a fixed order and finishes normalise / softmax / expectation (models/clip/model.py:203-212).
__global__ void __launch_bounds__(256) ebc_head_finish_kernel(const float* __restrict__ partial, int n_part,
                                                              const float* __restrict__ anchors, int n_bins, int n_win,
                                                              int gh, int gw, float* __restrict__ exp_out,
                                                              float* __restrict__ logits_out) {
  pdl_launch_dependents();
  pdl_wait();
  const int Hp = gh + 1, Wp = gw + 1;  // shared-border grid (kernels.h: resample_to_padded)
  const int S = 1 + n_bins;
  const int64_t n_cells = static_cast<int64_t>(n_win) * gh * gw;
  for (int64_t cell = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; cell < n_cells;
       cell += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int win = static_cast<int>(cell / (gh * gw));
    const int q = static_cast<int>(cell - static_cast<int64_t>(win) * gh * gw);
    const int y = q / gw, x = q - y * gw;
    const int64_t row = (static_cast<int64_t>(win) * Hp + y) * Wp + x;
    const float* pr = partial + row * n_part * S;
    float ss = 0.f;
    for (int k = 0; k < n_part; ++k) ss += pr[k * S];
    const float inv_norm = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    float lg[32];
    float mx = -INFINITY;
#pragma unroll 1
    for (int b = 0; b < n_bins; ++b) {
      float d = 0.f;
      for (int k = 0; k < n_part; ++k) d += pr[k * S + 1 + b];
      d *= inv_norm;
      lg[b] = d;
      mx = fmaxf(mx, d);
    }
    float den = 0.f, num = 0.f;
#pragma unroll 1
    for (int b = 0; b < n_bins; ++b) {
      const float e = __expf(lg[b] - mx);
      den += e;
      num += e * __ldg(anchors + b);
      if (logits_out != nullptr) logits_out[((static_cast<int64_t>(win) * n_bins + b) * gh + y) * gw + x] = lg[b];
    }
    exp_out[cell] = num / den;
  }
}

// One thread per output cell; windows covering the cell are visited in ascending window index (row-major i, j), which
// reproduces the fp32 summation order of the reference loop (utils/eval_utils.py:79-95) bit for bit. No atomics.
__global__ void __launch_bounds__(256) fold_average_kernel(const float* __restrict__ preds,
                                                           const int* __restrict__ row_cells,
                                                           const int* __restrict__ col_cells, int n_rows, int n_cols,
                                                           int gh, int gw, int Ho, int Wo, float* __restrict__ density) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = static_cast<int64_t>(Ho) * Wo;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int oy = static_cast<int>(idx / Wo), ox = static_cast<int>(idx - static_cast<int64_t>(oy) * Wo);
    float sum = 0.f, cnt = 0.f;
    for (int i = 0; i < n_rows; ++i) {
      const int dy = oy - __ldg(row_cells + i);
      if (dy < 0 || dy >= gh) continue;
      for (int j = 0; j < n_cols; ++j) {
        const int dx = ox - __ldg(col_cells + j);
        if (dx < 0 || dx >= gw) continue;
        sum += preds[(static_cast<int64_t>(i) * n_cols + j) * gh * gw + dy * gw + dx];
        cnt += 1.0f;
      }
    }
    density[idx] = sum / cnt;  // cnt == 0 -> NaN, as numpy's 0/0 in the reference
  }
}

// Deterministic single-block sum (fixed association order, independent of grid size / GPU count).
__global__ void __launch_bounds__(1024) sum_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sh[1024];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s += x[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

inline const char* last_err() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace

const char* ebc_head_finish(cudaStream_t stream, const float* partial, int n_part, const float* anchors, int n_bins, int n_win,
                            int gh, int gw, float* exp_out, float* logits_out) {
  if (n_bins < 1 || n_bins > 32) return "ebc_head: 1..32 bins supported";
  if (n_win <= 0 || n_part < 1) return "ebc_head: no windows";
  const int64_t cells = static_cast<int64_t>(n_win) * gh * gw;
  int64_t blocks = (cells + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  LaunchScope scope(stream, "ebc_head_finish", 0.0, static_cast<double>(cells) * (n_part * (1.0 + n_bins) * 4.0 + 4.0));
  cudaError_t e = launch_pdl(ebc_head_finish_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, 1, partial,
                             n_part, anchors, n_bins, n_win, gh, gw, exp_out, logits_out);
  return e != cudaSuccess ? cudaGetErrorString(e) : last_err();
}

const char* fold_average(cudaStream_t stream, const float* preds, const int* row_cells_dev, const int* col_cells_dev,
                         int n_rows, int n_cols, int gh, int gw, int Ho, int Wo, float* density, float* count_out) {
  if (n_rows <= 0 || n_cols <= 0 || Ho <= 0 || Wo <= 0) return "fold: empty problem";
  const int64_t total = static_cast<int64_t>(Ho) * Wo;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  {
    LaunchScope scope(stream, "fold", 0.0, 4.0 * (static_cast<double>(n_rows) * n_cols * gh * gw + static_cast<double>(total)));
    cudaError_t le = launch_pdl(fold_average_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, 1, preds,
                                row_cells_dev, col_cells_dev, n_rows, n_cols, gh, gw, Ho, Wo, density);
    if (le != cudaSuccess) return cudaGetErrorString(le);
  }
  const char* e = last_err();
  if (e) return e;
  if (count_out != nullptr) {
    LaunchScope scope(stream, "count_sum", 0.0, 4.0 * static_cast<double>(total));
    cudaError_t le = launch_pdl(sum_kernel, dim3(1), dim3(1024), 0, stream, 1, density, total, count_out);
    return le != cudaSuccess ? cudaGetErrorString(le) : last_err();
  }
  return nullptr;
}

}  // namespace cebc
