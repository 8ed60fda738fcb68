// Throughput of MUFU.EX2 and of the f32x2 -> bf16x2 conversion per SM sub-partition on sm_100a, alone and mixed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_probe profiles/probes/mufu_probe.cu && /tmp/mufu_probe
// One CTA on one SM, W warps per sub-partition (blockDim = 128 * W), every thread runs `iters` rounds of 16 independent ops.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(float* out, long long* cycles, int iters, float seed) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) {  // ex2 only
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      } else if (MODE == 1) {  // cvt pack only
        uint32_t p;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(x[i]), "f"(x[(i + 1) & 15]));
        acc ^= p;
      } else if (MODE == 2) {  // ex2 + the softmax's companions: fma, min, add, and a pack per pair
        float y = fminf(fmaf(x[i], 0.18f, -0.5f), 120.f);
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(y));
        x[i] = y * 0.5f;
        if (i & 1) {
          uint32_t p;
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(x[i]), "f"(x[i - 1]));
          acc ^= p;
        }
      } else if (MODE == 3) {  // fma only (reference: full-rate pipe)
        x[i] = fmaf(x[i], 0.999f, 0.001f);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc & 0x7fffff);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&cyc, 1024);
  const int iters = 2000;
  const char* names[4] = {"ex2.approx.ftz.f32", "cvt.rn.bf16x2.f32", "softmax mix (fma+min+ex2+mul, pack per pair)", "fma"};
  for (int mode = 0; mode < 4; ++mode)
    for (int w = 1; w <= 4; w *= 2) {
      long long c = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) probe<0><<<1, 128 * w>>>(out, cyc, iters, 0.3f);
        if (mode == 1) probe<1><<<1, 128 * w>>>(out, cyc, iters, 0.3f);
        if (mode == 2) probe<2><<<1, 128 * w>>>(out, cyc, iters, 0.3f);
        if (mode == 3) probe<3><<<1, 128 * w>>>(out, cyc, iters, 0.3f);
        cudaDeviceSynchronize();
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      }
      const double per = static_cast<double>(c) / (static_cast<double>(iters) * 16 * w);
      printf("%-48s %d warp(s)/sub-partition: %.2f cycles per warp instruction group (%.1f lanes/clk/sub-partition)\n",
             names[mode], w, per, 32.0 / per);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
