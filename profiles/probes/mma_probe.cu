// Cost of one tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) in a long stream issued by one thread, by shape and
// operand source -- the shapes of the attention kernel:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I clip_ebc_b200/csrc -o profiles/probes/_bin/mma_probe profiles/probes/mma_probe.cu
// One CTA. Operands are zeros in shared memory / TMEM (timing does not depend on the values).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include "common.cuh"

using namespace cebc;

__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr_bytes) {  // MN-major, 128B swizzle, 64 columns (attention_pp.cu)
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr_bytes & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(32768 >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn ? 1u : 0u) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d), "r"(a), "l"(db), "r"(id), "r"(acc) : "memory");
}

// mode: 0 SS K-major B, 1 SS MN-major B, 2 TS K-major B, 3 TS MN-major B.  n_acc accumulators used round-robin.
__global__ void __launch_bounds__(128, 1) probe(int mode, int N, int n_acc, int count, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_ptr);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem) + 32768;
    const uint32_t id = idesc(128, N, mode & 1);
    const uint32_t d_cols = N;  // accumulator i at column 256 + i * N (TS: A operand at columns [0, 128))
    for (int rep = 0; rep < 2; ++rep) {
      const long long t0 = clock64();
      // descriptors precomputed, straight-line groups of 8: the loop must not be what is measured
      uint64_t dbv[4], dav[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        dbv[q] = (mode & 1) ? desc_mn(b_addr + q * 2048) : umma_desc_sw128_kmajor(b_addr + q * 32);
        dav[q] = umma_desc_sw128_kmajor(a_addr + q * 32);
      }
      const uint32_t d0 = tb + 256, d1 = tb + 256 + (n_acc > 1 ? d_cols : 0);
      for (int i = 0; i < count; i += 8) {
        const uint32_t acc = i > 0 ? 1u : 0u;
        if (mode & 2) {
#pragma unroll
          for (int q = 0; q < 8; ++q) mma_ts((q & 1) ? d1 : d0, tb + q * 8, dbv[q & 3], id, (q < 2) ? acc : 1u);
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) umma_bf16_ss((q & 1) ? d1 : d0, dav[q & 3], dbv[q & 3], id, (q < 2) ? acc : 1u);
        }
      }
      const long long t_issued = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, rep & 1);
      const long long t1 = clock64();
      out[0] = t1 - t0;
      out[1] = t_issued - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tb); }
}

int main() {
  long long* out;
  cudaMalloc(&out, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* names[4] = {"SS, B K-major ", "SS, B MN-major", "TS, B K-major ", "TS, B MN-major"};
  const int count = 512;
  for (int mode = 0; mode < 4; ++mode)
    for (int N : {256, 128, 64, 16})
      for (int n_acc : {1, 2}) {
        if ((mode & 1) && N != 64) continue;  // the MN-major descriptor above is written for 64 columns
        if (N * n_acc > 256) continue;
        probe<<<1, 128, 100 * 1024>>>(mode, N, n_acc, count, out);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("%s  M=128 N=%3d K=16, %d accumulator(s): %6.1f cycles per MMA (issue alone %5.1f)  floor %3d  %s\n", names[mode], N,
               n_acc, static_cast<double>(h[0]) / count, static_cast<double>(h[1]) / count, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
