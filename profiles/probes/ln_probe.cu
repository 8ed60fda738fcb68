// Sweep of the streaming LayerNorm's ring depth / residency on the in-step shape (M = 64 x 197 rows of 768 fp32, X resident in
// L2 as inside a forward step), against the warp-per-row kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o profiles/probes/ln_probe profiles/probes/ln_probe.cu
#include <cstdio>
#include <vector>
#include "../../clip_ebc_b200/csrc/elementwise.cu"

namespace cebc {  // what api.cu / gemm2_tcgen05.cu provide in the library
LaunchScope::LaunchScope(cudaStream_t s, const char*, double, double) : stream_(s), slot_(-1) {}
LaunchScope::~LaunchScope() {}
bool pdl_enabled() { return true; }
const cudaAccessPolicyWindow* current_l2_window() { return nullptr; }
int device_num_sms() { return 148; }
}  // namespace cebc
using namespace cebc;

template <int STAGES, int CTAS>
float run(const float* x, const float* g, const float* b, uint16_t* out, int64_t rows, int reps) {
  auto kern = layernorm_stream_kernel<768, STAGES, CTAS>;
  const int smem = LnStream<768, STAGES>::kSmem;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1.f;
  int64_t blocks = (rows + kLnRows - 1) / kLnRows;
  if (blocks > 148 * CTAS) blocks = 148 * CTAS;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) launch_pdl(kern, dim3((unsigned)blocks), dim3(256), smem, 0, 1, x, g, b, out, rows, 1);
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) launch_pdl(kern, dim3((unsigned)blocks), dim3(256), smem, 0, 1, x, g, b, out, rows, 1);
  cudaEventRecord(e1);
  if (cudaDeviceSynchronize() != cudaSuccess) return -2.f;
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  return 1e3f * ms / reps;
}

int main() {
  for (int windows : {64, 96}) {
    const int64_t rows = (int64_t)windows * 197;
    float *x, *g, *b;
    uint16_t* out;
    cudaMalloc(&x, rows * 768 * 4); cudaMalloc(&g, 768 * 4); cudaMalloc(&b, 768 * 4); cudaMalloc(&out, rows * 768 * 2);
    std::vector<float> h(rows * 768);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)((i * 2654435761u) % 2001) / 1000.f - 1.f;
    cudaMemcpy(x, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    std::vector<float> one(768, 1.f), zero(768, 0.f);
    cudaMemcpy(g, one.data(), 768 * 4, cudaMemcpyHostToDevice); cudaMemcpy(b, zero.data(), 768 * 4, cudaMemcpyHostToDevice);
    const double mb = rows * 768 * 6.0 / 1e6;
    printf("%d windows: %lld rows, %.1f MB per launch (PDL-chained back-to-back launches, X L2-resident)\n", windows, (long long)rows, mb);
    auto show = [&](const char* name, float us) { printf("  %-34s %7.2f us  %6.0f GB/s\n", name, us, mb / us * 1e3); };
    show("stream 3 stages x 2 CTAs (shipped)", run<3, 2>(x, g, b, out, rows, 200));
    show("stream 2 stages x 2 CTAs", run<2, 2>(x, g, b, out, rows, 200));
    show("stream 4 stages x 2 CTAs", run<4, 2>(x, g, b, out, rows, 200));
    show("stream 2 stages x 3 CTAs", run<2, 3>(x, g, b, out, rows, 200));
    show("stream 3 stages x 3 CTAs", run<3, 3>(x, g, b, out, rows, 200));
    show("stream 2 stages x 4 CTAs", run<2, 4>(x, g, b, out, rows, 200));
    {
      auto kern = layernorm_kernel<768, true>;
      const int blocks = 148 * 8 < (rows + 7) / 8 ? 148 * 8 : (int)((rows + 7) / 8);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int i = 0; i < 5; ++i) launch_pdl(kern, dim3(blocks), dim3(256), 0, 0, 1, x, g, b, (void*)out, rows, 1, 1, 0, 1, (uint16_t*)nullptr);
      cudaEventRecord(e0);
      for (int i = 0; i < 200; ++i) launch_pdl(kern, dim3(blocks), dim3(256), 0, 0, 1, x, g, b, (void*)out, rows, 1, 1, 0, 1, (uint16_t*)nullptr);
      cudaEventRecord(e1); cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      show("warp per row (no ring)", 1e3f * ms / 200);
    }
    cudaFree(x); cudaFree(g); cudaFree(b); cudaFree(out);
  }
  return 0;
}
