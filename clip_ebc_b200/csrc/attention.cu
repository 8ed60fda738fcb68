// Streamed-K/V attention for the CLIP ViT blocks: softmax(q k^T / sqrt(64)) v per (window, head), 64-dim heads, no mask,
// ANY sequence length. Replaces nn.MultiheadAttention -> F.scaled_dot_product_attention of the reference
// (/root/reference/models/clip/_clip/blocks.py:25,35-37) for the windows the tcgen05 kernel (attention_pp.cu: one 256-key
// tile per item) does not take: more than 256 keys (224 x 224 windows of ViT-L/14: 257 live tokens + 32 prompts; 320 x 320
// or 448 x 448 windows of ViT-B/16) or a constant-key count that is not a multiple of 8.
//
// Deep-VPT rewrite (reference models/clip/model.py:164-183): the prompt tokens of layer l are constants that are
// discarded after the block, so they only ever act as keys/values. Their K/V rows are precomputed at pack time and
// appended here as `const_kv` keys; the per-window sequence holds only the live tokens (cls + patches).
#include "common.cuh"
#include "kernels.h"

namespace cebc {

namespace {

constexpr int kHeadDim = 64;
constexpr int kAttnThreads = 128;

__device__ __forceinline__ uint32_t sw_off(int row, int chunk) {  // byte offset inside a [rows][128 B] swizzled tile
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------------------------------------
// One CTA per (window, head, 64-query chunk); K / V stream through shared memory in double-buffered 64-key blocks
// (cp.async, XOR-swizzled 128 B rows), each warp owns 16 queries, scores stay in registers, softmax is online over the
// blocks. mma.sync m16n8k16: a completeness path, not a tuned one.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kLongQ = 64;                               // queries per CTA (4 warps x 16)
constexpr int kLongSmem = kLongQ * 128 + 2 * 2 * 64 * 128;  // Q chunk + 2 stages x (K, V) blocks = 40 KB

__global__ void __launch_bounds__(kAttnThreads) attention_h64_long_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                         const __nv_bfloat16* __restrict__ const_kv,
                                                                         int n_const, int t_live, int q_chunks, int heads,
                                                                         uint16_t* __restrict__ out, int out_fp16) {
  __shared__ __align__(128) uint8_t smem[kLongSmem];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sKV = sQ + kLongQ * 128;  // stage s: K at sKV + s * 16384, V at + 8192

  const int item = blockIdx.x / q_chunks, qc = blockIdx.x - item * q_chunks;
  const int win = item / heads, head = item - win * heads;
  const int width = heads * kHeadDim, kQkvLd = 3 * width;  // row pitch of qkv / const_kv; k at +width, v at +2 * width
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Tq = t_live, Tk = t_live + n_const;
  const int k_blocks = (Tk + 63) >> 6;
  const int qbase = qc * kLongQ;

  const __nv_bfloat16* base = qkv + static_cast<int64_t>(win) * t_live * kQkvLd + head * kHeadDim;
  for (int i = tid; i < kLongQ * 8; i += kAttnThreads) {
    const int row = i >> 3, ch = i & 7;
    const bool ok = qbase + row < Tq;
    cp_async_16(sQ + sw_off(row, ch), base + static_cast<int64_t>(ok ? qbase + row : 0) * kQkvLd + ch * 8, ok);
  }
  auto load_block = [&](int kb) {
    const uint32_t sK = sKV + (kb & 1) * 16384, sV = sK + 8192;
    for (int i = tid; i < 64 * 8; i += kAttnThreads) {
      const int row = i >> 3, ch = i & 7;
      const int key = kb * 64 + row;
      const bool ok = key < Tk;
      const __nv_bfloat16* src = base;
      if (ok) src = key < t_live ? base + static_cast<int64_t>(key) * kQkvLd
                                 : const_kv + static_cast<int64_t>(key - t_live) * kQkvLd + head * kHeadDim;
      cp_async_16(sK + sw_off(row, ch), src + width + ch * 8, ok);
      cp_async_16(sV + sw_off(row, ch), src + 2 * width + ch * 8, ok);
    }
  };
  load_block(0);
  cp_async_commit();

  const float kScaleLog2 = 0.125f * 1.4426950408889634f;
  const int g = lane >> 2, tq = lane & 3;
  const int q0 = warp << 4;  // this warp's 16 queries inside the chunk
  uint32_t qa[4][4];
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = 0.f; o[j][1] = 0.f; o[j][2] = 0.f; o[j][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  for (int kb = 0; kb < k_blocks; ++kb) {
    // everybody is done with the stage block kb + 1 goes to (it held block kb - 1)
    __syncthreads();
    if (kb + 1 < k_blocks) load_block(kb + 1);
    cp_async_commit();
    cp_async_wait<1>();  // block kb (and, first time, Q) has landed
    __syncthreads();
    if (kb == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int m = lane >> 3;
        ldmatrix_x4(qa[ks], sQ + sw_off(q0 + (m & 1) * 8 + (lane & 7), 2 * ks + (m >> 1)));
      }
    }
    // A warp whose 16 queries all lie beyond the sequence only takes part in the loads and barriers: 257 tokens (ViT-L/14)
    // are 4 chunks of 64 and one chunk with a single query, i.e. one busy warp in the fifth CTA.
    if (qbase + q0 >= Tq) continue;
    const uint32_t sK = sKV + (kb & 1) * 16384, sV = sK + 8192;
    const int k0 = kb << 6;
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = 0.f; s[j][1] = 0.f; s[j][2] = 0.f; s[j][3] = 0.f; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int krow = 8 * j + (lane & 7);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t kf[4];
        ldmatrix_x4(kf, sK + sw_off(krow, (lane >> 3) + 4 * half));
        mma_bf16_16816(s[j], qa[2 * half + 0], kf[0], kf[1]);
        mma_bf16_16816(s[j], qa[2 * half + 1], kf[2], kf[3]);
      }
    }
    if (k0 + 64 > Tk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = k0 + 8 * j + 2 * tq;
        if (key >= Tk) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (key + 1 >= Tk) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], mneg[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float m_new = fmaxf(m_run[r], mx[r]);
      corr[r] = fast_exp2((m_run[r] - m_new) * kScaleLog2);
      m_run[r] = m_new;
      mneg[r] = m_new * kScaleLog2;
      l_run[r] *= corr[r];
    }
    float ls[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = fast_exp2(s[j][0] * kScaleLog2 - mneg[0]);
      s[j][1] = fast_exp2(s[j][1] * kScaleLog2 - mneg[0]);
      s[j][2] = fast_exp2(s[j][2] * kScaleLog2 - mneg[1]);
      s[j][3] = fast_exp2(s[j][3] * kScaleLog2 - mneg[1]);
      ls[0] += s[j][0] + s[j][1];
      ls[1] += s[j][2] + s[j][3];
    }
    l_run[0] += ls[0];
    l_run[1] += ls[1];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= corr[0]; o[j][1] *= corr[0];
      o[j][2] *= corr[1]; o[j][3] *= corr[1];
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * ks][0], s[2 * ks][1]);
      pa[1] = pack_bf16x2(s[2 * ks][2], s[2 * ks][3]);
      pa[2] = pack_bf16x2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      pa[3] = pack_bf16x2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
      const int m = lane >> 3;
      const int vrow = 16 * ks + (m & 1) * 8 + (lane & 7);
#pragma unroll
      for (int jd = 0; jd < 8; jd += 2) {
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, sV + sw_off(vrow, jd + (m >> 1)));
        mma_bf16_16816(o[jd], pa, vf[0], vf[1]);
        mma_bf16_16816(o[jd + 1], pa, vf[2], vf[3]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
  const int r0 = qbase + q0 + g, r1 = r0 + 8;
  uint16_t* obase = out + static_cast<int64_t>(win) * t_live * width + head * kHeadDim + 2 * tq;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (r0 < Tq)
      *reinterpret_cast<uint32_t*>(obase + static_cast<int64_t>(r0) * width + 8 * j) = pack16x2(o[j][0] * inv0, o[j][1] * inv0, out_fp16);
    if (r1 < Tq)
      *reinterpret_cast<uint32_t*>(obase + static_cast<int64_t>(r1) * width + 8 * j) = pack16x2(o[j][2] * inv1, o[j][3] * inv1, out_fp16);
  }
}

}  // namespace

const char* attention_h64_long(cudaStream_t stream, const __nv_bfloat16* qkv, const __nv_bfloat16* const_kv, int n_const,
                               int n_win, int t_live, int heads, void* out, int out_fp16) {
  if (n_win <= 0 || t_live <= 0 || heads <= 0) return "attention: empty problem";
  if (n_const < 0 || (n_const > 0 && const_kv == nullptr)) return "attention: constant keys missing";
  const int q_chunks = (t_live + kLongQ - 1) / kLongQ;
  const int64_t blocks = static_cast<int64_t>(n_win) * heads * q_chunks;
  if (blocks > 0x7fffffff) return "attention: too many (window, head, query chunk) items";
  {
    const double tk = t_live + n_const, width = 64.0 * heads;
    LaunchScope scope(stream, "attention", 4.0 * n_win * heads * t_live * tk * 64.0, 2.0 * n_win * t_live * 4.0 * width);
    attention_h64_long_kernel<<<static_cast<unsigned>(blocks), kAttnThreads, 0, stream>>>(
        qkv, const_kv, n_const, t_live, q_chunks, heads, static_cast<uint16_t*>(out), out_fp16);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace cebc
