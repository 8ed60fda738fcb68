// tcgen05 / TMEM short-sequence attention: softmax(q k^T / 8) v per (window, head), 64-dim heads, <= 256 keys.
// Replaces nn.MultiheadAttention -> F.scaled_dot_product_attention (/root/reference/models/clip/_clip/blocks.py:25,35-37)
// with the deep-VPT constant prompt keys/values appended (reference models/clip/model.py:164-183, see attention.cu).
//
// One CTA per (window, head, 128-query tile), 128 threads, two CTAs resident per SM (96 KB smem, 256 TMEM columns each)
// so one CTA's softmax overlaps the other's loads and MMAs:
//   thread 0      TMA: Q tile [128 x 64], K and V tiles [256 x 64] (constant prompt rows first, then the window's rows;
//                 key order is irrelevant to softmax), all 128B-swizzled; then tcgen05.mma S = Q K^T (128 x 256 x 64)
//   all 4 warps   thread = query row = TMEM lane: row max, exp2, row sum straight from TMEM; P (bf16) is written into
//                 smem in the K-major UMMA layout, over the dead Q/K tiles
//   thread 0      tcgen05.mma O = P V (128 x 64 x keys), V consumed in place as an MN-major operand (no transpose);
//                 O overlays the first 64 columns of S in TMEM
//   all 4 warps   O * 1/rowsum -> bf16 -> global (one 128 B row segment per thread)
// Scores never leave TMEM / registers; the only HBM traffic is Q, K, V in and O out.
#include "common.cuh"
#include "kernels.h"

namespace cebc {

namespace {

constexpr int kThreadsA = 128;
constexpr int kQBytes = 128 * 128;        // 128 query rows x 64 dims bf16
constexpr int kKVBytes = 256 * 128;       // 256 key rows x 64 dims bf16
constexpr int kPBytes = 4 * 128 * 128;    // 4 key blocks of 64 keys: [128 rows x 128 B] each
constexpr int kOffQ = 0;
constexpr int kOffK = kQBytes;
constexpr int kOffV = kOffK + kKVBytes + 16384;  // Q | K | 16 KB pad | V   (P overlays Q | K | pad)
constexpr int kSmemA = kOffV + kKVBytes + 1024 /*align*/ + 64 /*barriers*/;
constexpr int kTmemColsA = 256;
constexpr int kQkvLd = 3 * 768;

// MN-major operand (V as B[N = dims][K = keys], stored [key][dim]) in 128B-swizzled rows: canonical layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -- 8 keys x 128 B per swizzle atom, SBO = 1024 B between 8-key
// groups, LBO = distance between 64-element blocks along N (only one block here).
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr_bytes & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(kKVBytes >> 4) << 16;  // LBO (unused: N = 64 is a single block)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;      // SBO
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;              // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreadsA, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                    const __grid_constant__ CUtensorMap tm_const, int n_const, int t_live, int q_tiles,
                    uint16_t* __restrict__ out, int out_fp16) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint8_t* sP = smem;  // overlays Q | K | pad once S has been computed
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffV + kKVBytes);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_o = bars + 3;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int item = blockIdx.x;
  const int qt = item % q_tiles; item /= q_tiles;
  const int head = item % 12;
  const int win = item / 12;
  const int Tk = n_const + t_live;
  const int row_base = win * t_live;  // first row of this window in qkv / out

  if (tid == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemColsA>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (tid == 0) {
    // ---- loads: constant prompt rows occupy key slots [0, n_const), the window's rows follow ----
    mbar_arrive_expect_tx(bar_qk, kQBytes + kKVBytes);
    tma_load_2d(sQ, &tm_q, bar_qk, head * 64, row_base + qt * 128);
    if (n_const > 0) tma_load_2d(sK, &tm_const, bar_qk, 768 + head * 64, 0);
    tma_load_2d(sK + n_const * 128, &tm_kv, bar_qk, 768 + head * 64, row_base);
    mbar_arrive_expect_tx(bar_v, kKVBytes);
    if (n_const > 0) tma_load_2d(sV, &tm_const, bar_v, 1536 + head * 64, 0);
    tma_load_2d(sV + n_const * 128, &tm_kv, bar_v, 1536 + head * 64, row_base);
    // ---- S = Q K^T : 128 x 256, K = 64 ----
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    constexpr uint32_t idesc_s = idesc_bf16(128, 256, false);
    const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16_ss(tmem_base, umma_desc_sw128_kmajor(q_addr + k * 32), umma_desc_sw128_kmajor(k_addr + k * 32), idesc_s,
                   k != 0 ? 1u : 0u);
    umma_commit(bar_s);
  }

  // ---- softmax: thread = query row (TMEM lane 32 * warp + lane) ----
  const int r = tid;                       // row inside the tile
  const int q_row = qt * 128 + r;          // row inside the window
  const int rows_valid = t_live - qt * 128;  // rows of this tile that exist
  const bool warp_active = warp * 32 < rows_valid;
  const int n_chunks = (Tk + 31) >> 5;     // 32-key chunks that contain at least one real key
  const float kScale = 0.125f * 1.4426950408889634f;
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  float row_sum = 0.f;

  mbar_wait(bar_s, 0);
  tc_fence_after();
  if (warp_active) {
    float mx = -INFINITY;
    for (int c = 0; c < n_chunks; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + c * 32, v);
      tmem_ld_wait();
      const int lim = Tk - c * 32;  // keys of this chunk that are real
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < lim) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    const float m_scaled = mx * kScale;
    for (int c = 0; c < n_chunks; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + c * 32, v);
      tmem_ld_wait();
      const int lim = Tk - c * 32;
      float p[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        p[j] = (j < lim) ? ex2(__uint_as_float(v[j]) * kScale - m_scaled) : 0.f;
        row_sum += p[j];
      }
      // P[r, 32c .. 32c+31] -> key block c/2, 16-byte pieces (c%2)*4 .. +3 of the row, 128B swizzle
      uint8_t* dst = sP + (c >> 1) * 16384 + r * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int piece = (c & 1) * 4 + g;
        *reinterpret_cast<uint4*>(dst + ((piece ^ (r & 7)) << 4)) =
            make_uint4(pack_bf16x2(p[8 * g + 0], p[8 * g + 1]), pack_bf16x2(p[8 * g + 2], p[8 * g + 3]),
                       pack_bf16x2(p[8 * g + 4], p[8 * g + 5]), pack_bf16x2(p[8 * g + 6], p[8 * g + 7]));
      }
    }
  } else {
    // rows beyond the window: P = 0 keeps the (discarded) O rows finite
    for (int c = 0; c < n_chunks; ++c) {
      uint8_t* dst = sP + (c >> 1) * 16384 + r * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(dst + ((((c & 1) * 4 + g) ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  // an odd number of chunks leaves half a key block unwritten: zero it (the PV loop runs over whole 16-key steps)
  if (n_chunks & 1) {
    uint8_t* dst = sP + (n_chunks >> 1) * 16384 + r * 128;
#pragma unroll
    for (int g = 4; g < 8; ++g) *reinterpret_cast<uint4*>(dst + ((g ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();  // generic-proxy writes of P -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();           // all of P written, all reads of S retired (O overlays S)
  tc_fence_after();

  if (tid == 0) {
    // ---- O = P V : 128 x 64, K = keys (whole 16-key steps) ----
    mbar_wait(bar_v, 0);
    tc_fence_after();
    constexpr uint32_t idesc_o = idesc_bf16(128, 64, true);
    const uint32_t p_addr = smem_u32(sP), v_addr = smem_u32(sV);
    const int k_steps = (Tk + 15) >> 4;
    for (int ks = 0; ks < k_steps; ++ks)
      umma_bf16_ss(tmem_base, umma_desc_sw128_kmajor(p_addr + (ks >> 2) * 16384 + (ks & 3) * 32),
                   umma_desc_sw128_mnmajor(v_addr + ks * 2048), idesc_o, ks != 0 ? 1u : 0u);
    umma_commit(bar_o);
  }

  mbar_wait(bar_o, 0);
  tc_fence_after();
  if (warp_active) {
    uint32_t o0[32], o1[32];
    tmem_ld_32x32b_x32(t_row, o0);
    tmem_ld_32x32b_x32(t_row + 32, o1);
    tmem_ld_wait();
    if (q_row < t_live) {
      const float inv = 1.0f / row_sum;
      uint4* dst = reinterpret_cast<uint4*>(out + static_cast<int64_t>(row_base + q_row) * 768 + head * 64);
#pragma unroll
      for (int g = 0; g < 4; ++g)
        dst[g] = make_uint4(pack16x2(__uint_as_float(o0[8 * g + 0]) * inv, __uint_as_float(o0[8 * g + 1]) * inv, out_fp16),
                            pack16x2(__uint_as_float(o0[8 * g + 2]) * inv, __uint_as_float(o0[8 * g + 3]) * inv, out_fp16),
                            pack16x2(__uint_as_float(o0[8 * g + 4]) * inv, __uint_as_float(o0[8 * g + 5]) * inv, out_fp16),
                            pack16x2(__uint_as_float(o0[8 * g + 6]) * inv, __uint_as_float(o0[8 * g + 7]) * inv, out_fp16));
#pragma unroll
      for (int g = 0; g < 4; ++g)
        dst[4 + g] = make_uint4(pack16x2(__uint_as_float(o1[8 * g + 0]) * inv, __uint_as_float(o1[8 * g + 1]) * inv, out_fp16),
                                pack16x2(__uint_as_float(o1[8 * g + 2]) * inv, __uint_as_float(o1[8 * g + 3]) * inv, out_fp16),
                                pack16x2(__uint_as_float(o1[8 * g + 4]) * inv, __uint_as_float(o1[8 * g + 5]) * inv, out_fp16),
                                pack16x2(__uint_as_float(o1[8 * g + 6]) * inv, __uint_as_float(o1[8 * g + 7]) * inv, out_fp16));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemColsA>(tmem_base);
  }
}

typedef CUresult (*PFN_encodeTiledA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_tmap_rows(CUtensorMap* map, const void* base, int64_t rows, int box_rows) {
  static PFN_encodeTiledA enc = nullptr;
  if (!enc) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !ptr)
      return false;
    enc = reinterpret_cast<PFN_encodeTiledA>(ptr);
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(kQkvLd), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(kQkvLd) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// Same contract as attention_h64 (kernels.h); additionally requires n_const to be a multiple of 8 (swizzle atom).
const char* attention_h64_tc(cudaStream_t stream, const __nv_bfloat16* qkv, const __nv_bfloat16* const_kv, int n_const,
                             int n_win, int t_live, void* out, int out_fp16) {
  if (n_win <= 0 || t_live <= 0) return "attention: empty problem";
  if (n_const < 0 || (n_const > 0 && const_kv == nullptr)) return "attention: constant keys missing";
  if (t_live + n_const > 256) return "attention: sequence longer than 256 keys is not supported";
  if (n_const % 8 != 0) return "attention(tc): constant key count must be a multiple of 8";
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemA);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    attr_set = true;
  }
  const int64_t rows = static_cast<int64_t>(n_win) * t_live;
  CUtensorMap tq, tkv, tc;
  if (!make_tmap_rows(&tq, qkv, rows, 128)) return "attention: cuTensorMapEncodeTiled(q) failed";
  if (!make_tmap_rows(&tkv, qkv, rows, 256 - n_const)) return "attention: cuTensorMapEncodeTiled(kv) failed";
  if (n_const > 0) {
    if (!make_tmap_rows(&tc, const_kv, n_const, n_const)) return "attention: cuTensorMapEncodeTiled(const) failed";
  } else {
    tc = tkv;
  }
  const int q_tiles = (t_live + 127) / 128;
  {
    const double tk = t_live + n_const;
    LaunchScope scope(stream, "attention", 4.0 * n_win * 12.0 * t_live * tk * 64.0,
                      2.0 * n_win * t_live * (2304.0 + 768.0));
    attention_tc_kernel<<<n_win * 12 * q_tiles, kThreadsA, kSmemA, stream>>>(tq, tkv, tc, n_const, t_live, q_tiles,
                                                                             static_cast<uint16_t*>(out), out_fp16);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace cebc
