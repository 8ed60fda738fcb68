// Persistent, warp-specialised tcgen05 attention: softmax(q k^T / 8) v per (window, head), 64-dim heads, <= 256 queries
// and <= 256 keys per item (229 for a 224x224 window with 32 prompt tokens).
// Replaces nn.MultiheadAttention -> F.scaled_dot_product_attention (/root/reference/models/clip/_clip/blocks.py:25,35-37);
// deep-VPT constant prompt keys/values are appended as in attention.cu (reference models/clip/model.py:164-183).
//
// One CTA per SM loops over (window, head) items; all stages of consecutive items overlap:
//   warp 0       TMA producer: Q [256 x 64], K, V [256 x 64] (constant prompt rows first) of item i+1 are fetched into
//                the second smem stage while item i is being processed
//   warps 1, 3   MMA issuers, one per 128-query tile t (so the two tiles run decoupled and their softmax phases stagger):
//                S_t = Q_t K^T (128 x 256 x 64), later O_t = P_t V with P_t read straight from TMEM (tcgen05.mma
//                A-operand in tensor memory) and V consumed in place as MN-major B
//   warp 2       TMEM allocator (all 512 columns: S_0 | S_1, each 128 lanes x 256 fp32)
//   warps 4-7    softmax group of tile 0, warps 8-11 of tile 1: thread = query row = TMEM lane; row max, exp2, row sum from
//                TMEM; P is written back to TMEM as packed bf16 over the consumed S columns (no smem round trip);
//                finally O_t * 1/rowsum -> bf16 -> global
// Scores and probabilities never leave the SM: HBM traffic is Q, K, V in and O out.
#include "common.cuh"
#include "kernels.h"

namespace cebc {

namespace {

constexpr int kThreadsF = 384;
constexpr int kQTileBytes = 128 * 128;   // one 128-query tile
constexpr int kQBytes = 2 * kQTileBytes; // 256 query rows x 64 dims bf16
constexpr int kKVBytes = 256 * 128;      // 256 key rows x 64 dims bf16
constexpr int kStageBytesF = kQBytes + 2 * kKVBytes;  // 96 KB
constexpr int kSmemF = 2 * kStageBytesF + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kQkvLdF = 3 * 768;

__device__ __forceinline__ uint64_t desc_sw128_mn(uint32_t smem_addr_bytes) {
  // MN-major operand in 128B-swizzled rows: ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units; SBO = 1024 B between
  // 8-key groups; LBO (distance between 64-element N blocks) is unused for N = 64.
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr_bytes & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(kKVBytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_f(int M, int N, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]: A (bf16 pairs packed per 32-bit column) is read from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// max over the real keys of one 32-key chunk (`lim` = number of real keys left; only the last chunk is partial)
__device__ __forceinline__ void chunk_max(const uint32_t (&v)[32], int lim, float& mx) {
  if (lim >= 32) {
#pragma unroll
    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < lim) mx = fmaxf(mx, __uint_as_float(v[j]));
  }
}
// p_j = exp2(min(s_j * scale - m_scaled, 120)) (0 beyond the real keys), row sum in fp32, P as packed bf16 pairs into TMEM.
// The P columns [16c, 16c+16) overlay S columns that have already been consumed (16c + 16 <= 32c + 32).
__device__ __forceinline__ void chunk_exp_store(const uint32_t (&v)[32], int lim, float scale, float m_scaled,
                                                float& row_sum, uint32_t p_taddr) {
  uint32_t pk[16];
  if (lim >= 32) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float p0 = ex2f(fminf(__uint_as_float(v[2 * j]) * scale - m_scaled, 120.f));
      const float p1 = ex2f(fminf(__uint_as_float(v[2 * j + 1]) * scale - m_scaled, 120.f));
      row_sum += p0 + p1;
      pk[j] = pack_bf16x2(p0, p1);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float p0 = (2 * j < lim) ? ex2f(fminf(__uint_as_float(v[2 * j]) * scale - m_scaled, 120.f)) : 0.f;
      const float p1 = (2 * j + 1 < lim) ? ex2f(fminf(__uint_as_float(v[2 * j + 1]) * scale - m_scaled, 120.f)) : 0.f;
      row_sum += p0 + p1;
      pk[j] = pack_bf16x2(p0, p1);
    }
  }
  tmem_st_32x32b_x16(p_taddr, pk);
}

__global__ void __launch_bounds__(kThreadsF, 1)
attention_fa_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                    const __grid_constant__ CUtensorMap tm_const, int n_const, int t_live, int n_items,
                    uint16_t* __restrict__ out, int out_fp16) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kStageBytesF);
  uint64_t* qk_full = bars + 0;     // [2] TMA -> MMA
  uint64_t* v_full = bars + 2;      // [2]
  uint64_t* qk_empty = bars + 4;    // [2] MMA (commit) -> TMA
  uint64_t* v_empty = bars + 6;     // [2]
  uint64_t* s_full = bars + 8;      // [2 tiles] MMA (commit) -> softmax group
  uint64_t* p_ready = bars + 10;    // [2 tiles] softmax group (4 warps) -> MMA
  uint64_t* o_full = bars + 12;     // [2 tiles] MMA (commit) -> softmax group
  uint64_t* tmem_free = bars + 14;  // [2 tiles] softmax group (4 warps) -> MMA
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Tk = n_const + t_live;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_const);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qk_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&qk_empty[i], 2); mbar_init(&v_empty[i], 2);
      mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&o_full[i], 1); mbar_init(&tmem_free[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      const int head = item % 12, win = item / 12;
      const int row_base = win * t_live;
      uint8_t* sQ = smem + s * kStageBytesF;
      uint8_t* sK = sQ + kQBytes;
      uint8_t* sV = sK + kKVBytes;
      mbar_wait(&qk_empty[s], ph ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&qk_full[s], kQBytes + kKVBytes);
        tma_load_2d(sQ, &tm_q, &qk_full[s], head * 64, row_base);
        if (n_const > 0) tma_load_2d(sK, &tm_const, &qk_full[s], 768 + head * 64, 0);
        tma_load_2d(sK + n_const * 128, &tm_kv, &qk_full[s], 768 + head * 64, row_base);
      }
      __syncwarp();
      mbar_wait(&v_empty[s], ph ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&v_full[s], kKVBytes);
        if (n_const > 0) tma_load_2d(sV, &tm_const, &v_full[s], 1536 + head * 64, 0);
        tma_load_2d(sV + n_const * 128, &tm_kv, &v_full[s], 1536 + head * 64, row_base);
      }
      __syncwarp();
    }
  } else if (warp == 1 || warp == 3) {
    // ------------------------------------ MMA issuer of query tile t ------------------------------------
    const int t = warp >> 1;  // warp 1 -> tile 0, warp 3 -> tile 1
    constexpr uint32_t idesc_s = idesc_f(128, 256, false);
    constexpr uint32_t idesc_o = idesc_f(128, 64, true);
    const int k_steps = (Tk + 15) >> 4;
    const uint32_t s_tmem = tmem_base + t * 256;  // S_t; packed bf16 P_t overlays columns [0, 128)
    const uint32_t o_tmem = s_tmem + 128;         // O_t overlays S_t columns [128, 192)
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1, ip = it & 1;
      const uint32_t q_addr = smem_u32(smem + s * kStageBytesF) + t * kQTileBytes;
      const uint32_t k_addr = smem_u32(smem + s * kStageBytesF) + kQBytes;
      const uint32_t v_addr = k_addr + kKVBytes;
      mbar_wait(&qk_full[s], ph);
      mbar_wait(&tmem_free[t], ip ^ 1);  // O_t of the previous item has been read out
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(s_tmem, umma_desc_sw128_kmajor(q_addr + k * 32), umma_desc_sw128_kmajor(k_addr + k * 32), idesc_s,
                       k != 0 ? 1u : 0u);
        umma_commit(&s_full[t]);
        umma_commit(&qk_empty[s]);  // Q/K of this stage are dead once both tiles' S MMAs have retired (2 arrivals)
      }
      __syncwarp();
      mbar_wait(&v_full[s], ph);
      mbar_wait(&p_ready[t], ip);
      tc_fence_after();
      if (lane == 0) {
        for (int ks = 0; ks < k_steps; ++ks)
          umma_bf16_ts(o_tmem, s_tmem + ks * 8, desc_sw128_mn(v_addr + ks * 2048), idesc_o, ks != 0 ? 1u : 0u);
        umma_commit(&o_full[t]);
        umma_commit(&v_empty[s]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------ softmax groups ------------------------------------
    const int t = (warp - 4) >> 2;            // query tile of this group
    const int q = warp & 3;                   // TMEM lane quadrant
    const int r = q * 32 + lane;              // row inside the tile
    const int q_row = t * 128 + r;            // row inside the window
    const bool warp_active = t * 128 + q * 32 < t_live;
    const int n_chunks = (Tk + 31) >> 5;
    const float kScale = 0.125f * 1.4426950408889634f;
    const uint32_t s_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * 256;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int ip = it & 1;
      const int head = item % 12, win = item / 12;
      float row_sum = 0.f;
      mbar_wait(&s_full[t], ip);
      tc_fence_after();
      if (warp_active) {
        // Single pass over S (TMEM reads are the scarce resource: ~64 B/clk/SM). Softmax is shift invariant, so the
        // reference maximum only has to keep exp2 in range: the max over the first 32 keys is used and the exponent is
        // clamped at +120 (a row whose other scores exceed that reference by > 660 would saturate instead of overflow).
        uint32_t va[32], vb[32];
        float mx = -INFINITY;
        tmem_ld_32x32b_x32(s_row, va);
        tmem_ld_wait();
        chunk_max(va, Tk, mx);
        // pass 2: p = exp2((s - max) / 8 * log2 e), row sum, packed bf16 P back into TMEM
        const float m_scaled = mx * kScale;
#pragma unroll 1
        for (int c = 0; c < n_chunks; c += 2) {
          if (c + 1 < n_chunks) tmem_ld_32x32b_x32(s_row + (c + 1) * 32, vb);
          chunk_exp_store(va, Tk - c * 32, kScale, m_scaled, row_sum, s_row + c * 16);
          tmem_ld_wait();
          if (c + 1 < n_chunks) {
            if (c + 2 < n_chunks) tmem_ld_32x32b_x32(s_row + (c + 2) * 32, va);
            chunk_exp_store(vb, Tk - (c + 1) * 32, kScale, m_scaled, row_sum, s_row + (c + 1) * 16);
            tmem_ld_wait();
          }
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[t]);

      mbar_wait(&o_full[t], ip);
      tc_fence_after();
      if (warp_active) {
        uint32_t o0[32], o1[32];
        tmem_ld_32x32b_x32(s_row + 128, o0);
        tmem_ld_32x32b_x32(s_row + 160, o1);
        tmem_ld_wait();
        // O is in registers: hand the TMEM tile back before the global stores so the next S MMA can start
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_free[t]);
        if (q_row < t_live) {
          const float inv = 1.0f / row_sum;
          uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(win) * t_live + q_row) * 768 + head * 64);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            dst[g] = make_uint4(pack16x2(__uint_as_float(o0[8 * g + 0]) * inv, __uint_as_float(o0[8 * g + 1]) * inv, out_fp16),
                                pack16x2(__uint_as_float(o0[8 * g + 2]) * inv, __uint_as_float(o0[8 * g + 3]) * inv, out_fp16),
                                pack16x2(__uint_as_float(o0[8 * g + 4]) * inv, __uint_as_float(o0[8 * g + 5]) * inv, out_fp16),
                                pack16x2(__uint_as_float(o0[8 * g + 6]) * inv, __uint_as_float(o0[8 * g + 7]) * inv, out_fp16));
#pragma unroll
          for (int g = 0; g < 4; ++g)
            dst[4 + g] =
                make_uint4(pack16x2(__uint_as_float(o1[8 * g + 0]) * inv, __uint_as_float(o1[8 * g + 1]) * inv, out_fp16),
                           pack16x2(__uint_as_float(o1[8 * g + 2]) * inv, __uint_as_float(o1[8 * g + 3]) * inv, out_fp16),
                           pack16x2(__uint_as_float(o1[8 * g + 4]) * inv, __uint_as_float(o1[8 * g + 5]) * inv, out_fp16),
                           pack16x2(__uint_as_float(o1[8 * g + 6]) * inv, __uint_as_float(o1[8 * g + 7]) * inv, out_fp16));
        }
      }
      if (!warp_active) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_free[t]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

typedef CUresult (*PFN_encodeTiledF)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_tmap_rows_f(CUtensorMap* map, const void* base, int64_t rows, int box_rows) {
  static PFN_encodeTiledF enc = nullptr;
  if (!enc) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !ptr)
      return false;
    enc = reinterpret_cast<PFN_encodeTiledF>(ptr);
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(kQkvLdF), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(kQkvLdF) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// Same contract as attention_h64 (kernels.h); additionally requires n_const to be a multiple of 8 (swizzle atom).
const char* attention_h64_fa(cudaStream_t stream, const __nv_bfloat16* qkv, const __nv_bfloat16* const_kv, int n_const,
                             int n_win, int t_live, void* out, int out_fp16) {
  if (n_win <= 0 || t_live <= 0) return "attention: empty problem";
  if (n_const < 0 || (n_const > 0 && const_kv == nullptr)) return "attention: constant keys missing";
  if (t_live + n_const > 256) return "attention: sequence longer than 256 keys is not supported";
  if (n_const % 8 != 0) return "attention(fa): constant key count must be a multiple of 8";
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_fa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemF);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    attr_set = true;
  }
  const int64_t rows = static_cast<int64_t>(n_win) * t_live;
  CUtensorMap tq, tkv, tc;
  if (!make_tmap_rows_f(&tq, qkv, rows, 256)) return "attention: cuTensorMapEncodeTiled(q) failed";
  if (!make_tmap_rows_f(&tkv, qkv, rows, 256 - n_const)) return "attention: cuTensorMapEncodeTiled(kv) failed";
  if (n_const > 0) {
    if (!make_tmap_rows_f(&tc, const_kv, n_const, n_const)) return "attention: cuTensorMapEncodeTiled(const) failed";
  } else {
    tc = tkv;
  }
  const int n_items = n_win * 12;
  const int grid = n_items < device_num_sms() ? n_items : device_num_sms();
  {
    const double tk = t_live + n_const;
    LaunchScope scope(stream, "attention", 4.0 * n_win * 12.0 * t_live * tk * 64.0,
                      2.0 * n_win * t_live * (2304.0 + 768.0));
    attention_fa_kernel<<<grid, kThreadsF, kSmemF, stream>>>(tq, tkv, tc, n_const, t_live, n_items,
                                                             static_cast<uint16_t*>(out), out_fp16);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace cebc
