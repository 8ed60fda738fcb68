// Persistent warp-specialised tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T  (bf16 x bf16 -> fp32 in TMEM)
// with the CLIP-EBC epilogues fused (bias, QuickGELU, fp32 residual stream add, BN-folded bias + ReLU for the
// decoder convs, hi/lo bf16 split for the fp32-accurate projection).
//
// Replaces, on the reference's hot path (all via PyTorch library kernels there):
//   nn.MultiheadAttention in_proj / out_proj      /root/reference/models/clip/_clip/blocks.py:25,37,40
//   mlp.c_fc + QuickGELU + mlp.c_proj             /root/reference/models/clip/_clip/blocks.py:27-31,41
//   image_encoder.conv1 (patchify, as a GEMM)     /root/reference/models/clip/model.py:147
//   image_decoder BasicBlock conv3x3+BN(+ReLU)    /root/reference/models/utils.py:290-303   (implicit GEMM: 9 shifted K-segments)
//   projection 1x1 conv                           /root/reference/models/clip/model.py:198
//
// Structure (one CTA per SM, 256 threads):
//   warp 0   TMA producer: A/B k-blocks (128B-swizzled, K-major) into a STAGES-deep smem ring
//   warp 1   MMA issuer: one lane issues tcgen05.mma 128xBLOCK_Nx16, accumulators double-buffered in TMEM
//   warp 2   TMEM allocator / deallocator
//   warps 4-7 epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue -> global
// The A operand may be a sequence of K-segments, each with its own (row shift, column start) in the A tensor: this is
// how the 3x3 convolution over a zero-bordered NHWC grid (segment = filter tap, row shift = dy*Wp+dx) and the
// [hi|lo|hi] x [Whi|Whi|Wlo] split-precision projection are expressed without materialising im2col buffers.
#include "common.cuh"
#include "kernels.h"

namespace cebc {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 256;

constexpr int kStageRowBytes = 144;                       // 128 B payload + 16 B pad: conflict-free 16 B accesses
constexpr int kStagingPerWarp = 32 * kStageRowBytes;      // epilogue transpose buffer of one warp (32 rows)
constexpr int kStagingBytes = 4 * kStagingPerWarp;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N > 192) ? 4 : (BLOCK_N > 128 ? 5 : 6);
  static constexpr int kTmemCols = (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float quick_gelu(float x) {
  // x * sigmoid(1.702 x)   (reference: blocks.py:17-19)
  return __fdividef(x, 1.0f + __expf(-1.702f * x));
}

template <int EPI>
constexpr bool epi_out_is_bf16() {
  return EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_RELU_MASK_BF16;
}

// ---- epilogues with bf16 output: 64 accumulator columns per pass --------------------------------------------------
// Phase 1 (thread = output row, as tcgen05.ld delivers it): bias + activation, pack to bf16, write the row's 128 B into
// the warp's staging tile. Phase 2 (after the transpose through smem): every warp store instruction writes 4 complete
// 128 B row segments, so global stores are fully coalesced instead of 32 scattered 16 B pieces.
template <int EPI>
__device__ __forceinline__ void epilogue_bf16_chunk64(const GemmParams& p, uint8_t* stg, int lane, int row0, int n,
                                                      const uint32_t (&r0)[32], const uint32_t (&r1)[32]) {
  const int row = row0 + lane;
  bool border = false;
  if constexpr (EPI == EPI_BIAS_RELU_MASK_BF16) {
    // zero-bordered grid: rows are (image, py, px) over a (mask_hp x mask_wp) padded grid; border rows must stay 0
    const int rpi = p.mask_hp * p.mask_wp;
    const int q = row % rpi;
    const int py = q / p.mask_wp, px = q - py * p.mask_wp;
    border = (py == p.mask_hp - 1) || (px == p.mask_wp - 1) || (p.mask_lead && (py == 0 || px == 0));
  }
  uint4* my = reinterpret_cast<uint4*>(stg + lane * kStageRowBytes);
  const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
  for (int j = 0; j < 8; ++j) {  // 8 columns per iteration
    const uint32_t* src = (j < 4) ? &r0[8 * j] : &r1[8 * (j - 4)];
    const float4 ba = __ldg(b4 + 2 * j), bb = __ldg(b4 + 2 * j + 1);
    float v[8];
    v[0] = __uint_as_float(src[0]) + ba.x; v[1] = __uint_as_float(src[1]) + ba.y;
    v[2] = __uint_as_float(src[2]) + ba.z; v[3] = __uint_as_float(src[3]) + ba.w;
    v[4] = __uint_as_float(src[4]) + bb.x; v[5] = __uint_as_float(src[5]) + bb.y;
    v[6] = __uint_as_float(src[6]) + bb.z; v[7] = __uint_as_float(src[7]) + bb.w;
    if constexpr (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = quick_gelu(v[k]);
    }
    if constexpr (EPI == EPI_BIAS_RELU_MASK_BF16) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = border ? 0.0f : fmaxf(v[k], 0.0f);
    }
    my[j] = make_uint4(pack16x2(v[0], v[1], p.out_fp16), pack16x2(v[2], v[3], p.out_fp16), pack16x2(v[4], v[5], p.out_fp16),
                   pack16x2(v[6], v[7], p.out_fp16));
  }
  __syncwarp();
  const int piece = lane & 7, rsub = lane >> 3;
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + rsub;
    const uint4 u = *reinterpret_cast<const uint4*>(stg + rr * kStageRowBytes + piece * 16);
    if (row0 + rr < p.M)
      *reinterpret_cast<uint4*>(out + static_cast<size_t>(row0 + rr) * p.ldo + n + piece * 8) = u;
  }
  __syncwarp();
}

// ---- epilogues that finish in fp32 (residual stream, projection, conv2 hi/lo split): 32 columns per pass ----------
// The raw accumulators are transposed through smem first; bias, residual add and the store all happen in the
// coalesced layout (lane -> 4 consecutive columns of one row, a warp instruction covers 4 rows x 128 B).
// The residual of the NEXT chunk is requested before the current chunk is processed (the output may alias the residual
// -- the in-place residual stream -- so the compiler cannot hoist those loads itself).
template <int EPI>
constexpr bool epi_has_resid() { return EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_RESID_RELU_SPLIT; }

template <int EPI>
__device__ __forceinline__ void load_resid_chunk32(const GemmParams& p, int lane, int row0, int n, float4 (&x)[8]) {
  if constexpr (epi_has_resid<EPI>()) {
    const int c4 = (lane & 7) * 4, rsub = lane >> 3;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int grow = row0 + it * 4 + rsub;
      x[it] = (grow < p.M) ? *reinterpret_cast<const float4*>(p.resid + static_cast<size_t>(grow) * p.ldr + n + c4)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_f32_chunk32(const GemmParams& p, uint8_t* stg, int lane, int row0, int n,
                                                     const uint32_t (&r)[32], const float4 (&x)[8]) {
  uint4* my = reinterpret_cast<uint4*>(stg + lane * kStageRowBytes);
#pragma unroll
  for (int j = 0; j < 8; ++j) my[j] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
  __syncwarp();
  const int c4 = (lane & 7) * 4, rsub = lane >> 3;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (EPI != EPI_F32 || p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + n + c4));
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + rsub;
    const int grow = row0 + rr;
    float4 v = *reinterpret_cast<const float4*>(stg + rr * kStageRowBytes + c4 * 4);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    if constexpr (epi_has_resid<EPI>()) { v.x += x[it].x; v.y += x[it].y; v.z += x[it].z; v.w += x[it].w; }
    if (grow < p.M) {
      if constexpr (EPI == EPI_BIAS_RESID_RELU_SPLIT) {
        // relu, then split into hi + lo bf16 so the next GEMM can recover ~fp32 accuracy: x ~= hi + lo
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        const float hx = round16(v.x, p.out_fp16), hy = round16(v.y, p.out_fp16);
        const float hz = round16(v.z, p.out_fp16), hw = round16(v.w, p.out_fp16);
        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(grow) * p.ldo + n + c4;
        *reinterpret_cast<uint2*>(o) = make_uint2(pack16x2(hx, hy, p.out_fp16), pack16x2(hz, hw, p.out_fp16));
        *reinterpret_cast<uint2*>(o + p.N) =
            make_uint2(pack16x2(v.x - hx, v.y - hy, p.out_fp16), pack16x2(v.z - hz, v.w - hw, p.out_fp16));
      } else {
        *reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<size_t>(grow) * p.ldo + n + c4) = v;
      }
    }
  }
  __syncwarp();
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + STAGES * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tmem_full_bar = bars + 2 * STAGES;  // [2]
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
  const int n_tiles = p.N / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.K / kBlockK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    uint32_t stage = 0, phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t / n_tiles, n_blk = t - m_blk * n_tiles;
      const int m0 = m_blk * kBlockM, n0 = n_blk * BLOCK_N;
      int seg = 0, kk = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(sa, &tma_a, &full_bar[stage], p.seg_col_start[seg] + kk * kBlockK, m0 + p.seg_row_shift[seg]);
          tma_load_2d(sb, &tma_b, &full_bar[stage], kb * kBlockK, n0);
        }
        __syncwarp();
        if (++kk == p.seg_kblocks) { kk = 0; ++seg; }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    const uint32_t idesc = umma_idesc_bf16_f32(kBlockM, BLOCK_N, p.ab_fp16);
    uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BLOCK_N;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t da = umma_desc_sw128_kmajor(a_addr + k * kUmmaK * 2);
            const uint64_t db = umma_desc_sw128_kmajor(b_addr + k * kUmmaK * 2);
            umma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                          // frees the smem slot when the MMAs retire
          if (kb == num_kb - 1) umma_commit(&tmem_full_bar[as]);  // accumulator ready for the epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------- epilogue -------------------------------
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    uint8_t* stg = staging + q * kStagingPerWarp;
    uint32_t as = 0, aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t / n_tiles, n_blk = t - m_blk * n_tiles;
      const int row0 = m_blk * kBlockM + q * 32;
      const int n0 = n_blk * BLOCK_N;
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;
      if constexpr (epi_out_is_bf16<EPI>()) {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 64; ++c) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(t_addr + c * 64, r0);
          tmem_ld_32x32b_x32(t_addr + c * 64 + 32, r1);
          tmem_ld_wait();
          epilogue_bf16_chunk64<EPI>(p, stg, lane, row0, n0 + c * 64, r0, r1);
        }
      } else {
        float4 xa[8], xb[8];
        load_resid_chunk32<EPI>(p, lane, row0, n0, xa);
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; c += 2) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_addr + c * 32, r);
          load_resid_chunk32<EPI>(p, lane, row0, n0 + (c + 1) * 32, xb);  // in flight while chunk c is finished
          tmem_ld_wait();
          epilogue_f32_chunk32<EPI>(p, stg, lane, row0, n0 + c * 32, r, xa);
          tmem_ld_32x32b_x32(t_addr + (c + 1) * 32, r);
          if (c + 2 < BLOCK_N / 32) load_resid_chunk32<EPI>(p, lane, row0, n0 + (c + 2) * 32, xa);
          tmem_ld_wait();
          epilogue_f32_chunk32<EPI>(p, stg, lane, row0, n0 + (c + 1) * 32, r, xb);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  // resolved at run time so the library links against cudart only (loads on hosts without libcuda.so.1)
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

// 2D bf16 tensor [rows, cols] with row pitch ld (elements); box = [box_rows, 64 cols], 128B swizzle.
bool make_tmap_bf16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BLOCK_N, int EPI>
cudaError_t launch_one(cudaStream_t stream, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                       int num_sms) {
  using Cfg = GemmCfg<BLOCK_N>;
  static bool attr_set = false;
  auto kern = gemm_tcgen05_kernel<BLOCK_N, EPI>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
  const int num_tiles = m_tiles * (p.N / BLOCK_N);
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  {
    // algorithmic work: 2*M*N*K flops; bytes = A + W + out (+ resid), each touched once
    LaunchScope scope(stream, "gemm", 2.0 * p.M * static_cast<double>(p.N) * p.K,
                      2.0 * p.M * static_cast<double>(p.K) + 2.0 * p.N * static_cast<double>(p.K) +
                          4.0 * p.M * static_cast<double>(p.N));
    kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  }
  return cudaGetLastError();
}

template <int BLOCK_N>
cudaError_t launch_epi(cudaStream_t stream, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                       int num_sms) {
  switch (epi) {
    case EPI_F32: return launch_one<BLOCK_N, EPI_F32>(stream, ta, tb, p, num_sms);
    case EPI_BIAS_F32: return launch_one<BLOCK_N, EPI_BIAS_F32>(stream, ta, tb, p, num_sms);
    case EPI_BIAS_BF16: return launch_one<BLOCK_N, EPI_BIAS_BF16>(stream, ta, tb, p, num_sms);
    case EPI_BIAS_GELU_BF16: return launch_one<BLOCK_N, EPI_BIAS_GELU_BF16>(stream, ta, tb, p, num_sms);
    case EPI_BIAS_RESID_F32: return launch_one<BLOCK_N, EPI_BIAS_RESID_F32>(stream, ta, tb, p, num_sms);
    case EPI_BIAS_RELU_MASK_BF16: return launch_one<BLOCK_N, EPI_BIAS_RELU_MASK_BF16>(stream, ta, tb, p, num_sms);
    case EPI_BIAS_RESID_RELU_SPLIT: return launch_one<BLOCK_N, EPI_BIAS_RESID_RELU_SPLIT>(stream, ta, tb, p, num_sms);
    default: return cudaErrorInvalidValue;
  }
}

// Persistent grid of one CTA per SM: the slowest SM executes ceil(tiles / SMs) tiles. Pick the tile width whose
// (rounds x tile cost) is smallest; narrower tiles pay a fixed per-tile overhead and a lower MMA/smem efficiency.
int pick_block_n(int M, int N, int num_sms) {
  const int m_tiles = (M + kBlockM - 1) / kBlockM;
  int best = 0;
  double best_cost = 0.0;
  const int cand[3] = {256, 192, 128};
  for (int i = 0; i < 3; ++i) {
    const int bn = cand[i];
    if (N % bn != 0) continue;
    const long tiles = static_cast<long>(m_tiles) * (N / bn);
    const long rounds = (tiles + num_sms - 1) / num_sms;
    // measured on B200 (profiles/gemm_bench.py, K = 768): time per tile ~ (bn + 136); 128-wide tiles are
    // shared-memory-bandwidth bound and cost about as much as 192-wide ones
    const double tile_cost = (bn == 128) ? 320.0 : bn + 136.0;
    const double cost = rounds * tile_cost;
    if (best == 0 || cost < best_cost) { best = bn; best_cost = cost; }
  }
  return best ? best : 128;
}

}  // namespace

int device_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// See kernels.h for the contract.
const char* gemm_bf16_tn(cudaStream_t stream, int epi, const __nv_bfloat16* A, int64_t a_rows, int64_t a_cols,
                         int64_t lda, const __nv_bfloat16* W, int64_t ldw, GemmParams p, int block_n) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return "gemm: empty problem";
  if (p.K % kBlockK != 0) return "gemm: K must be a multiple of 64";
  if (p.n_seg < 1 || p.n_seg > kMaxGemmSegs) return "gemm: bad segment count";
  if (p.seg_kblocks * p.n_seg * kBlockK != p.K) return "gemm: segments do not tile K";
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15)) return "gemm: operands must be 16B aligned";
  if ((lda * 2) % 16 != 0 || (ldw * 2) % 16 != 0) return "gemm: row pitch must be a multiple of 16 bytes";
  if (p.N % 64 != 0) return "gemm: N must be a multiple of 64";
  const int num_sms = device_num_sms();
  if (block_n == 0) block_n = pick_block_n(p.M, p.N, num_sms);
  if (block_n != 128 && block_n != 192 && block_n != 256) return "gemm: block_n must be 128, 192 or 256";
  if (p.N % block_n != 0) return "gemm: N must be a multiple of block_n";
  if (epi != EPI_F32 && p.bias == nullptr) return "gemm: epilogue needs a bias";
  if ((epi == EPI_BIAS_RESID_F32 || epi == EPI_BIAS_RESID_RELU_SPLIT) && p.resid == nullptr) return "gemm: epilogue needs a residual";
  if (epi == EPI_BIAS_RELU_MASK_BF16 && (p.mask_hp < 2 || p.mask_wp < 2)) return "gemm: mask grid missing";

  CUtensorMap ta, tb;
  if (!make_tmap_bf16(&ta, A, a_rows, a_cols, lda, kBlockM)) return "gemm: cuTensorMapEncodeTiled(A) failed";
  if (!make_tmap_bf16(&tb, W, p.N, p.K, ldw, block_n)) return "gemm: cuTensorMapEncodeTiled(W) failed";
  cudaError_t e = (block_n == 256)   ? launch_epi<256>(stream, epi, ta, tb, p, num_sms)
                  : (block_n == 192) ? launch_epi<192>(stream, epi, ta, tb, p, num_sms)
                                     : launch_epi<128>(stream, epi, ta, tb, p, num_sms);
  if (e != cudaSuccess) return cudaGetErrorString(e);
  return nullptr;
}

}  // namespace cebc
