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

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N > 128) ? 4 : 6;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float quick_gelu(float x) {
  // x * sigmoid(1.702 x)   (reference: blocks.py:17-19)
  return __fdividef(x, 1.0f + __expf(-1.702f * x));
}

// One thread owns one output row; v[32] are 32 consecutive accumulator columns starting at global column n.
template <int EPI>
__device__ __forceinline__ void epilogue_row32(const GemmParams& p, int row, int n, const uint32_t (&r)[32]) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);

  if constexpr (EPI != EPI_F32) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  } else {
    if (p.bias != nullptr) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = __ldg(b4 + j);
        v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    }
  }

  if constexpr (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_RESID_RELU_SPLIT) {
    const float4* r4 = reinterpret_cast<const float4*>(p.resid + static_cast<size_t>(row) * p.ldr + n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 x = r4[j];
      v[4 * j + 0] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
    }
  }

  if constexpr (EPI == EPI_F32 || EPI == EPI_BIAS_F32 || EPI == EPI_BIAS_RESID_F32) {
    float4* o4 = reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + n);
#pragma unroll
    for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_RELU_MASK_BF16) {
    if constexpr (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
    }
    if constexpr (EPI == EPI_BIAS_RELU_MASK_BF16) {
      // zero-bordered grid: rows are (image, py, px) over a (mask_hp x mask_wp) padded grid; border rows must stay 0
      const int rpi = p.mask_hp * p.mask_wp;
      const int q = row % rpi;
      const int py = q / p.mask_wp, px = q - py * p.mask_wp;
      const bool border = (py == 0) || (py == p.mask_hp - 1) || (px == 0) || (px == p.mask_wp - 1);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = border ? 0.0f : fmaxf(v[j], 0.0f);
    }
    uint4* o4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldo + n);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 u;
      u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
      u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
      u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
      u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
      o4[j] = u;
    }
  } else if constexpr (EPI == EPI_BIAS_RESID_RELU_SPLIT) {
    // relu, then split into hi + lo bf16 so the next GEMM can recover ~fp32 accuracy: x ~= hi + lo
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldo + n;
    uint4* ohi = reinterpret_cast<uint4*>(o);
    uint4* olo = reinterpret_cast<uint4*>(o + p.N);
    float lo[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v[j] = fmaxf(v[j], 0.0f);
      const float h = __bfloat162float(__float2bfloat16_rn(v[j]));
      lo[j] = v[j] - h;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 u, w;
      u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
      u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
      u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
      u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
      w.x = pack_bf16x2(lo[8 * j + 0], lo[8 * j + 1]);
      w.y = pack_bf16x2(lo[8 * j + 2], lo[8 * j + 3]);
      w.z = pack_bf16x2(lo[8 * j + 4], lo[8 * j + 5]);
      w.w = pack_bf16x2(lo[8 * j + 6], lo[8 * j + 7]);
      ohi[j] = u;
      olo[j] = w;
    }
  }
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
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
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kBlockM, BLOCK_N);
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
    uint32_t as = 0, aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t / n_tiles, n_blk = t - m_blk * n_tiles;
      const int row = m_blk * kBlockM + q * 32 + lane;
      const int n0 = n_blk * BLOCK_N;
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_addr + c * 32, r);
        tmem_ld_wait();
        if (row < p.M) epilogue_row32<EPI>(p, row, n0 + c * 32, r);
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
  if (p.N % 32 != 0) return "gemm: N must be a multiple of 32";
  if (block_n == 0) block_n = (p.N % 256 == 0) ? 256 : 128;
  if (block_n != 128 && block_n != 256) return "gemm: block_n must be 128 or 256";
  if (p.N % block_n != 0) return "gemm: N must be a multiple of block_n";
  if (epi != EPI_F32 && p.bias == nullptr) return "gemm: epilogue needs a bias";
  if ((epi == EPI_BIAS_RESID_F32 || epi == EPI_BIAS_RESID_RELU_SPLIT) && p.resid == nullptr) return "gemm: epilogue needs a residual";
  if (epi == EPI_BIAS_RELU_MASK_BF16 && (p.mask_hp < 3 || p.mask_wp < 3)) return "gemm: mask grid missing";

  CUtensorMap ta, tb;
  if (!make_tmap_bf16(&ta, A, a_rows, a_cols, lda, kBlockM)) return "gemm: cuTensorMapEncodeTiled(A) failed";
  if (!make_tmap_bf16(&tb, W, p.N, p.K, ldw, block_n)) return "gemm: cuTensorMapEncodeTiled(W) failed";
  const int num_sms = device_num_sms();
  cudaError_t e = (block_n == 256) ? launch_epi<256>(stream, epi, ta, tb, p, num_sms)
                                   : launch_epi<128>(stream, epi, ta, tb, p, num_sms);
  if (e != cudaSuccess) return cudaGetErrorString(e);
  return nullptr;
}

}  // namespace cebc
