// CTA-pair (cta_group::2) variant of the tcgen05 GEMM:  D[M,N] = A[M,K] * W[N,K]^T, bf16 x bf16 -> fp32 in TMEM.
//
// Why a pair: with one CTA per 128xBLOCK_N tile the shared-memory port carries the TMA writes *and* the MMA operand
// reads of A (16 KB) + B (32 KB) per 64-deep k-block (96 KB per 512 MMA cycles > 128 B/clk) -- measured ~800 cycles per
// k-block on B200 (profiles/r01_*). A CTA pair computes a 256xBLOCK_N tile with M split over the two SMs and each SM
// holding only half of B: 32 KB written + 32 KB read per k-block and SM.
//
// Cluster of 2 CTAs (one per SM of a TPC), 384 threads each:
//   warp 0      TMA producer (both CTAs): own 128 rows of A + own half (BLOCK_N/2 rows) of W; transaction bytes of both
//               CTAs are reported to the LEADER's full barrier (cp.async.bulk.tensor ... cta_group::2)
//   warp 1      MMA issuer (leader CTA only): tcgen05.mma.cta_group::2 256 x BLOCK_N x 16; tcgen05.commit multicast
//               releases the smem slot in both CTAs and publishes the accumulator to both epilogues
//   warp 2      TMEM allocator (both CTAs, cta_group::2)
//   warps 4-11  epilogue (both CTAs): 2 warps per TMEM lane quadrant, each owning half of the tile's columns;
//               tcgen05.ld -> XOR-swizzled smem transpose -> coalesced global accesses; the tile's bias lives in smem
// Fused epilogues and K-segment addressing: kernels.h.
//
// Tried and not kept (measured on B200): clusters of two pairs sharing the weight tile by TMA multicast
// (only 33 clusters of 4 fit on 148 SMs; chip rate unchanged), LayerNorm folded into the epilogues either side of it
// (+0.6 % on the step for three more epilogues), a TMA round trip of the residual tile. Their code lives in the history
// of this file (round 1).
#include "common.cuh"
#include "kernels.h"

namespace cebc {

namespace {

constexpr int kBlockM = 128;   // rows per CTA (256 per pair)
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int kStagingPerWarp = 32 * 128;  // 32 rows x 128 B, XOR-swizzled 16 B slots

template <int BLOCK_N>
struct Cfg2 {
  static constexpr int kABytes = kBlockM * kBlockK * 2;          // 16 KB
  static constexpr int kBBytes = (BLOCK_N / 2) * kBlockK * 2;    // this CTA's half of the W tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N > 192) ? 5 : 6;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int kStagingBytes = kEpiWarps * kStagingPerWarp;
  static constexpr int kBiasBytes = 2 * BLOCK_N * 4;  // [2][BLOCK_N] bias, double-buffered by accumulator stage
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kBiasBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// QuickGELU x * sigmoid(1.702 x) (reference: blocks.py:17-19) through sigmoid(z) = (1 + tanh(z / 2)) / 2: ONE
// special-function op (MUFU.TANH) per element instead of two (EX2 + RCP). The c_fc epilogue is MUFU-bound (32768 elements per 128 x 256 tile and SM against a ~9.4k-cycle
// mainloop); tanh.approx has 2^-11 relative error, the size of the fp16 rounding that follows.
__device__ __forceinline__ float quick_gelu2(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

template <int EPI>
constexpr bool out_is_bf16() {
  return EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_RELU_MASK_BF16 ||
         EPI == EPI_BIAS_RESID16_RELU_MASK_BF16;
}
template <int EPI>
constexpr bool has_resid() {
  return EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_RESID_RELU_SPLIT || EPI == EPI_BIAS_UPSKIP_RELU_SPLIT;
}
template <int EPI>
constexpr bool is_relu_split() { return EPI == EPI_BIAS_RESID_RELU_SPLIT || EPI == EPI_BIAS_UPSKIP_RELU_SPLIT; }

// ---- bf16-output epilogues, 32 accumulator columns per pass -------------------------------------------------------
// phase 1: thread = output row (as tcgen05.ld delivers it): bias (smem broadcast) + activation, pack, 4 x 16 B into the
// warp's staging tile (64 B rows, slot ^= (row >> 1) & 3). phase 2: lane -> (row = it*8 + lane/4, slot = lane%4): one
// warp store writes 8 complete 64 B row segments.
template <int EPI>
__device__ __forceinline__ void epi_bf16_chunk32(const GemmParams& p, uint8_t* stg, const float* bias_s, int lane, int row0,
                                                 int n, const uint32_t (&r)[32], int tma_half = -1) {
  const int row = row0 + lane;
  bool border = false;
  if constexpr (EPI == EPI_BIAS_RELU_MASK_BF16 || EPI == EPI_BIAS_RESID16_RELU_MASK_BF16) {
    const int rpi = p.mask_hp * p.mask_wp;
    const int q = row % rpi;
    const int py = q / p.mask_wp, px = q - py * p.mask_wp;
    border = (py == p.mask_hp - 1) || (px == p.mask_wp - 1) || (p.mask_lead && (py == 0 || px == 0));
  }
  const float4* b4 = reinterpret_cast<const float4*>(bias_s);
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 ba = b4[2 * j], bb = b4[2 * j + 1];
    float v[8];
    v[0] = __uint_as_float(r[8 * j + 0]) + ba.x; v[1] = __uint_as_float(r[8 * j + 1]) + ba.y;
    v[2] = __uint_as_float(r[8 * j + 2]) + ba.z; v[3] = __uint_as_float(r[8 * j + 3]) + ba.w;
    v[4] = __uint_as_float(r[8 * j + 4]) + bb.x; v[5] = __uint_as_float(r[8 * j + 5]) + bb.y;
    v[6] = __uint_as_float(r[8 * j + 6]) + bb.z; v[7] = __uint_as_float(r[8 * j + 7]) + bb.w;
    if constexpr (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = quick_gelu2(v[k]);
    }
    if constexpr (EPI == EPI_BIAS_RESID16_RELU_MASK_BF16) {
      // identity branch: 8 values of this thread's row in the 16-bit output format
      uint4 rv = make_uint4(0u, 0u, 0u, 0u);
      if (row < p.M && !border)
        rv = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.resid16) + static_cast<size_t>(row) * p.ldr + n + 8 * j);
      const uint32_t ru[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 f;
        if (p.out_fp16) f = __half22float2(*reinterpret_cast<const __half2*>(&ru[k]));
        else f = make_float2(__uint_as_float(ru[k] << 16), __uint_as_float(ru[k] & 0xFFFF0000u));
        v[2 * k] += f.x; v[2 * k + 1] += f.y;
      }
    }
    if constexpr (EPI == EPI_BIAS_RELU_MASK_BF16 || EPI == EPI_BIAS_RESID16_RELU_MASK_BF16) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = border ? 0.0f : fmaxf(v[k], 0.0f);
    }
    const uint4 pk = make_uint4(pack16x2(v[0], v[1], p.out_fp16), pack16x2(v[2], v[3], p.out_fp16),
                                pack16x2(v[4], v[5], p.out_fp16), pack16x2(v[6], v[7], p.out_fp16));
    if (tma_half >= 0)  // row `lane` of a 32 x 64 box in the 128B-swizzled layout a bulk tensor store reads
      *reinterpret_cast<uint4*>(stg + lane * 128 + (((tma_half * 4 + j) ^ (lane & 7)) << 4)) = pk;
    else
      *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ sw) << 4)) = pk;
  }
  if (tma_half >= 0) return;
  __syncwarp();
  const int slot = lane & 3, rsub = lane >> 2;
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out);
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int rr = it * 8 + rsub;
    const uint4 u = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((slot ^ ((rr >> 1) & 3)) << 4));
    if (row0 + rr < p.M) *reinterpret_cast<uint4*>(out + static_cast<size_t>(row0 + rr) * p.ldo + n + slot * 8) = u;
  }
  __syncwarp();
}

// ---- fp32-finishing epilogues, 32 columns per pass (128 B rows, slot ^= row & 7) ----------------------------------
template <int EPI>
__device__ __forceinline__ void load_resid32(const GemmParams& p, int lane, int row0, int n, float4 (&x)[8]) {
  if constexpr (EPI == EPI_BIAS_UPSKIP_RELU_SPLIT) {
    // residual = bilinear_up(Y)[cell of the row] (F.interpolate, align_corners = False; models/clip/model.py:195-196),
    // four coarse rows per fine cell, weights as in resample_to_padded_kernel; border rows of the grid add nothing
    const int c4 = (lane & 7) * 4, rsub = lane >> 3;
    const int gh = p.mask_hp - 1, gw = p.mask_wp - 1, rpi = p.mask_hp * p.mask_wp;
    const float inv_sy = static_cast<float>(p.up_hp) / static_cast<float>(gh), inv_sx = static_cast<float>(p.up_wp) / static_cast<float>(gw);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int grow = row0 + it * 4 + rsub;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (grow < p.M) {
        const int win = grow / rpi, q = grow - win * rpi;
        const int py = q / p.mask_wp, px = q - py * p.mask_wp;
        if (py < gh && px < gw) {
          int y0, y1, x0, x1;
          float ly, lx;
          bilinear_src(py, inv_sy, p.up_hp, y0, y1, ly);
          bilinear_src(px, inv_sx, p.up_wp, x0, x1, lx);
          const float* base = p.resid + static_cast<size_t>(win) * p.up_hp * p.up_wp * p.ldr + n + c4;
          const float4 a = *reinterpret_cast<const float4*>(base + static_cast<size_t>(y0 * p.up_wp + x0) * p.ldr);
          const float4 b = *reinterpret_cast<const float4*>(base + static_cast<size_t>(y0 * p.up_wp + x1) * p.ldr);
          const float4 c = *reinterpret_cast<const float4*>(base + static_cast<size_t>(y1 * p.up_wp + x0) * p.ldr);
          const float4 d = *reinterpret_cast<const float4*>(base + static_cast<size_t>(y1 * p.up_wp + x1) * p.ldr);
          const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
          v.x = w00 * a.x + w01 * b.x + w10 * c.x + w11 * d.x;
          v.y = w00 * a.y + w01 * b.y + w10 * c.y + w11 * d.y;
          v.z = w00 * a.z + w01 * b.z + w10 * c.z + w11 * d.z;
          v.w = w00 * a.w + w01 * b.w + w10 * c.w + w11 * d.w;
        }
      }
      x[it] = v;
    }
  } else if constexpr (has_resid<EPI>()) {
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
__device__ __forceinline__ void epi_f32_chunk32(const GemmParams& p, uint8_t* stg, const float* bias_s, int lane, int row0,
                                                int n, const uint32_t (&r)[32], const float4 (&x)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
        make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
  __syncwarp();
  const int slot = lane & 7, c4 = slot * 4, rsub = lane >> 3;
  const float4 b = *reinterpret_cast<const float4*>(bias_s + c4);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + rsub;
    const int grow = row0 + rr;
    float4 v = *reinterpret_cast<const float4*>(stg + rr * 128 + ((slot ^ (rr & 7)) << 4));
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    if constexpr (has_resid<EPI>()) { v.x += x[it].x; v.y += x[it].y; v.z += x[it].z; v.w += x[it].w; }
    if (grow < p.M) {
      if constexpr (is_relu_split<EPI>()) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        const float hx = round16(v.x, p.out_fp16), hy = round16(v.y, p.out_fp16);
        const float hz = round16(v.z, p.out_fp16), hw = round16(v.w, p.out_fp16);
        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(grow) * p.ldo + n + c4;
        *reinterpret_cast<uint2*>(o) = make_uint2(pack16x2(hx, hy, p.out_fp16), pack16x2(hz, hw, p.out_fp16));
        if (p.split_lo)
          *reinterpret_cast<uint2*>(o + p.N) =
              make_uint2(pack16x2(v.x - hx, v.y - hy, p.out_fp16), pack16x2(v.z - hz, v.w - hw, p.out_fp16));
      } else {
        *reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<size_t>(grow) * p.ldo + n + c4) = v;
      }
    }
  }
  __syncwarp();
}

// ---- projection fused with the first half of the EBC head (EPI_BIAS_HEAD_PARTIAL) ---------------------------------
// thread = output row; over its `ncols` columns of the tile: ss += f^2, dot[b] += f * tmat[b][col], f = acc + bias.
// tm_s: the tile's slice of the text matrix in smem, [NB][tile_n] f32 (rows >= head_bins are zero); reads are warp
// broadcasts of 16 B. Nothing of f is written to memory.
template <int NB>
__device__ __forceinline__ void head_partial_cols(uint32_t t_addr, const float* tm_s, int tile_n, const float* bias_s, int ncols,
                                                  float& ss, float (&dot)[NB]) {
#pragma unroll 1
  for (int c = 0; c < ncols; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_addr + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c + j);
      const float f0 = __uint_as_float(r[j]) + b4.x, f1 = __uint_as_float(r[j + 1]) + b4.y;
      const float f2 = __uint_as_float(r[j + 2]) + b4.z, f3 = __uint_as_float(r[j + 3]) + b4.w;
      ss += (f0 * f0 + f1 * f1) + (f2 * f2 + f3 * f3);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float4 t = *reinterpret_cast<const float4*>(tm_s + b * tile_n + c + j);
        dot[b] += (f0 * t.x + f1 * t.y) + (f2 * t.z + f3 * t.w);
      }
    }
  }
}

template <int NB>
__device__ __forceinline__ void head_partial_tile(const GemmParams& p, uint32_t t_addr, const float* tm_s, int tile_n,
                                                  const float* bias_s, int ncols, int row, int part) {
  float ss = 0.f, dot[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) dot[b] = 0.f;
  head_partial_cols<NB>(t_addr, tm_s, tile_n, bias_s, ncols, ss, dot);
  if (row < p.M) {
    float* o = static_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + static_cast<size_t>(part) * (1 + p.head_bins);
    o[0] = ss;
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (b < p.head_bins) o[1 + b] = dot[b];
  }
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                     const __grid_constant__ CUtensorMap tma_o, const __grid_constant__ GemmParams p) {
  using Cfg = Cfg2<BLOCK_N>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int HALF_N = BLOCK_N / 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + STAGES * Cfg::kStageBytes;
  float* bias_s = reinterpret_cast<float*>(staging + Cfg::kStagingBytes);  // [2][BLOCK_N]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_s) + Cfg::kBiasBytes);
  uint64_t* full_bar = bars;                         // [STAGES]  used in the leader CTA only
  uint64_t* empty_bar = bars + STAGES;               // [STAGES]  one set per CTA
  uint64_t* tmem_full_bar = bars + 2 * STAGES;       // [2]       one set per CTA
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;  // [2]       used in the leader CTA only
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // rank inside the CTA pair
  const bool leader = rank == 0;

  const int m_tiles = (p.M + 2 * kBlockM - 1) / (2 * kBlockM);  // 256-row pair tiles
  const int n_tiles = p.N / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.K / kBlockK;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (p.tma_out) tma_prefetch_desc(&tma_o);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);   // leader's arrive.expect_tx + the peer producer's remote arrive
      mbar_init(&empty_bar[s], 1);  // tcgen05.commit multicast from the leader
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 2 * kEpiWarps);  // every epilogue warp of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair<Cfg::kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // everything above overlapped the tail of the previous kernel; its results (A, the residual) are needed from here on
  pdl_launch_dependents();
  // The weights do not depend on the previous kernel of the stream (p.w_prefetch; pack-time launches whose weights were
  // written by the kernel just before them clear it): the producer requests the weight half-tiles of the first k-blocks
  // of its first tile BEFORE the dependency wait, so that only the activation loads are exposed after it.
  int pre_kb = 0;
  if (warp == 0 && p.w_prefetch && pair_id < num_tiles) {
    pre_kb = num_kb < STAGES ? num_kb : STAGES;
    if (lane == 0) {
      const int n_blk = pair_id % n_tiles;
      const int n0 = n_blk * BLOCK_N + static_cast<int>(rank) * HALF_N;
      for (int kb = 0; kb < pre_kb; ++kb) {  // fresh barriers: every slot is free
        const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[kb]), 0);
        if (leader) mbar_arrive_expect_tx(&full_bar[kb], 2u * Cfg::kABytes + 2u * Cfg::kBBytes);
        else mbar_arrive_cluster(leader_full);
        tma_load_2d_pair(smem + kb * Cfg::kStageBytes + Cfg::kABytes, &tma_b, leader_full, kb * kBlockK, n0);
      }
    }
    __syncwarp();
  }
  pdl_wait();

  if (warp == 0) {
    // ------------------------------- TMA producer (both CTAs) -------------------------------
    uint32_t stage = 0, phase = 0;
    for (int t = pair_id; t < num_tiles; t += num_pairs) {
      const int m_blk = t / n_tiles, n_blk = t - m_blk * n_tiles;
      const int m0 = m_blk * 2 * kBlockM + static_cast<int>(rank) * kBlockM;
      const int n0 = n_blk * BLOCK_N + static_cast<int>(rank) * HALF_N;
      int seg = 0, kk = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const bool weights_done = t == pair_id && kb < pre_kb;  // barrier armed, weights requested before the wait
        if (!weights_done) mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (!weights_done) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2u * Cfg::kABytes + 2u * Cfg::kBBytes);
            else mbar_arrive_cluster(leader_full);
          }
          tma_load_2d_pair(sa, &tma_a, leader_full, p.seg_col_start[seg] + kk * kBlockK, m0 + p.seg_row_shift[seg]);
          if (!weights_done) tma_load_2d_pair(sa + Cfg::kABytes, &tma_b, leader_full, kb * kBlockK, n0);
        }
        __syncwarp();
        if (++kk == p.seg_kblocks) { kk = 0; ++seg; }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------- MMA issuer (leader CTA) -------------------------------
    const uint32_t idesc = umma_idesc_bf16_f32(2 * kBlockM, BLOCK_N, p.ab_fp16);
    uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
    constexpr uint16_t kBothCtas = 3;
    for (int t = pair_id; t < num_tiles; t += num_pairs) {
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
            umma_bf16_ss_pair(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair_mcast(&empty_bar[stage], kBothCtas);                          // the pair is done with the slot
          if (kb == num_kb - 1) umma_commit_pair_mcast(&tmem_full_bar[as], kBothCtas);  // accumulator ready in both CTAs
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------- epilogue (both CTAs, 8 warps) -------------------------------
    const int ew = warp - 4;
    const int q = ew & 3;        // TMEM lane quadrant (== warp % 4)
    const int half = ew >> 2;    // which half of the tile's columns
    const int et = threadIdx.x - 128;
    uint8_t* stg = staging + ew * kStagingPerWarp;
    const uint32_t leader_tmem_empty0 = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);
    const uint32_t leader_tmem_empty1 = mapa_u32(smem_u32(&tmem_empty_bar[1]), 0);
    uint32_t as = 0, aphase = 0;
    for (int t = pair_id; t < num_tiles; t += num_pairs) {
      const int m_blk = t / n_tiles, n_blk = t - m_blk * n_tiles;
      const int row0 = m_blk * 2 * kBlockM + static_cast<int>(rank) * kBlockM + q * 32;
      const int n0 = n_blk * BLOCK_N;
      // the tile's bias -> smem (double-buffered by accumulator stage); overlaps the wait for the accumulator
      float* bs = bias_s + as * BLOCK_N;
      if constexpr (EPI == EPI_BIAS_HEAD_PARTIAL) {
        // the tile's slice of the text matrix lives in the (otherwise unused) staging area: wait until every warp is
        // done with the previous tile's slice before overwriting it
        named_bar_sync(1, kEpiWarps * 32);
        float* tm_s = reinterpret_cast<float*>(staging);
        const int nb_pad = p.head_bins <= 8 ? 8 : (p.head_bins <= 16 ? 16 : 32);
        for (int i = et; i < nb_pad * BLOCK_N; i += kEpiWarps * 32) {
          const int b = i / BLOCK_N, col = i - b * BLOCK_N;
          tm_s[i] = b < p.head_bins ? __ldg(p.head_tmat + static_cast<size_t>(b) * p.N + n0 + col) : 0.0f;
        }
      }
      if (et < BLOCK_N) bs[et] = (EPI != EPI_F32 || p.bias != nullptr) ? __ldg(p.bias + n0 + et) : 0.0f;
      named_bar_sync(1, kEpiWarps * 32);
      const int cbase = half * HALF_N;
      float4 xa[8], xb[8];
      if constexpr (!out_is_bf16<EPI>()) {
        if (!(EPI == EPI_BIAS_RESID_F32 && p.tma_out)) load_resid32<EPI>(p, lane, row0, n0 + cbase, xa);
      }
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + cbase;
      if constexpr (EPI == EPI_BIAS_HEAD_PARTIAL) {
        const float* tm_s = reinterpret_cast<const float*>(staging) + cbase;
        const int row = row0 + lane, part = 2 * n_blk + half;
        if (p.head_bins <= 8) head_partial_tile<8>(p, t_addr, tm_s, BLOCK_N, bs + cbase, HALF_N, row, part);
        else if (p.head_bins <= 16) head_partial_tile<16>(p, t_addr, tm_s, BLOCK_N, bs + cbase, HALF_N, row, part);
        else head_partial_tile<32>(p, t_addr, tm_s, BLOCK_N, bs + cbase, HALF_N, row, part);
      } else if constexpr (out_is_bf16<EPI>()) {
        if (HALF_N % 64 == 0 && p.tma_out) {
          // 16-bit outputs through bulk tensor stores: thread = accumulator row writes its 64 columns into a 32 x 64
          // box (128B-swizzled, conflict-free 16 B stores), lane 0 sends the box; no read-back of the staging tile, no
          // per-lane global stores, full 128 B lines per row; rows >= M are clipped by the hardware
#pragma unroll 1
          for (int b = 0; b < HALF_N / 64; ++b) {
            if (lane == 0) bulk_wait_group_read0();  // the previous box has left the staging tile
            __syncwarp();
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
              const int c = 2 * b + cc;
              uint32_t r[32];
              tmem_ld_32x32b_x32(t_addr + c * 32, r);
              tmem_ld_wait();
              epi_bf16_chunk32<EPI>(p, stg, bs + cbase + c * 32, lane, row0, n0 + cbase + c * 32, r, cc);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tma_o, stg, n0 + cbase + b * 64, row0);
              bulk_commit_group();
            }
          }
        } else
#pragma unroll 1
        for (int c = 0; c < HALF_N / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_addr + c * 32, r);
          tmem_ld_wait();
          epi_bf16_chunk32<EPI>(p, stg, bs + cbase + c * 32, lane, row0, n0 + cbase + c * 32, r);
        }
      } else if (EPI == EPI_BIAS_RESID_F32 && p.tma_out) {
        // In-place residual update as a memory-side reduction: thread = accumulator row stages acc + bias (fp32) in a
        // 128B-swizzled 32 x 32 box and lane 0 issues cp.reduce.async.bulk.tensor .add -- X += tile happens in L2. The
        // SM never reads the old rows (38.7 MB per launch at 64 windows), there are no per-lane global accesses and no
        // read-back of the staging tile; x + (acc + bias) has the bits of (acc + bias) + x, and every element receives
        // exactly one addition per launch, so the result is the same, deterministically.
        constexpr int NC = HALF_N / 32;
#pragma unroll 1
        for (int c = 0; c < NC; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_addr + c * 32, r);
          if (lane == 0) bulk_wait_group_read0();  // the previous box has left the staging tile
          __syncwarp();
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(bs + cbase + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = b4[j];
            *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_float4(__uint_as_float(r[4 * j]) + b.x, __uint_as_float(r[4 * j + 1]) + b.y,
                            __uint_as_float(r[4 * j + 2]) + b.z, __uint_as_float(r[4 * j + 3]) + b.w);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tma_o, stg, n0 + cbase + c * 32, row0);
            bulk_commit_group();
          }
        }
      } else {
        constexpr int NC = HALF_N / 32;
#pragma unroll 1
        for (int c = 0; c < NC; c += 2) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_addr + c * 32, r);
          if (c + 1 < NC) load_resid32<EPI>(p, lane, row0, n0 + cbase + (c + 1) * 32, xb);
          tmem_ld_wait();
          epi_f32_chunk32<EPI>(p, stg, bs + cbase + c * 32, lane, row0, n0 + cbase + c * 32, r, xa);
          if (c + 1 < NC) {
            tmem_ld_32x32b_x32(t_addr + (c + 1) * 32, r);
            if (c + 2 < NC) load_resid32<EPI>(p, lane, row0, n0 + cbase + (c + 2) * 32, xa);
            tmem_ld_wait();
            epi_f32_chunk32<EPI>(p, stg, bs + cbase + (c + 1) * 32, lane, row0, n0 + cbase + (c + 1) * 32, r, xb);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(as == 0 ? leader_tmem_empty0 : leader_tmem_empty1);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  if constexpr (out_is_bf16<EPI>() || EPI == EPI_BIAS_RESID_F32) {
    if (warp >= 4 && lane == 0 && p.tma_out) bulk_wait_group0();  // this warp's bulk stores / reductions are complete
  }
  tc_fence_before();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the peer may still signal its barriers / read its smem
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn2() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

bool make_tmap2(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn2();
  if (!enc) return false;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fp32 [rows, cols] pitch ld: 32 x 32 boxes, 128B-swizzled (the tile of the in-place residual reduction)
bool make_tmap2_f32(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld) {
  PFN_encodeTiled enc = get_encode_fn2();
  if (!enc) return false;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr int kMaxDevices = 64;
int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
  return dev < 0 ? 0 : (dev >= kMaxDevices ? kMaxDevices - 1 : dev);
}

// CTA pairs (1 CTA per SM) that can be co-resident on the current device: 74 on B200. Cached per device; the first call on
// a device also sets the kernel's dynamic shared-memory attribute there (a per-device attribute).
template <int BLOCK_N, int EPI>
int max_pairs2(int num_sms) {
  static int cached[kMaxDevices] = {};
  const int dev = current_device();
  if (cached[dev]) return cached[dev];
  using Cfg = Cfg2<BLOCK_N>;
  auto kern = gemm2_tcgen05_kernel<BLOCK_N, EPI>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * 64); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms / 2; }
  cached[dev] = n;
  return n;
}

template <int BLOCK_N, int EPI>
cudaError_t launch_one2(cudaStream_t stream, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                        int num_sms) {
  using Cfg = Cfg2<BLOCK_N>;
  CUtensorMap to = ta;  // placeholder for the epilogues that do not use it
  GemmParams pl = p;
  pl.tma_out = 0;
  if constexpr (out_is_bf16<EPI>() && (BLOCK_N / 2) % 64 == 0) {
    // 16-bit outputs leave through bulk tensor stores when the output allows it (else per-lane stores)
    if ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && (p.ldo * 2) % 16 == 0 && make_tmap2(&to, p.out, p.M, p.N, p.ldo, 32))
      pl.tma_out = 1;
  }
  if constexpr (EPI == EPI_BIAS_RESID_F32) {
    // in place (out == resid): the residual add becomes a bulk tensor reduction (else the register path)
    if (p.out == static_cast<const void*>(p.resid) && p.ldo == p.ldr && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 &&
        p.ldo % 4 == 0 && make_tmap2_f32(&to, p.out, p.M, p.N, p.ldo))
      pl.tma_out = 1;
  }
  auto kern = gemm2_tcgen05_kernel<BLOCK_N, EPI>;
  const int max_pairs = max_pairs2<BLOCK_N, EPI>(num_sms);  // also sets the dynamic smem attribute (once per device)
  if (max_pairs <= 0) return cudaErrorInvalidConfiguration;
  const int m_tiles = (p.M + 2 * kBlockM - 1) / (2 * kBlockM);
  const int num_tiles = m_tiles * (p.N / BLOCK_N);
  const int pairs = num_tiles < max_pairs ? num_tiles : max_pairs;
  cudaError_t e;
  {
    LaunchScope scope(stream, "gemm", 2.0 * p.M * static_cast<double>(p.N) * p.K,
                      2.0 * p.M * static_cast<double>(p.K) + 2.0 * p.N * static_cast<double>(p.K) +
                          4.0 * p.M * static_cast<double>(p.N));
    e = launch_pdl(kern, dim3(2 * pairs), dim3(kThreads), Cfg::kSmemBytes, stream, 2, ta, tb, to, pl);
  }
  return e != cudaSuccess ? e : cudaGetLastError();
}

template <int BLOCK_N>
cudaError_t launch_epi2(cudaStream_t stream, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                        int num_sms) {
  switch (epi) {
    case EPI_F32: return launch_one2<BLOCK_N, EPI_F32>(stream, ta, tb, p, num_sms);
    case EPI_BIAS_F32: return launch_one2<BLOCK_N, EPI_BIAS_F32>(stream, ta, tb, p, num_sms);
    case EPI_BIAS_BF16: return launch_one2<BLOCK_N, EPI_BIAS_BF16>(stream, ta, tb, p, num_sms);
    case EPI_BIAS_GELU_BF16: return launch_one2<BLOCK_N, EPI_BIAS_GELU_BF16>(stream, ta, tb, p, num_sms);
    case EPI_BIAS_RESID_F32: return launch_one2<BLOCK_N, EPI_BIAS_RESID_F32>(stream, ta, tb, p, num_sms);
    case EPI_BIAS_RELU_MASK_BF16: return launch_one2<BLOCK_N, EPI_BIAS_RELU_MASK_BF16>(stream, ta, tb, p, num_sms);
    case EPI_BIAS_RESID_RELU_SPLIT: return launch_one2<BLOCK_N, EPI_BIAS_RESID_RELU_SPLIT>(stream, ta, tb, p, num_sms);
    case EPI_BIAS_UPSKIP_RELU_SPLIT: return launch_one2<BLOCK_N, EPI_BIAS_UPSKIP_RELU_SPLIT>(stream, ta, tb, p, num_sms);
    case EPI_BIAS_RESID16_RELU_MASK_BF16:
      return launch_one2<BLOCK_N, EPI_BIAS_RESID16_RELU_MASK_BF16>(stream, ta, tb, p, num_sms);
    case EPI_BIAS_HEAD_PARTIAL:
      if constexpr (BLOCK_N == 256) return launch_one2<256, EPI_BIAS_HEAD_PARTIAL>(stream, ta, tb, p, num_sms);
      else return cudaErrorInvalidValue;
    default: return cudaErrorInvalidValue;
  }
}

// ceil(tiles / pairs) rounds; measured on B200 (profiles/gemm_bench.py) the time of one round is ~ (block_n + 320):
// pick the cheapest tile width that divides N
int pick_block_n2(int M, int N, int num_sms) {
  const int m_tiles = (M + 2 * kBlockM - 1) / (2 * kBlockM);
  const int pairs = num_sms / 2;
  int best = 0;
  double best_cost = 0.0;
  const int cand[3] = {256, 192, 128};
  for (int i = 0; i < 3; ++i) {
    const int bn = cand[i];
    if (N % bn != 0) continue;
    const long tiles = static_cast<long>(m_tiles) * (N / bn);
    const long rounds = (tiles + pairs - 1) / pairs;
    const double cost = rounds * (bn + 320.0);
    if (best == 0 || cost < best_cost) { best = bn; best_cost = cost; }
  }
  return best ? best : 128;
}

}  // namespace

int device_num_sms() {
  static int cached[kMaxDevices] = {};
  const int dev = current_device();
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
    cached[dev] = n;
  }
  return cached[dev];
}

int gemm2_pick_block_n(int M, int N) { return pick_block_n2(M, N, device_num_sms()); }

const char* gemm2_bf16_tn(cudaStream_t stream, int epi, const __nv_bfloat16* A, int64_t a_rows, int64_t a_cols,
                          int64_t lda, const __nv_bfloat16* W, int64_t ldw, GemmParams p, int block_n) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return "gemm: empty problem";
  if (p.K % kBlockK != 0) return "gemm: K must be a multiple of 64";
  if (p.n_seg < 1 || p.n_seg > kMaxGemmSegs) return "gemm: bad segment count";
  if (p.seg_kblocks * p.n_seg * kBlockK != p.K) return "gemm: segments do not tile K";
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15)) return "gemm: operands must be 16B aligned";
  if ((lda * 2) % 16 != 0 || (ldw * 2) % 16 != 0) return "gemm: row pitch must be a multiple of 16 bytes";
  if (p.N % 64 != 0) return "gemm: N must be a multiple of 64";
  const int num_sms = device_num_sms();
  if (block_n == 0) block_n = pick_block_n2(p.M, p.N, num_sms);
  if (block_n != 128 && block_n != 192 && block_n != 256) return "gemm: block_n must be 128, 192 or 256";
  if (p.N % block_n != 0) return "gemm: N must be a multiple of block_n";
  if (epi != EPI_F32 && p.bias == nullptr) return "gemm: epilogue needs a bias";
  if ((epi == EPI_BIAS_RESID_F32 || epi == EPI_BIAS_RESID_RELU_SPLIT || epi == EPI_BIAS_UPSKIP_RELU_SPLIT) && p.resid == nullptr)
    return "gemm: epilogue needs a residual";
  if (epi == EPI_BIAS_UPSKIP_RELU_SPLIT && (p.mask_hp < 2 || p.mask_wp < 2 || p.up_hp < 1 || p.up_wp < 1))
    return "gemm: upsampled-skip epilogue needs the grid (mask_hp, mask_wp) and the patch grid (up_hp, up_wp)";
  if ((epi == EPI_BIAS_RELU_MASK_BF16 || epi == EPI_BIAS_RESID16_RELU_MASK_BF16) && (p.mask_hp < 2 || p.mask_wp < 2))
    return "gemm: mask grid missing";
  if (epi == EPI_BIAS_RESID16_RELU_MASK_BF16 &&
      (p.resid16 == nullptr || (reinterpret_cast<uintptr_t>(p.resid16) & 15) || (p.ldr * 2) % 16 != 0))
    return "gemm: 16-bit residual missing or not 16-byte aligned";
  if (epi == EPI_BIAS_HEAD_PARTIAL) {
    if (p.head_tmat == nullptr || p.head_bins < 1 || p.head_bins > 32) return "gemm: head epilogue needs the text matrix and 1..32 bins";
    if (p.N % 256 != 0) return "gemm: head epilogue needs N to be a multiple of 256";
    block_n = 256;
  }

  CUtensorMap ta, tb;
  if (!make_tmap2(&ta, A, a_rows, a_cols, lda, kBlockM)) return "gemm: cuTensorMapEncodeTiled(A) failed";
  if (!make_tmap2(&tb, W, p.N, p.K, ldw, block_n / 2)) return "gemm: cuTensorMapEncodeTiled(W) failed";
  const cudaError_t e = (block_n == 256)   ? launch_epi2<256>(stream, epi, ta, tb, p, num_sms)
                        : (block_n == 192) ? launch_epi2<192>(stream, epi, ta, tb, p, num_sms)
                                           : launch_epi2<128>(stream, epi, ta, tb, p, num_sms);
  if (e != cudaSuccess) return cudaGetErrorString(e);
  return nullptr;
}

}  // namespace cebc
