// Persistent, warp-specialised tcgen05 attention: softmax(q k^T / 8) v per (window, head), 64-dim heads, <= 256 queries
// and <= 256 keys per item (229 for a 224x224 window with 32 prompt tokens).
// Replaces nn.MultiheadAttention -> F.scaled_dot_product_attention (/root/reference/models/clip/_clip/blocks.py:25,35-37);
// deep-VPT constant prompt keys/values are appended as in attention.cu (reference models/clip/model.py:164-183).
//
// One CTA per SM loops over (window, head) items; all stages of consecutive items overlap:
//   warp 0       TMA producer: Q [256 x 64], K, V [256 x 64] (constant prompt rows first) of item i+1 are fetched into
//                the second smem stage while item i is being processed
//   warp 1       MMA issuer: S = Q_t K^T (128 x 256 x 64) of the NEXT 128-query tile into the free TMEM buffer, and
//                O = P V of the current one with P read straight from TMEM (tcgen05.mma A-operand in tensor memory) and
//                V consumed in place as MN-major B
//   warp 2       TMEM allocator (all 512 columns: two S buffers of 128 lanes x 256 fp32)
//   warps 4-11   two softmax groups that split the keys of one tile (group h: keys [128h, 128h+128)): thread = query row
//                = TMEM lane; exp2 and row sum from TMEM in a single pass, P written back to TMEM as packed bf16 over
//                the consumed S columns (no smem round trip); finally O * 1/rowsum -> 16-bit -> global
// Scores and probabilities never leave the SM: HBM traffic is Q, K, V in and O out.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace cebc {

namespace {

constexpr int kThreadsF = 384;
constexpr int kQTileBytes = 128 * 128;   // one 128-query tile
constexpr int kQBytes = 2 * kQTileBytes; // 256 query rows x 64 dims bf16
constexpr int kKVBytes = 256 * 128;      // 256 key rows x 64 dims bf16
constexpr int kStageBytesF = kQBytes + 2 * kKVBytes;  // 96 KB
constexpr int kXchgBytes = 2 * 2 * 2 * 128 * 4;   // row max / row sum exchange between the two softmax groups
constexpr int kOutStageBytes = 8 * 32 * 64;        // per softmax warp: 32 rows x 64 B of packed output, XOR-swizzled
constexpr int kSmemF = 2 * kStageBytesF + kXchgBytes + kOutStageBytes + 1024 /*align*/ + 256 /*barriers*/;
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
                                                float& row_sum, uint32_t p_taddr, int dbg = 0) {
  uint32_t pk[16];
  if (dbg & 1) {  // EXPERIMENT: no exp
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float p0 = __uint_as_float(v[2 * j]) * scale - m_scaled, p1 = __uint_as_float(v[2 * j + 1]) * scale - m_scaled;
      row_sum += p0 + p1;
      pk[j] = pack_bf16x2(p0, p1);
    }
  } else if (lim >= 32) {
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
  if (!(dbg & 4)) tmem_st_32x32b_x16(p_taddr, pk);
}

// Kernel structure: a CTA walks over "tiles" u = (item, 128-query tile); tile u lives in TMEM buffer u & 1 (256 columns):
//   S (fp32, 256 cols)  ->  P0 packed bf16 [0,64) | O fp32 [64,128) | P1 packed bf16 [128,192)
// The two softmax groups split the COLUMNS (keys) of one tile: group h owns keys [128h, 128h+128). While they work on
// tile u, the tensor core computes S of tile u+1 into the other buffer and O of tile u-1 behind them; the epilogue of
// tile u-1 is interleaved in the middle of tile u's softmax so its buffer is free again in time for S of tile u+1.
__global__ void __launch_bounds__(kThreadsF, 1)
attention_fa_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                    const __grid_constant__ CUtensorMap tm_const, int n_const, int t_live, int n_items,
                    uint16_t* __restrict__ out, int out_fp16, int dbg, long long* trace) {
#define ATT_TRACE(slot) do { if (trace != nullptr && blockIdx.x == 0 && lane == 0 && (slot) < 256) trace[(slot)] = clock64(); } while (0)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* xchg = reinterpret_cast<float*>(smem + 2 * kStageBytesF);  // [2 buffers][2 kinds: max, sum][2 halves][128 rows]
  uint8_t* out_stage = smem + 2 * kStageBytesF + kXchgBytes;  // [8 warps][32 rows][64 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kStageBytesF + kXchgBytes + kOutStageBytes);
  uint64_t* qk_full = bars + 0;    // [2 stages] TMA -> MMA
  uint64_t* v_full = bars + 2;     // [2 stages]
  uint64_t* qk_empty = bars + 4;   // [2 stages] MMA (commit) -> TMA
  uint64_t* v_empty = bars + 6;    // [2 stages]
  uint64_t* s_full = bars + 8;     // [2 buffers] MMA (commit) -> softmax groups
  uint64_t* p_ready = bars + 10;   // [2 buffers] softmax groups (8 warps) -> MMA
  uint64_t* o_full = bars + 12;    // [2 buffers] MMA (commit) -> softmax groups
  uint64_t* buf_free = bars + 14;  // [2 buffers] softmax groups (8 warps) -> MMA
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Tk = n_const + t_live;
  const int n_qt = (t_live + 127) >> 7;  // 128-query tiles per item (1 or 2)
  const int n_local = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int total_tiles = n_local * n_qt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_const);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qk_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&qk_empty[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 8); mbar_init(&o_full[i], 1); mbar_init(&buf_free[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();  // QKV of this layer comes from the previous kernel
  if (warp == 3) ATT_TRACE(0);

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    for (int it = 0; it < n_local; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int s = it & 1, ph = (it >> 1) & 1;
      const int head = item % 12, win = item / 12;
      const int row_base = win * t_live;
      uint8_t* sQ = smem + s * kStageBytesF;
      uint8_t* sK = sQ + kQBytes;
      uint8_t* sV = sK + kKVBytes;
      mbar_wait(&qk_empty[s], ph ^ 1);
      if (lane == 0) {
        if (dbg & 64) { mbar_arrive(&qk_full[s]); } else {
        mbar_arrive_expect_tx(&qk_full[s], kQBytes + kKVBytes);
        tma_load_2d(sQ, &tm_q, &qk_full[s], head * 64, row_base);
        if (n_const > 0) tma_load_2d(sK, &tm_const, &qk_full[s], 768 + head * 64, 0);
        tma_load_2d(sK + n_const * 128, &tm_kv, &qk_full[s], 768 + head * 64, row_base);
        }
      }
      __syncwarp();
      mbar_wait(&v_empty[s], ph ^ 1);
      if (lane == 0) {
        if (dbg & 64) { mbar_arrive(&v_full[s]); } else {
        mbar_arrive_expect_tx(&v_full[s], kKVBytes);
        if (n_const > 0) tma_load_2d(sV, &tm_const, &v_full[s], 1536 + head * 64, 0);
        tma_load_2d(sV + n_const * 128, &tm_kv, &v_full[s], 1536 + head * 64, row_base);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ------------------------------------
    constexpr uint32_t idesc_s = idesc_f(128, 256, false);
    // experiment (timing only, results garbage): other N for the P.V MMAs
    const uint32_t idesc_o = (dbg & 128) ? idesc_f(128, 128, true) : (dbg & 256) ? idesc_f(128, 16, true) : idesc_f(128, 64, true);
    const int k_steps = (Tk + 15) >> 4;
    // S of tile u: Q_t K^T into buffer u & 1
    auto issue_s = [&](int u) {
      const int it = u / n_qt, t = u - it * n_qt;
      const int s = it & 1, ph = (it >> 1) & 1;
      const int b = u & 1, j = u >> 1;
      const uint32_t q_addr = smem_u32(smem + s * kStageBytesF) + t * kQTileBytes;
      const uint32_t k_addr = smem_u32(smem + s * kStageBytesF) + kQBytes;
      mbar_wait(&qk_full[s], ph);
      ATT_TRACE(8 + 8 * u + 0);
      mbar_wait(&buf_free[b], (j & 1) ^ 1);  // the tile that used this buffer two tiles ago has been written out
      ATT_TRACE(8 + 8 * u + 1);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < ((dbg & 16) ? 0 : 4); ++k)
          umma_bf16_ss(tmem_base + b * 256, umma_desc_sw128_kmajor(q_addr + k * 32),
                       umma_desc_sw128_kmajor(k_addr + k * 32), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[b]);
        if (t == n_qt - 1) umma_commit(&qk_empty[s]);  // Q / K of this stage are dead after the item's last S
      }
      __syncwarp();
    };
    if (total_tiles > 0) issue_s(0);
    if (total_tiles > 1) issue_s(1);
    for (int u = 0; u < total_tiles; ++u) {
      const int it = u / n_qt, t = u - it * n_qt;
      const int s = it & 1, ph = (it >> 1) & 1;
      const int b = u & 1, j = u >> 1;
      const uint32_t v_addr = smem_u32(smem + s * kStageBytesF) + kQBytes + kKVBytes;
      mbar_wait(&v_full[s], ph);
      mbar_wait(&p_ready[b], j & 1);
      ATT_TRACE(8 + 8 * u + 2);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t buf = tmem_base + b * 256;
        for (int ks = 0; ks < ((dbg & 8) ? 0 : k_steps); ++ks) {
          // keys [0,128): packed P at columns [0,64); keys [128,256): packed P at columns [128,192)
          const uint32_t p_tmem = buf + (ks < 8 ? ks * 8 : 128 + (ks - 8) * 8);
          umma_bf16_ts(buf + 64, p_tmem, desc_sw128_mn(v_addr + ks * 2048), idesc_o, ks != 0 ? 1u : 0u);
        }
        umma_commit(&o_full[b]);
        if (t == n_qt - 1) umma_commit(&v_empty[s]);
      }
      __syncwarp();
      if (trace != nullptr && blockIdx.x == 0) {  // experiment: when does the PV chain retire?
        mbar_wait(&o_full[b], j & 1);
        ATT_TRACE(128 + 8 * u + 4);
      }
      if (u + 2 < total_tiles) issue_s(u + 2);
    }
  } else if (warp >= 4) {
    // ------------------------------------ softmax groups ------------------------------------
    const int h = (warp - 4) >> 2;  // column half owned by this group: keys [128h, 128h + 128)
    const int q = warp & 3;         // TMEM lane quadrant
    const int r = q * 32 + lane;    // row inside the tile
    const float kScale = 0.125f * 1.4426950408889634f;
    const int keys_mine = min(max(Tk - 128 * h, 0), 128);  // real keys in this half
    const int n_chunks = (keys_mine + 31) >> 5;            // 0..4
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    // O of tile u (this group's 32 of the 64 output dims) -> global, then release the buffer
    auto epilogue = [&](int u) {
      const int it = u / n_qt, t = u - it * n_qt;
      const int item = blockIdx.x + it * gridDim.x;
      const int head = item % 12, win = item / 12;
      const int b = u & 1, j = u >> 1;
      const int q_row = t * 128 + r;
      const bool active = t * 128 + q * 32 < t_live;
      mbar_wait(&o_full[b], j & 1);
      if (warp == 4) ATT_TRACE(8 + 8 * u + 5);
      tc_fence_after();
      if (active) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(lane_base + b * 256 + 64 + h * 32, o);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&buf_free[b]);
        if (!(dbg & 32)) {
          // thread = row holds 64 B of its output row; a direct store would touch 32 rows with 16 B each. Stage the
          // warp's 32 x 64 B in smem (slot ^= (row >> 1) & 3: conflict-free both ways) and write 8 complete 64 B row
          // segments per instruction instead.
          const float* sums = xchg + (b * 2 + 1) * 256;  // [2 halves][128 rows]
          const float inv = 1.0f / (sums[r] + sums[128 + r]);
          uint8_t* stg = out_stage + (warp - 4) * (32 * 64);
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((g ^ sw) << 4)) =
                make_uint4(pack16x2(__uint_as_float(o[8 * g + 0]) * inv, __uint_as_float(o[8 * g + 1]) * inv, out_fp16),
                           pack16x2(__uint_as_float(o[8 * g + 2]) * inv, __uint_as_float(o[8 * g + 3]) * inv, out_fp16),
                           pack16x2(__uint_as_float(o[8 * g + 4]) * inv, __uint_as_float(o[8 * g + 5]) * inv, out_fp16),
                           pack16x2(__uint_as_float(o[8 * g + 6]) * inv, __uint_as_float(o[8 * g + 7]) * inv, out_fp16));
          __syncwarp();
          const int slot = lane & 3, rsub = lane >> 2;
          const int row0 = t * 128 + q * 32;
          uint16_t* obase = out + (static_cast<int64_t>(win) * t_live + row0) * 768 + head * 64 + h * 32 + slot * 8;
#pragma unroll
          for (int it4 = 0; it4 < 4; ++it4) {
            const int rr = it4 * 8 + rsub;
            const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((slot ^ ((rr >> 1) & 3)) << 4));
            if (row0 + rr < t_live) *reinterpret_cast<uint4*>(obase + static_cast<int64_t>(rr) * 768) = v;
          }
          __syncwarp();
        }
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&buf_free[b]);
      }
    };

    for (int u = 0; u < total_tiles; ++u) {
      const int it = u / n_qt, t = u - it * n_qt;
      const int b = u & 1, j = u >> 1;
      const bool active = (t * 128 + q * 32 < t_live) && n_chunks > 0;
      const uint32_t s_row = lane_base + b * 256 + h * 128;  // this group's S columns; its P overlays their first half
      float* maxs = xchg + (b * 2 + 0) * 256;
      float* sums = xchg + (b * 2 + 1) * 256;
      float row_sum = 0.f;
      uint32_t va[32], vb[32];
      if (warp == 4) ATT_TRACE(8 + 8 * u + 6);
      mbar_wait(&s_full[b], j & 1);
      if (warp == 4) ATT_TRACE(8 + 8 * u + 3);
      tc_fence_after();
      // Reference maximum (softmax is shift invariant; it only has to keep exp2 in range): max over the first 32 keys of
      // each half, combined across the two groups through smem. The exponent is clamped at +120.
      float mx = -INFINITY;
      if (active) {
        if (!(dbg & 2)) tmem_ld_32x32b_x32(s_row, va);
        tmem_ld_wait();
        chunk_max(va, keys_mine, mx);
      }
      maxs[h * 128 + r] = mx;
      if (warp == 4) ATT_TRACE(128 + 8 * u + 0);
      named_bar_sync(1, 256);
      if (warp == 4) ATT_TRACE(128 + 8 * u + 1);
      const float m_scaled = fmaxf(mx, maxs[(1 - h) * 128 + r]) * kScale;
      if (active) {
        if (n_chunks > 1) { if (!(dbg & 2)) tmem_ld_32x32b_x32(s_row + 32, vb); }
        chunk_exp_store(va, keys_mine, kScale, m_scaled, row_sum, s_row, dbg);
        tmem_ld_wait();
        if (n_chunks > 1) {
          if (n_chunks > 2) { if (!(dbg & 2)) tmem_ld_32x32b_x32(s_row + 64, va); }
          chunk_exp_store(vb, keys_mine - 32, kScale, m_scaled, row_sum, s_row + 16, dbg);
          tmem_ld_wait();
        }
      }
      if (warp == 4) ATT_TRACE(128 + 8 * u + 2);
      if (u >= 1) epilogue(u - 1);  // mid-softmax: frees the other buffer in time for S of tile u + 1
      if (warp == 4) ATT_TRACE(128 + 8 * u + 3);
      if (active) {
        if (n_chunks > 2) {
          if (n_chunks > 3) { if (!(dbg & 2)) tmem_ld_32x32b_x32(s_row + 96, vb); }
          chunk_exp_store(va, keys_mine - 64, kScale, m_scaled, row_sum, s_row + 32, dbg);
          tmem_ld_wait();
          if (n_chunks > 3) chunk_exp_store(vb, keys_mine - 96, kScale, m_scaled, row_sum, s_row + 48, dbg);
        }
        tmem_st_wait();
      }
      sums[h * 128 + r] = row_sum;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[b]);
      if (warp == 4) ATT_TRACE(8 + 8 * u + 4);
    }
    if (total_tiles > 0) epilogue(total_tiles - 1);
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
    static int dbg = getenv("CLIPEBC_ATTN_DBG") ? atoi(getenv("CLIPEBC_ATTN_DBG")) : 0;  // experiment knob
    static const bool trace_env = getenv("CLIPEBC_ATTN_TRACE") != nullptr;  // experiment: time line of CTA 0
    static long long* trace_dev = nullptr;
    if (trace_env) {
      if (!trace_dev) cudaMalloc(&trace_dev, 256 * sizeof(long long));
      cudaMemsetAsync(trace_dev, 0, 256 * sizeof(long long), stream);
    }
    cudaError_t le = launch_pdl(attention_fa_kernel, dim3(grid), dim3(kThreadsF), kSmemF, stream, 1, tq, tkv, tc, n_const,
                                t_live, n_items, static_cast<uint16_t*>(out), out_fp16, dbg, trace_env ? trace_dev : nullptr);
    if (le != cudaSuccess) return cudaGetErrorString(le);
    if (trace_env) {
      long long h[256];
      cudaStreamSynchronize(stream);
      cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost);
      const long long t0 = h[0];
      printf("[attn trace] columns: qk_full buf_free(S issued) p_ready(PV issued) | softmax: wait_s s_full p_done o_full\n");
      for (int u = 0; u < 10 && h[8 + 8 * u + 3]; ++u)
        printf("  tile%d: %6lld %6lld %6lld (pv done %6lld) | %6lld %6lld %6lld %6lld | max_written %6lld exchanged %6lld half %6lld epi_done %6lld\n",
               u, h[8 + 8 * u] - t0, h[8 + 8 * u + 1] - t0, h[8 + 8 * u + 2] - t0, h[128 + 8 * u + 4] - t0, h[8 + 8 * u + 6] - t0,
               h[8 + 8 * u + 3] - t0, h[8 + 8 * u + 4] - t0, h[8 + 8 * u + 5] - t0, h[128 + 8 * u] - t0, h[128 + 8 * u + 1] - t0,
               h[128 + 8 * u + 2] - t0, h[128 + 8 * u + 3] - t0);
    }
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace cebc
