// Persistent tcgen05 attention for windows of 257..320 keys (ViT-L/14 at 224 x 224: 257 live tokens + 32 prompts = 289):
// softmax(q k^T / 8) v per (window, head), 64-dim heads. The two-chain kernel of attention_pp.cu with the keys of a tile taken
// in TWO blocks inside the chain's 256-column TMEM buffer. Replaces nn.MultiheadAttention -> F.scaled_dot_product_attention
// (/root/reference/models/clip/_clip/blocks.py:25,35-37) for the backbone of /root/reference/models/clip/model.py:16-24 whose
// sequence does not fit one 256-key tile; deep-VPT constant prompt keys as in attention_pp.cu (models/clip/model.py:164-183).
//
// Why two blocks work without a running maximum: the softmax of attention_pp.cu is single pass with a shift-invariant
// reference maximum (the first 32 keys of the row, exponent clamped at +120), so probabilities of different key blocks share
// one scale and O = P_A V_A + P_B V_B needs no correction. Per 128-query tile and chain (TMEM columns of the chain's buffer):
//
//   S_A = Q K[0:256)^T  -> [0, 256)                      MMA warp
//   softmax A: P_A bf16 -> [0, 128)                      softmax group (thread = query row)
//   S_B = Q K[256:320)^T -> [192, 256)   (consumed S_A)  MMA warp, issued BEFORE  O = P_A V[0:256) -> [128, 192)
//   softmax B: P_B bf16 -> [192, 224)                    overlaps the 16 MMAs of P_A V_A
//   O += P_B V[256:Tk)                                   MMA warp
//   O / rowsum -> global                                 softmax group
//
// Shared memory: K and V of an item are 320 rows each (80 KB per stage, two stages), so the queries no longer fit next to
// them; they travel per 128-query tile through their own three-slot ring (16 KB each). 224 KB + staging in total.
//   warp 0        TMA producer (K: constant prompt rows first, then the live rows in two boxes of (320 - n_const) / 2 rows;
//                 Q per tile; V)
//   warps 1, 3    MMA issuers of chain 0 / 1;  warp 2: TMEM allocator;  warps 4-7 / 8-11: softmax + output of chain 0 / 1
#include "common.cuh"
#include "kernels.h"

namespace cebc {

namespace {

constexpr int kThreadsL = 384;
constexpr int kKeysL = 320;                  // key rows of a K / V tile
constexpr int kKeysA = 256;                  // first key block
constexpr int kQTileBytesL = 128 * 128;      // one 128-query tile
constexpr int kKVBytesL = kKeysL * 128;      // 40 KB
constexpr int kStageBytesL = 2 * kKVBytesL;  // K | V
constexpr int kQSlotsL = 3;
constexpr int kMaxQtL = 3;                   // 128-query tiles per item (t_live <= 384)
constexpr int kOutStageBytesL = 8 * 32 * 64;
constexpr int kSmemL = 2 * kStageBytesL + kQSlotsL * kQTileBytesL + kOutStageBytesL + 1024 /*align*/ + 256 /*barriers*/;
static_assert(kSmemL <= 227 * 1024, "shared memory of the long-sequence attention kernel");

__device__ __forceinline__ uint64_t desc_sw128_mn_l(uint32_t smem_addr_bytes) {
  // MN-major operand in 128B-swizzled rows; SBO = 1024 B between 8-key groups; LBO unused for N = 64.
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr_bytes & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(kKVBytesL >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_l(int M, int N, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_ts_l(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
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
__device__ __forceinline__ void tmem_st_x16_l(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait_l() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f_l(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// p_j = exp2(min(s_j * scale - m_scaled, 120)) for the `lim` real keys of a 32-key chunk (0 beyond), fp32 row sum,
// P as packed bf16 pairs into TMEM over S columns that have already been consumed.
__device__ __forceinline__ void chunk_exp_store_l(const uint32_t (&v)[32], int lim, float scale, float m_scaled,
                                                  float& row_sum, uint32_t p_taddr) {
  uint32_t pk[16];
  if (lim >= 32) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float p0 = ex2f_l(fminf(__uint_as_float(v[2 * j]) * scale - m_scaled, 120.f));
      const float p1 = ex2f_l(fminf(__uint_as_float(v[2 * j + 1]) * scale - m_scaled, 120.f));
      row_sum += p0 + p1;
      pk[j] = pack_bf16x2(p0, p1);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float p0 = 0.f, p1 = 0.f;
      if (2 * j < lim) p0 = ex2f_l(fminf(__uint_as_float(v[2 * j]) * scale - m_scaled, 120.f));
      if (2 * j + 1 < lim) p1 = ex2f_l(fminf(__uint_as_float(v[2 * j + 1]) * scale - m_scaled, 120.f));
      row_sum += p0 + p1;
      pk[j] = pack_bf16x2(p0, p1);
    }
  }
  tmem_st_x16_l(p_taddr, pk);
}

__global__ void __launch_bounds__(kThreadsL, 1)
attention_ppl_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     const __grid_constant__ CUtensorMap tm_const, int n_const, int t_live, int n_items, int heads,
                     uint16_t* __restrict__ out, int out_fp16) {
  const int width = heads * 64;  // q | k | v thirds of a qkv row; row pitch of the output
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_ring = smem + 2 * kStageBytesL;                  // [3 slots][128 rows][128 B]
  uint8_t* out_stage = q_ring + kQSlotsL * kQTileBytesL;      // [8 warps][32 rows][64 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + kOutStageBytesL);
  uint64_t* k_full = bars + 0;     // [2 stages] TMA -> MMA
  uint64_t* v_full = bars + 2;     // [2 stages]
  uint64_t* k_empty = bars + 4;    // [2 stages] MMA (commit behind S_B of every tile of the item) -> TMA
  uint64_t* v_empty = bars + 6;    // [2 stages] MMA (commit behind P_B V_B of every tile of the item) -> TMA
  uint64_t* q_full = bars + 8;     // [3 slots] TMA -> MMA
  uint64_t* q_empty = bars + 11;   // [3 slots] MMA (commit behind S_B of the tile) -> TMA
  uint64_t* s_full = bars + 14;    // [2 chains] MMA (commit) -> softmax group: S_A
  uint64_t* p_ready = bars + 16;   // [2 chains] softmax group (4 warps) -> MMA: P_A written, S_A consumed
  uint64_t* sb_full = bars + 18;   // [2 chains] MMA (commit) -> softmax group: S_B
  uint64_t* pb_ready = bars + 20;  // [2 chains] softmax group -> MMA: P_B written
  uint64_t* o_full = bars + 22;    // [2 chains] MMA (commit) -> softmax group
  uint64_t* buf_free = bars + 24;  // [2 chains] softmax group (4 warps) -> MMA
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Tk = n_const + t_live;       // 257..320
  const int n_b = Tk - kKeysA;           // keys of the second block, 1..64
  const int half_live = (kKeysL - n_const) >> 1;  // rows of one live-key box
  const int n_qt = (t_live + 127) >> 7;  // 128-query tiles per item (1..3)
  // Work is dealt in 128-query tiles, a contiguous range per CTA; consecutive tiles alternate between the chains; the tiles
  // of an item share one K / V stage, an item cut by a range boundary is loaded by both CTAs.
  const int64_t n_tiles_all = static_cast<int64_t>(n_items) * n_qt;
  const int g0 = static_cast<int>(n_tiles_all * blockIdx.x / gridDim.x);
  const int g1 = static_cast<int>(n_tiles_all * (blockIdx.x + 1) / gridDim.x);
  const int total_tiles = g1 - g0;
  const int item0 = g0 / n_qt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    if (n_const > 0) tma_prefetch_desc(&tm_const);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], kMaxQtL); mbar_init(&v_empty[i], kMaxQtL);  // one commit per tile slot of the item
      mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&sb_full[i], 1); mbar_init(&pb_ready[i], 4);
      mbar_init(&o_full[i], 1); mbar_init(&buf_free[i], 4);
    }
    for (int i = 0; i < kQSlotsL; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();  // QKV of this layer comes from the previous kernel

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    int prev_item = -1;
    for (int u = 0; u < total_tiles; ++u) {
      const int g = g0 + u;
      const int item = g / n_qt, t = g - item * n_qt;
      const int it = item - item0;
      const int s = it & 1, ph = (it >> 1) & 1;
      const int head = item % heads, win = item / heads;
      const int row_base = win * t_live;
      const bool first = item != prev_item;  // first tile of the item in this CTA: its K / V have to be fetched
      uint8_t* sK = smem + s * kStageBytesL;
      uint8_t* sV = sK + kKVBytesL;
      if (first) {
        mbar_wait(&k_empty[s], ph ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(&k_full[s], kKVBytesL);
          if (n_const > 0) tma_load_2d(sK, &tm_const, &k_full[s], width + head * 64, 0);
          tma_load_2d(sK + n_const * 128, &tm_kv, &k_full[s], width + head * 64, row_base);
          tma_load_2d(sK + (n_const + half_live) * 128, &tm_kv, &k_full[s], width + head * 64, row_base + half_live);
        }
        __syncwarp();
      }
      const int slot = u % kQSlotsL, qph = (u / kQSlotsL) & 1;
      mbar_wait(&q_empty[slot], qph ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&q_full[slot], kQTileBytesL);
        tma_load_2d(q_ring + slot * kQTileBytesL, &tm_q, &q_full[slot], head * 64, row_base + t * 128);
      }
      __syncwarp();
      if (first) {
        mbar_wait(&v_empty[s], ph ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(&v_full[s], kKVBytesL);
          if (n_const > 0) tma_load_2d(sV, &tm_const, &v_full[s], 2 * width + head * 64, 0);
          tma_load_2d(sV + n_const * 128, &tm_kv, &v_full[s], 2 * width + head * 64, row_base);
          tma_load_2d(sV + (n_const + half_live) * 128, &tm_kv, &v_full[s], 2 * width + head * 64, row_base + half_live);
        }
        __syncwarp();
        prev_item = item;
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ------------------------------------ MMA issuer of chain b ------------------------------------
    const int b = warp >> 1;
    constexpr uint32_t idesc_sa = idesc_l(128, 256, false);
    constexpr uint32_t idesc_sb = idesc_l(128, 64, false);
    constexpr uint32_t idesc_o = idesc_l(128, 64, true);
    const int k_steps_b = (n_b + 15) >> 4;  // 1..4
    const uint32_t buf = tmem_base + b * 256;
    // De-phase the chains (attention_pp.cu): chain 1 starts when chain 0 has finished its first softmax.
    if (b == 1 && total_tiles > 1) mbar_wait(&p_ready[0], 0);
    int k = 0;
    for (int u = b; u < total_tiles; u += 2, ++k) {
      const int g = g0 + u;
      const int item = g / n_qt;
      const int it = item - item0;
      const int s = it & 1, ph = (it >> 1) & 1;
      const int slot = u % kQSlotsL, qph = (u / kQSlotsL) & 1;
      // a K / V stage is released by kMaxQtL commits: one per tile of the item this CTA processes, the missing ones (items
      // of fewer tiles, items cut by the range boundary) from the first of them
      const int lo = item * n_qt > g0 ? item * n_qt : g0, hi = (item + 1) * n_qt < g1 ? (item + 1) * n_qt : g1;
      const int n_commits = 1 + (g == lo ? kMaxQtL - (hi - lo) : 0);
      const uint32_t q_addr = smem_u32(q_ring + slot * kQTileBytesL);
      const uint32_t k_addr = smem_u32(smem + s * kStageBytesL);
      const uint32_t v_addr = k_addr + kKVBytesL;
      // S_A = Q K[0:256)^T once Q / K have landed and the previous tile of this chain has been read out
      mbar_wait(&k_full[s], ph);
      mbar_wait(&q_full[slot], qph);
      mbar_wait(&buf_free[b], (k & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_ss(buf, umma_desc_sw128_kmajor(q_addr + kk * 32), umma_desc_sw128_kmajor(k_addr + kk * 32), idesc_sa,
                       kk != 0 ? 1u : 0u);
        umma_commit(&s_full[b]);
      }
      __syncwarp();
      // S_B over the consumed columns [192, 256) as soon as the softmax group is through S_A, then O = P_A V_A behind it
      mbar_wait(&p_ready[b], k & 1);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_ss(buf + 192, umma_desc_sw128_kmajor(q_addr + kk * 32),
                       umma_desc_sw128_kmajor(k_addr + kKeysA * 128 + kk * 32), idesc_sb, kk != 0 ? 1u : 0u);
        umma_commit(&sb_full[b]);
        umma_commit(&q_empty[slot]);  // the tile's queries are dead
        for (int c = 0; c < n_commits; ++c) umma_commit(&k_empty[s]);
      }
      __syncwarp();
      mbar_wait(&v_full[s], ph);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll 4
        for (int ks = 0; ks < kKeysA / 16; ++ks)
          umma_bf16_ts_l(buf + 128, buf + ks * 8, desc_sw128_mn_l(v_addr + ks * 2048), idesc_o, ks != 0 ? 1u : 0u);
      }
      __syncwarp();
      // O += P_B V_B once the softmax group has written P_B
      mbar_wait(&pb_ready[b], k & 1);
      tc_fence_after();
      if (lane == 0) {
        for (int ks = 0; ks < k_steps_b; ++ks)
          umma_bf16_ts_l(buf + 128, buf + 192 + ks * 8, desc_sw128_mn_l(v_addr + (kKeysA / 16 + ks) * 2048), idesc_o, 1u);
        umma_commit(&o_full[b]);
        for (int c = 0; c < n_commits; ++c) umma_commit(&v_empty[s]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------ softmax + output of chain b ------------------------------------
    const int b = (warp - 4) >> 2;
    const int q = warp & 3;       // TMEM lane quadrant
    const float kScale = 0.125f * 1.4426950408889634f;
    const uint32_t row_base_t = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * 256;
    uint8_t* stg = out_stage + (warp - 4) * (32 * 64);
    int k = 0;
    for (int u = b; u < total_tiles; u += 2, ++k) {
      const int g = g0 + u;
      const int item = g / n_qt, t = g - item * n_qt;
      const int head = item % heads, win = item / heads;
      const int row0 = t * 128 + q * 32;          // first row of this warp inside the window
      const bool active = row0 < t_live;          // warps whose 32 rows are all padding only keep the protocol going
      float row_sum = 0.f, m_scaled = 0.f;
      mbar_wait(&s_full[b], k & 1);
      tc_fence_after();
      if (active) {
        // Block A: 8 full chunks of 32 keys. Reference maximum = maximum over the first 32 keys, exponent clamped at +120.
        uint32_t va[32], vb[32];
        tmem_ld_32x32b_x32(row_base_t, va);
        tmem_ld_wait();
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(va[j]));
        m_scaled = mx * kScale;
#pragma unroll 1
        for (int c = 0; c < 8; c += 2) {
          tmem_ld_32x32b_x32(row_base_t + (c + 1) * 32, vb);
          chunk_exp_store_l(va, 32, kScale, m_scaled, row_sum, row_base_t + c * 16);
          tmem_ld_wait();
          if (c + 2 < 8) tmem_ld_32x32b_x32(row_base_t + (c + 2) * 32, va);
          chunk_exp_store_l(vb, 32, kScale, m_scaled, row_sum, row_base_t + (c + 1) * 16);
          tmem_ld_wait();
        }
        tmem_st_wait_l();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[b]);

      // Block B: the remaining n_b <= 64 keys, S_B at columns [192, 256), P_B over its consumed columns from 192 on
      mbar_wait(&sb_full[b], k & 1);
      tc_fence_after();
      if (active) {
        uint32_t va[32], vb[32];
        tmem_ld_32x32b_x32(row_base_t + 192, va);
        if (n_b > 32) tmem_ld_32x32b_x32(row_base_t + 224, vb);
        tmem_ld_wait();
        chunk_exp_store_l(va, n_b, kScale, m_scaled, row_sum, row_base_t + 192);
        if (n_b > 32) chunk_exp_store_l(vb, n_b - 32, kScale, m_scaled, row_sum, row_base_t + 208);
        tmem_st_wait_l();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pb_ready[b]);

      // O / rowsum -> 16-bit -> global, two halves of 32 dims through the warp's smem staging tile
      mbar_wait(&o_full[b], k & 1);
      tc_fence_after();
      if (active) {
        const float inv = 1.0f / row_sum;
        const int sw = (lane >> 1) & 3;
        const int slot = lane & 3, rsub = lane >> 2;
        uint32_t o0[32], o1[32];
        tmem_ld_32x32b_x32(row_base_t + 128, o0);
        tmem_ld_32x32b_x32(row_base_t + 160, o1);
        tmem_ld_wait();
        // O is in registers: hand the TMEM buffer back before the stores so the next S_A of this chain can start
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&buf_free[b]);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t(&o)[32] = hh == 0 ? o0 : o1;
#pragma unroll
          for (int gq = 0; gq < 4; ++gq)
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((gq ^ sw) << 4)) =
                make_uint4(pack16x2(__uint_as_float(o[8 * gq + 0]) * inv, __uint_as_float(o[8 * gq + 1]) * inv, out_fp16),
                           pack16x2(__uint_as_float(o[8 * gq + 2]) * inv, __uint_as_float(o[8 * gq + 3]) * inv, out_fp16),
                           pack16x2(__uint_as_float(o[8 * gq + 4]) * inv, __uint_as_float(o[8 * gq + 5]) * inv, out_fp16),
                           pack16x2(__uint_as_float(o[8 * gq + 6]) * inv, __uint_as_float(o[8 * gq + 7]) * inv, out_fp16));
          __syncwarp();
          uint16_t* obase = out + (static_cast<int64_t>(win) * t_live + row0) * width + head * 64 + hh * 32 + slot * 8;
#pragma unroll
          for (int it4 = 0; it4 < 4; ++it4) {
            const int rr = it4 * 8 + rsub;
            const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((slot ^ ((rr >> 1) & 3)) << 4));
            if (row0 + rr < t_live) *reinterpret_cast<uint4*>(obase + static_cast<int64_t>(rr) * width) = v;
          }
          __syncwarp();
        }
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&buf_free[b]);
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

typedef CUresult (*PFN_encodeTiledL)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_tmap_rows_l(CUtensorMap* map, const void* base, int64_t rows, int box_rows, int ld) {
  static PFN_encodeTiledL enc = nullptr;
  if (!enc) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !ptr)
      return false;
    enc = reinterpret_cast<PFN_encodeTiledL>(ptr);
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(ld), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool attention_h64_ppl_takes(int n_const, int t_live) {
  const int tk = t_live + n_const;
  return tk > kKeysA && tk <= kKeysL && n_const >= 0 && n_const % 16 == 0 && t_live <= 128 * kMaxQtL;
}

// Contract: kernels.h. Needs 256 < t_live + n_const <= 320, n_const % 16 == 0 (two live-key boxes of whole swizzle atoms).
const char* attention_h64_ppl(cudaStream_t stream, const __nv_bfloat16* qkv, const __nv_bfloat16* const_kv, int n_const,
                              int n_win, int t_live, int heads, void* out, int out_fp16) {
  if (n_win <= 0 || t_live <= 0 || heads <= 0) return "attention: empty problem";
  if (n_const < 0 || (n_const > 0 && const_kv == nullptr)) return "attention: constant keys missing";
  if (!attention_h64_ppl_takes(n_const, t_live)) return "attention(ppl): needs 257..320 keys and a constant-key count that is a multiple of 16";
  static unsigned long long attr_done = 0;
  cudaError_t ea = ensure_dyn_smem(attention_ppl_kernel, kSmemL, &attr_done);
  if (ea != cudaSuccess) return cudaGetErrorString(ea);
  const int64_t rows = static_cast<int64_t>(n_win) * t_live;
  const int ld = 3 * 64 * heads;
  CUtensorMap tq, tkv, tc;
  if (!make_tmap_rows_l(&tq, qkv, rows, 128, ld)) return "attention: cuTensorMapEncodeTiled(q) failed";
  if (!make_tmap_rows_l(&tkv, qkv, rows, (kKeysL - n_const) / 2, ld)) return "attention: cuTensorMapEncodeTiled(kv) failed";
  if (n_const > 0) {
    if (!make_tmap_rows_l(&tc, const_kv, n_const, n_const, ld)) return "attention: cuTensorMapEncodeTiled(const) failed";
  } else {
    tc = tkv;
  }
  const int n_items = n_win * heads;
  // ViT-L/14: 257 = 2 * 128 + 1 rows, so every third tile holds ONE query and still costs its chain most of a period. Handing
  // that row to a warp-per-row kernel in a second launch was measured and dropped: each of its warps re-reads the item's whole
  // K / V (113 MB per launch at 96 windows), 137 us for both launches against 85 us for the three-tile form.
  const int64_t n_tiles_all = static_cast<int64_t>(n_items) * ((t_live + 127) / 128);
  const int grid = n_tiles_all < device_num_sms() ? static_cast<int>(n_tiles_all) : device_num_sms();
  {
    const double tk = t_live + n_const;
    LaunchScope scope(stream, "attention", 4.0 * n_items * t_live * tk * 64.0, 2.0 * n_win * t_live * 4.0 * 64.0 * heads);
    cudaError_t le = launch_pdl(attention_ppl_kernel, dim3(grid), dim3(kThreadsL), kSmemL, stream, 1, tq, tkv, tc, n_const,
                                t_live, n_items, heads, static_cast<uint16_t*>(out), out_fp16);
    if (le != cudaSuccess) return cudaGetErrorString(le);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace cebc
