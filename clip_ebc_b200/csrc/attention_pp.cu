// Persistent tcgen05 attention, "two chains" generation: softmax(q k^T / 8) v per (window, head), 64-dim heads, <= 256
// queries and <= 256 keys per item (229 for a 224x224 window with 32 prompt tokens).
// Replaces nn.MultiheadAttention -> F.scaled_dot_product_attention (/root/reference/models/clip/_clip/blocks.py:25,35-37);
// deep-VPT constant prompt keys/values are appended as in attention.cu (reference models/clip/model.py:164-183).
//
// Why two chains. With two softmax groups splitting the keys of ONE tile (the previous generation of this kernel, see
// DESIGN.md 4.2) both warps of every SM sub-partition execute the same phase at the same time: the ALU part, the MUFU
// (exp2) part and the TMEM / epilogue part of a tile add up instead of overlapping (measured with the parts switched off
// one at a time: 17 + 10 + 13 us per layer at 64 windows). Here a CTA runs two independent chains, one per TMEM buffer;
// chain b owns the 128-query tiles u = b, b + 2, ...:
//
//     S = Q_t K^T  ->  softmax (one thread = one query row, all keys)  ->  O = P V  ->  O / rowsum -> global
//       MMA warp b          softmax group b (4 warps)                   MMA warp b      softmax group b
//
// While chain b waits for its tensor work (the P.V chain alone is ~2200 cycles: 15 MMAs of 128x64x16 with A in TMEM),
// the other chain's softmax has the sub-partition's issue slots and MUFU pipe to itself. A thread owns a whole row, so
// there is no row-max / row-sum exchange and no CTA-level barrier in the steady state.
//
//   warp 0        TMA producer: Q [256 x 64], K, V [256 x 64] (constant prompt rows first) of item i+1 are fetched
//                 into the second smem stage while item i is being processed
//   warps 1, 3    MMA issuers of chain 0 / chain 1 (P is read straight from TMEM as the A operand, V is consumed in
//                 place as MN-major B)
//   warp 2        TMEM allocator (512 columns: one 256-column buffer per chain)
//   warps 4-7     softmax + output of chain 0, warps 8-11 of chain 1 (thread = query row = TMEM lane)
// TMEM buffer of a tile: S fp32 [0,256)  ->  P packed bf16 [0,128) | O fp32 [128,192); in the shipped two-block form (SPLIT = 1, see
// the kernel): P of keys [0,128) at [0,64) | O [64,128) | P of keys [128,256) at [128,192)
// Scores and probabilities never leave the SM: HBM traffic is Q, K, V in and O out.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace cebc {

namespace {

#ifdef CLIPEBC_ATTN_TRACE
// Build with `make EXTRA=-DCLIPEBC_ATTN_TRACE` for profiles/probes/attn_trace.py: clock64 time line of CTA 0, one slot per
// (warp, tile of the warp's loop, event); plain stores, no atomics. Compiled out otherwise.
__device__ long long g_tr[20 * 16 * 16];
__device__ __forceinline__ void tr_rec(int ev, int k) {
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && k < 16) g_tr[((threadIdx.x >> 5) * 16 + k) * 16 + (ev & 15)] = clock64();
}
#define TR(ev, k) tr_rec(ev, k)
#else
#define TR(ev, k)
#endif
constexpr int kThreadsP = 384;
constexpr int kSplitP = 1;  // 1: keys of a tile in two blocks of 128, P_A V_A under the second half of the softmax (0: one block)
constexpr int kPolyP = 0;  // share 1/kPolyP of the softmax exponentials on the FMA pipe; 0 = none (measured: no gain)
constexpr int kQTileBytesP = 128 * 128;    // one 128-query tile
constexpr int kQBytesP = 2 * kQTileBytesP; // 256 query rows x 64 dims, 16-bit
constexpr int kKVBytesP = 256 * 128;       // 256 key rows x 64 dims, 16-bit
constexpr int kStageBytesP = kQBytesP + 2 * kKVBytesP;  // 96 KB
constexpr int kOutStageBytesP = 8 * 32 * 64;            // per softmax warp: 32 rows x 64 B, XOR-swizzled
constexpr int kSmemP = 2 * kStageBytesP + kOutStageBytesP + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint64_t desc_sw128_mn_p(uint32_t smem_addr_bytes) {
  // MN-major operand in 128B-swizzled rows; SBO = 1024 B between 8-key groups; LBO unused for N = 64.
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr_bytes & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(kKVBytesP >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_p(int M, int N, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_ts_p(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
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
__device__ __forceinline__ void tmem_st_x16_p(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait_p() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f_p(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA pipe instead of the special-function pipe (4 lanes / clk and sub-partition, the busiest pipe of this
// kernel): x = n + f with n = floor(x) read off the low mantissa bits of x + 1.5 * 2^23 (rounded down), 2^f by a degree-3
// minimax polynomial on [0, 1) (relative error 8.8e-5, 40x below the bf16 rounding P gets next), 2^n added into the
// exponent field. x in [-126, 120] keeps 127 + n a valid exponent.
__device__ __forceinline__ float ex2_poly_p(float x) {
  x = fmaxf(x, -126.f);
  float xr;
  asm("add.rm.ftz.f32 %0, %1, %2;" : "=f"(xr) : "f"(x), "f"(12582912.f));
  const float f = x - (xr - 12582912.f);
  const float p = fmaf(fmaf(fmaf(0.077119089663028717f, f, 0.227564394474029541f), f, 0.695146143436431885f), f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}
// Element e of a chunk goes to the polynomial when POLY > 0 and e % POLY == POLY - 1 (POLY = 4: every fourth exponential).
template <int POLY>
__device__ __forceinline__ float ex2_sel_p(float x, int e) {
  if (POLY > 0 && (e % (POLY > 0 ? POLY : 1)) == POLY - 1) return ex2_poly_p(x);
  return ex2f_p(x);
}

// p_j = exp2(min(s_j * scale - m_scaled, 120)) for the `lim` real keys of a 32-key chunk (0 beyond), fp32 row sum,
// P as packed bf16 pairs into TMEM over S columns that have already been consumed.
template <int POLY>
__device__ __forceinline__ void chunk_exp_store_p(const uint32_t (&v)[32], int lim, float scale, float m_scaled,
                                                  float& row_sum, uint32_t p_taddr) {
  uint32_t pk[16];
  if (lim >= 32) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float p0 = ex2_sel_p<POLY>(fminf(__uint_as_float(v[2 * j]) * scale - m_scaled, 120.f), 2 * j);
      const float p1 = ex2_sel_p<POLY>(fminf(__uint_as_float(v[2 * j + 1]) * scale - m_scaled, 120.f), 2 * j + 1);
      row_sum += p0 + p1;
      pk[j] = pack_bf16x2(p0, p1);
    }
  } else {
    // partial last chunk: exp2 only for the real keys (lim is warp-uniform, so these are uniform branches)
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float p0 = 0.f, p1 = 0.f;
      if (2 * j < lim) p0 = ex2f_p(fminf(__uint_as_float(v[2 * j]) * scale - m_scaled, 120.f));
      if (2 * j + 1 < lim) p1 = ex2f_p(fminf(__uint_as_float(v[2 * j + 1]) * scale - m_scaled, 120.f));
      row_sum += p0 + p1;
      pk[j] = pack_bf16x2(p0, p1);
    }
  }
  tmem_st_x16_p(p_taddr, pk);
}

// SPLIT = 1: the keys of a tile are taken in two blocks of 128 (attention_ppl.cu does the same for 256 + 64): the softmax group
// announces P of keys [0, 128) half way, the MMA warp runs O = P_A V_A under the second half of the softmax, and only the k-steps
// of keys [128, Tk) are left behind it. TMEM buffer of a tile: S fp32 [0,256) -> P_A bf16 [0,64) | O fp32 [64,128) | P_B [128,192).
template <int POLY, int SPLIT>
__global__ void __launch_bounds__(kThreadsP, 1)
attention_pp_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                    const __grid_constant__ CUtensorMap tm_const, int n_const, int t_live, int n_items, int heads,
                    uint16_t* __restrict__ out, int out_fp16) {
  const int width = heads * 64;  // q | k | v thirds of a qkv row; row pitch of the output
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* out_stage = smem + 2 * kStageBytesP;  // [8 warps][32 rows][64 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kStageBytesP + kOutStageBytesP);
  uint64_t* qk_full = bars + 0;    // [2 stages] TMA -> MMA
  uint64_t* v_full = bars + 2;     // [2 stages]
  uint64_t* qk_empty = bars + 4;   // [2 stages] MMA (commit of every tile of the item) -> TMA
  uint64_t* v_empty = bars + 6;    // [2 stages]
  uint64_t* s_full = bars + 8;     // [2 chains] MMA (commit) -> softmax group
  uint64_t* p_ready = bars + 10;   // [2 chains] softmax group (4 warps) -> MMA
  uint64_t* o_full = bars + 12;    // [2 chains] MMA (commit) -> softmax group
  uint64_t* buf_free = bars + 14;  // [2 chains] softmax group (4 warps) -> MMA
  uint64_t* pb_ready = bars + 16;  // [2 chains] SPLIT: softmax group -> MMA, P of keys [128, Tk) written
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Tk = n_const + t_live;
  const int n_qt = (t_live + 127) >> 7;  // 128-query tiles per item (1 or 2)
  // Work is dealt in 128-query tiles, a contiguous range per CTA (1536 tiles on 148 CTAs = 10 or 11 each; dealing whole
  // items would give 10 or 12). The two tiles of an item are consecutive, so they land on the two chains and share one
  // K / V stage; an item cut by a range boundary is simply loaded by both CTAs.
  const int64_t n_tiles_all = static_cast<int64_t>(n_items) * n_qt;
  const int g0 = static_cast<int>(n_tiles_all * blockIdx.x / gridDim.x);
  const int g1 = static_cast<int>(n_tiles_all * (blockIdx.x + 1) / gridDim.x);
  const int total_tiles = g1 - g0;
  const int item0 = g0 / n_qt;
  const int n_local = total_tiles > 0 ? (g1 - 1) / n_qt - item0 + 1 : 0;  // items touched by this CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    if (n_const > 0) tma_prefetch_desc(&tm_const);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qk_full[i], 1); mbar_init(&v_full[i], 1);
      mbar_init(&qk_empty[i], 2); mbar_init(&v_empty[i], 2);  // two commits per item (see the MMA issuers)
      mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&o_full[i], 1); mbar_init(&buf_free[i], 4);
      mbar_init(&pb_ready[i], 4);
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

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    for (int it = 0; it < n_local; ++it) {
      const int item = item0 + it;
      const int s = it & 1, ph = (it >> 1) & 1;
      const int head = item % heads, win = item / heads;
      const int row_base = win * t_live;
      uint8_t* sQ = smem + s * kStageBytesP;
      uint8_t* sK = sQ + kQBytesP;
      uint8_t* sV = sK + kKVBytesP;
      TR(20, it);
      mbar_wait(&qk_empty[s], ph ^ 1);
      TR(21, it);
      if (lane == 0) {
        mbar_arrive_expect_tx(&qk_full[s], kQBytesP + kKVBytesP);
        tma_load_2d(sQ, &tm_q, &qk_full[s], head * 64, row_base);
        if (n_const > 0) tma_load_2d(sK, &tm_const, &qk_full[s], width + head * 64, 0);
        tma_load_2d(sK + n_const * 128, &tm_kv, &qk_full[s], width + head * 64, row_base);
      }
      __syncwarp();
      mbar_wait(&v_empty[s], ph ^ 1);
      TR(22, it);
      if (lane == 0) {
        mbar_arrive_expect_tx(&v_full[s], kKVBytesP);
        if (n_const > 0) tma_load_2d(sV, &tm_const, &v_full[s], 2 * width + head * 64, 0);
        tma_load_2d(sV + n_const * 128, &tm_kv, &v_full[s], 2 * width + head * 64, row_base);
      }
      __syncwarp();
    }
  } else if (warp == 1 || warp == 3) {
    // ------------------------------------ MMA issuer of chain b ------------------------------------
    const int b = warp >> 1;
    constexpr uint32_t idesc_s = idesc_p(128, 256, false);
    constexpr uint32_t idesc_o = idesc_p(128, 64, true);
    const int k_steps = (Tk + 15) >> 4;
    const uint32_t buf = tmem_base + b * 256;
    // De-phase the chains: left alone they start together and stay in lockstep (both softmax groups fight for the MUFU
    // pipe, then both wait for the tensor pipe). Chain 1 therefore starts when chain 0 has finished its first softmax;
    // from then on one chain's softmax runs while the other is in its P.V / output / next-S phase.
    if (b == 1 && total_tiles > 1) mbar_wait(SPLIT ? &pb_ready[0] : &p_ready[0], 0);
    constexpr uint32_t o_col = SPLIT ? 64 : 128;
    int k = 0;
    for (int u = b; u < total_tiles; u += 2, ++k) {
      const int g = g0 + u;
      const int item = g / n_qt, t = g - item * n_qt;
      const int it = item - item0;
      const int s = it & 1, ph = (it >> 1) & 1;
      // the stage of an item is released by two commits: one per tile, or both from here when this CTA only has one
      // tile of the item (single-tile windows, or an item cut by the range boundary)
      const int lo = item * n_qt > g0 ? item * n_qt : g0, hi = (item + 1) * n_qt < g1 ? (item + 1) * n_qt : g1;
      const bool sole = (hi - lo) == 1;
      const uint32_t stage_addr = smem_u32(smem + s * kStageBytesP);
      const uint32_t q_addr = stage_addr + t * kQTileBytesP;
      const uint32_t k_addr = stage_addr + kQBytesP;
      const uint32_t v_addr = k_addr + kKVBytesP;
      // S = Q_t K^T once Q / K have landed and the previous tile of this chain has been written out
      TR(0, k);
      mbar_wait(&qk_full[s], ph);
      TR(4, k);
      mbar_wait(&buf_free[b], (k & 1) ^ 1);
      TR(1, k);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_ss(buf, umma_desc_sw128_kmajor(q_addr + kk * 32), umma_desc_sw128_kmajor(k_addr + kk * 32), idesc_s,
                       kk != 0 ? 1u : 0u);
        umma_commit(&s_full[b]);
        umma_commit(&qk_empty[s]);  // Q / K of the stage are dead once every tile of the item has done this
        if (sole) umma_commit(&qk_empty[s]);
      }
      __syncwarp();
      // O = P V once the softmax group has written P
      mbar_wait(&v_full[s], ph);
      mbar_wait(&p_ready[b], k & 1);
      TR(2, k);
      tc_fence_after();
      if (SPLIT) {
        const int ks_a = k_steps < 8 ? k_steps : 8;
        if (lane == 0)
          for (int ks = 0; ks < ks_a; ++ks)
            umma_bf16_ts_p(buf + o_col, buf + ks * 8, desc_sw128_mn_p(v_addr + ks * 2048), idesc_o, ks != 0 ? 1u : 0u);
        __syncwarp();
        mbar_wait(&pb_ready[b], k & 1);
        tc_fence_after();
      }
      if (lane == 0) {
        for (int ks = SPLIT ? 8 : 0; ks < k_steps; ++ks)
          umma_bf16_ts_p(buf + o_col, buf + (SPLIT ? 64 : 0) + ks * 8, desc_sw128_mn_p(v_addr + ks * 2048), idesc_o,
                         (SPLIT || ks != 0) ? 1u : 0u);
        umma_commit(&o_full[b]);
        umma_commit(&v_empty[s]);
        if (sole) umma_commit(&v_empty[s]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------ softmax + output of chain b ------------------------------------
    const int b = (warp - 4) >> 2;
    const int q = warp & 3;       // TMEM lane quadrant
    const int r = q * 32 + lane;  // row inside the tile
    const float kScale = 0.125f * 1.4426950408889634f;
    const int n_chunks = (Tk + 31) >> 5;  // 1..8
    const uint32_t row_base_t = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * 256;
    uint8_t* stg = out_stage + (warp - 4) * (32 * 64);
    int k = 0;
    for (int u = b; u < total_tiles; u += 2, ++k) {
      const int g = g0 + u;
      const int item = g / n_qt, t = g - item * n_qt;
      const int head = item % heads, win = item / heads;
      const int row0 = t * 128 + q * 32;          // first row of this warp inside the window
      const bool active = row0 < t_live;          // warps whose 32 rows are all padding only keep the protocol going
      float row_sum = 0.f;
      bool pa_done = false;
      TR(10, k);
      mbar_wait(&s_full[b], k & 1);
      TR(11, k);
      tc_fence_after();
      if (active) {
        // Single pass over S. Softmax is shift invariant, so the reference maximum only has to keep exp2 in range: the
        // maximum over the first 32 keys is used and the exponent is clamped at +120.
        uint32_t va[32], vb[32];
        tmem_ld_32x32b_x32(row_base_t, va);
        tmem_ld_wait();
        float mx = -INFINITY;
        const int lim0 = Tk < 32 ? Tk : 32;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < lim0) mx = fmaxf(mx, __uint_as_float(va[j]));
        const float m_scaled = mx * kScale;
        // P of chunk c: packed pairs over consumed S columns -- [16c, 16c + 16), or from column 128 on for keys >= 128 (SPLIT)
        const uint32_t p_hi = SPLIT ? 64 : 0;
#pragma unroll 1
        for (int c = 0; c < n_chunks; c += 2) {
          if (c + 1 < n_chunks) tmem_ld_32x32b_x32(row_base_t + (c + 1) * 32, vb);
          chunk_exp_store_p<POLY>(va, Tk - c * 32, kScale, m_scaled, row_sum, row_base_t + (c >= 4 ? p_hi : 0) + c * 16);
          tmem_ld_wait();
          if (c + 1 < n_chunks) {
            if (c + 2 < n_chunks) tmem_ld_32x32b_x32(row_base_t + (c + 2) * 32, va);
            chunk_exp_store_p<POLY>(vb, Tk - (c + 1) * 32, kScale, m_scaled, row_sum,
                                    row_base_t + (c >= 4 ? p_hi : 0) + (c + 1) * 16);
            tmem_ld_wait();
          }
          if (SPLIT && c == 2) {  // keys [0, 128) are done: let the MMA warp start O = P_A V_A
            tmem_st_wait_p();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[b]);
            pa_done = true;
          }
        }
        tmem_st_wait_p();
      }
      TR(12, k);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!SPLIT || !pa_done) mbar_arrive(&p_ready[b]);
        if (SPLIT) mbar_arrive(&pb_ready[b]);
      }

      // O / rowsum -> 16-bit -> global, two halves of 32 dims through the warp's smem staging tile (thread = row holds
      // 64 B of its row; staged, one instruction writes 8 complete 64 B row segments)
      mbar_wait(&o_full[b], k & 1);
      TR(13, k);
      tc_fence_after();
      if (active) {
        const float inv = 1.0f / row_sum;
        const int sw = (lane >> 1) & 3;
        const int slot = lane & 3, rsub = lane >> 2;
        uint32_t o0[32], o1[32];
        tmem_ld_32x32b_x32(row_base_t + (SPLIT ? 64 : 128), o0);
        tmem_ld_32x32b_x32(row_base_t + (SPLIT ? 96 : 160), o1);
        tmem_ld_wait();
        // O is in registers: hand the TMEM buffer back before the stores so the next S of this chain can start
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&buf_free[b]);
        TR(14, k);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t(&o)[32] = hh == 0 ? o0 : o1;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((g ^ sw) << 4)) =
                make_uint4(pack16x2(__uint_as_float(o[8 * g + 0]) * inv, __uint_as_float(o[8 * g + 1]) * inv, out_fp16),
                           pack16x2(__uint_as_float(o[8 * g + 2]) * inv, __uint_as_float(o[8 * g + 3]) * inv, out_fp16),
                           pack16x2(__uint_as_float(o[8 * g + 4]) * inv, __uint_as_float(o[8 * g + 5]) * inv, out_fp16),
                           pack16x2(__uint_as_float(o[8 * g + 6]) * inv, __uint_as_float(o[8 * g + 7]) * inv, out_fp16));
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
      TR(15, k);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

typedef CUresult (*PFN_encodeTiledP)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_tmap_rows_p(CUtensorMap* map, const void* base, int64_t rows, int box_rows, int ld) {
  static PFN_encodeTiledP enc = nullptr;
  if (!enc) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !ptr)
      return false;
    enc = reinterpret_cast<PFN_encodeTiledP>(ptr);
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

// Contract: kernels.h. Needs t_live + n_const <= 256 and n_const % 8 == 0 (swizzle atom of the K / V tiles).
const char* attention_h64_pp(cudaStream_t stream, const __nv_bfloat16* qkv, const __nv_bfloat16* const_kv, int n_const,
                             int n_win, int t_live, int heads, void* out, int out_fp16) {
  if (n_win <= 0 || t_live <= 0 || heads <= 0) return "attention: empty problem";
  if (n_const < 0 || (n_const > 0 && const_kv == nullptr)) return "attention: constant keys missing";
  if (t_live + n_const > 256) return "attention: sequence longer than 256 keys is not supported by the tcgen05 kernel";
  if (n_const % 8 != 0) return "attention(pp): constant key count must be a multiple of 8";
  static unsigned long long attr_done = 0;
#ifdef CLIPEBC_ATTN_POLY_AB
  // A/B build only (make EXTRA=-DCLIPEBC_ATTN_POLY_AB): CLIPEBC_ATTN_POLY = 0 | 2 | 3 | 4 picks the share of exponentials
  // evaluated on the FMA pipe. The shipped library has the one instantiation below.
  static unsigned long long attr_done_ab[4] = {0, 0, 0, 0};
  static const int poly_env = [] { const char* e = getenv("CLIPEBC_ATTN_POLY"); return e ? atoi(e) : kPolyP; }();
  static unsigned long long attr_done_split = 0;
  static const int split_env = [] { const char* e = getenv("CLIPEBC_ATTN_SPLIT"); return e ? atoi(e) : kSplitP; }();
  auto kern = attention_pp_kernel<kPolyP, kSplitP>;
  unsigned long long* mask = &attr_done;
  if (split_env != kSplitP) {
    kern = attention_pp_kernel<kPolyP, 1 - kSplitP>; mask = &attr_done_split;
  } else if (poly_env != kPolyP) {
    if (poly_env == 0) { kern = attention_pp_kernel<0, kSplitP>; mask = &attr_done_ab[0]; }
    else if (poly_env == 2) { kern = attention_pp_kernel<2, kSplitP>; mask = &attr_done_ab[1]; }
    else if (poly_env == 3) { kern = attention_pp_kernel<3, kSplitP>; mask = &attr_done_ab[2]; }
    else if (poly_env == 4) { kern = attention_pp_kernel<4, kSplitP>; mask = &attr_done_ab[3]; }
  }
#else
  auto kern = attention_pp_kernel<kPolyP, kSplitP>;
  unsigned long long* mask = &attr_done;
#endif
  cudaError_t ea = ensure_dyn_smem(kern, kSmemP, mask);
  if (ea != cudaSuccess) return cudaGetErrorString(ea);
  const int64_t rows = static_cast<int64_t>(n_win) * t_live;
  const int ld = 3 * 64 * heads;
  CUtensorMap tq, tkv, tc;
  if (!make_tmap_rows_p(&tq, qkv, rows, 256, ld)) return "attention: cuTensorMapEncodeTiled(q) failed";
  if (!make_tmap_rows_p(&tkv, qkv, rows, 256 - n_const, ld)) return "attention: cuTensorMapEncodeTiled(kv) failed";
  if (n_const > 0) {
    if (!make_tmap_rows_p(&tc, const_kv, n_const, n_const, ld)) return "attention: cuTensorMapEncodeTiled(const) failed";
  } else {
    tc = tkv;
  }
  const int n_items = n_win * heads;
  const int64_t n_tiles_all = static_cast<int64_t>(n_items) * ((t_live + 127) / 128);
  const int grid = n_tiles_all < device_num_sms() ? static_cast<int>(n_tiles_all) : device_num_sms();
  {
    const double tk = t_live + n_const;
    LaunchScope scope(stream, "attention", 4.0 * n_items * t_live * tk * 64.0, 2.0 * n_win * t_live * 4.0 * 64.0 * heads);
    cudaError_t le = launch_pdl(kern, dim3(grid), dim3(kThreadsP), kSmemP, stream, 1, tq, tkv, tc, n_const,
                                t_live, n_items, heads, static_cast<uint16_t*>(out), out_fp16);
    if (le != cudaSuccess) return cudaGetErrorString(le);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

#ifdef CLIPEBC_ATTN_TRACE
extern "C" int clipebc_debug_attn_trace(long long* out, int cap) {
  const int n = 20 * 16 * 16;
  if (cap < n) return -1;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_tr, sizeof(long long) * n);
  static long long zeros[20 * 16 * 16];
  cudaMemcpyToSymbol(g_tr, zeros, sizeof(long long) * n);
  return n;
}
#endif

}  // namespace cebc
