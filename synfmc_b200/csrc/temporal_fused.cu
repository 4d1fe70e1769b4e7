// Fused temporal self-attention for the CameraAdapter motion module (C = 320, 8 heads x 40):
//     [q | k | v] = m W_qkv^T            per head, on tcgen05 (M = 128 tokens = 128/f sequences, N = 128, K = 320)
//     o = softmax(q k^T * scale) v        over the f frames of each latent position (block-diagonal 128 x 128 tile)
// in ONE kernel: the [token, 1088] q|k|v tensor of the un-fused chain (fmc_gemm_bf16 -> fmc_temporal_attn_bf16) never
// exists -- per (token tile, head) the projection lands in tensor memory and is turned into MMA operands on chip;
// only o[token, 320] goes back to HBM.
//
// Replaces to_q / to_k / to_v + head_to_batch_dim + baddbmm + softmax + bmm + batch_to_head_dim of the temporal
// processors (fmc/models/attention_processor.py:46-67 AttnProcessor, :259-281 PoseAdaptorAttnProcessor) as reached
// from TemporalSelfAttention.forward (fmc/models/motion_module.py:349-389).
//
// Work decomposition: persistent CTAs (one per SM) walk token tiles; the tile's input rows [128, 320] stay resident in
// shared memory (SWIZZLE_128B, five 64-column k-blocks, reloaded k-block by k-block behind the last head) while the
// per-head weight slices [128, 320] stream through a 6-stage TMA ring (one head = 5 stages, so the next head's weights
// are always in flight).  q and the probabilities P never touch shared memory: they are written back to TENSOR MEMORY
// as packed bf16 and consumed as the A operand of the score / PV MMAs (tcgen05.mma with A in TMEM), q in place over
// its fp32 accumulator columns, P in place over the scores (lane = row, column c = elements 2c, 2c+1; checked by
// profiles/microbench/mma_bench.cu).  Only k and v (B operands) are staged in shared memory.
//   warp 0       TMA producer (input tile, weight ring)
//   warp 1       tcgen05.mma issuer: per item m = (tile, head):  S(m), G(m+2), PV(m)  -- the projection of head m+2
//                keeps the tensor pipe busy while head m is in its softmax; one issuing thread = in-order execution,
//                which is what makes the in-place q / P operands safe against the next accumulator write
//   warp 2       TMEM allocator
//   warps 4-7    WG-A: q -> bf16 -> TMEM, k -> bf16 -> smem; block-diagonal softmax; P -> bf16 -> TMEM
//   warps 8-11   WG-B: v -> bf16 -> smem (plus a ones column so the PV MMA also yields the softmax row sum);
//                o epilogue (normalise, bf16, store)
// TMEM (512 columns): projection accumulators 2 x 128 | S / P 128 | O 2 x 48.
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fmc {

constexpr int TF_THREADS = 384;
constexpr int TF_C = 320;                 // channels
constexpr int TF_D = 40;                  // head width
constexpr int TF_DK = 48;                 // head width padded to the MMA K granularity
constexpr int TF_HEADS = 8;
constexpr int TF_HEAD_ROWS = 128;         // weight rows per head: q(40) | k(40) | v(40) | 8 zero rows
constexpr int TF_KB = TF_C / 64;          // k-blocks of the resident input tile
constexpr int TF_KBLK_BYTES = 128 * 128;  // [128 rows x 64 bf16] SWIZZLE_128B block
constexpr int TF_A_BYTES = TF_KB * TF_KBLK_BYTES;
constexpr int TF_W_STAGE_BYTES = TF_HEAD_ROWS * 128;  // [128 rows x 64 bf16]
constexpr int TF_W_STAGES = 6;
constexpr int TF_OFF_K = TF_A_BYTES;
constexpr int TF_OFF_V = TF_OFF_K + TF_KBLK_BYTES;
constexpr int TF_OFF_W = TF_OFF_V + TF_KBLK_BYTES;
constexpr int TF_SMEM_BYTES = TF_OFF_W + TF_W_STAGES * TF_W_STAGE_BYTES + 1024;
constexpr uint32_t TF_COL_G = 0;     // 2 x 128 projection accumulators (q bf16 is rewritten in place over columns 0..23)
constexpr uint32_t TF_COL_S = 256;   // 128 score columns (P bf16 is rewritten in place over columns 0..63)
constexpr uint32_t TF_COL_O = 384;   // 2 x 48 output accumulators
static_assert(TF_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(TF_W_STAGE_BYTES % 1024 == 0, "SW128 tiles stay 1024-byte aligned");

struct TfParams {
  int B, F, F_log2, HW;
  int seqs_per_tile;  // 128 / F
  int tiles_per_b;    // ceil(HW / seqs_per_tile)
  float scale_log2e;
  __nv_bfloat16* O;
  long long ldo;
  long long* timeline;  // diagnostics (fmc_debug_set_timeline): clock64() of pipeline events of CTA 0, or nullptr
};

// timeline[role][item][event], role 0 = MMA issuer, 1 = WG-A (warp 4), 2 = WG-B (warp 8), 3 = TMA producer
constexpr int TF_TL_ITEMS = 64, TF_TL_EVENTS = 8;
static long long* g_tf_timeline = nullptr;
#define TF_MARK(role, item, ev)                                                                              \
  do {                                                                                                       \
    if (p.timeline != nullptr && blockIdx.x == 0 && (item) < TF_TL_ITEMS)                                     \
      p.timeline[((role) * TF_TL_ITEMS + (item)) * TF_TL_EVENTS + (ev)] = clock64();                          \
  } while (0)

__device__ __forceinline__ float tf_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// D[tmem] (+)= A[tmem, packed bf16] * B[smem]^T
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// 40 fp32 values (a[0..31], b[0..7]) -> bf16 -> row `r` of a [128 x 64] SW128 smem tile; columns 40..47 (the padding of
// the head width to the MMA K granularity) become zero, or (1, 0, ..., 0) with `ones_col40`
__device__ __forceinline__ void tf_store40(const uint32_t (&a)[32], const uint32_t (&b)[8], uint32_t tile, int r,
                                           bool ones_col40) {
  const uint32_t row = tile + static_cast<uint32_t>(r) * 128u;
  const uint32_t sw = static_cast<uint32_t>(r & 7);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    st_shared_v4(row + ((static_cast<uint32_t>(j) ^ sw) << 4),
                 pack_bf16x2(__uint_as_float(a[8 * j]), __uint_as_float(a[8 * j + 1])),
                 pack_bf16x2(__uint_as_float(a[8 * j + 2]), __uint_as_float(a[8 * j + 3])),
                 pack_bf16x2(__uint_as_float(a[8 * j + 4]), __uint_as_float(a[8 * j + 5])),
                 pack_bf16x2(__uint_as_float(a[8 * j + 6]), __uint_as_float(a[8 * j + 7])));
  st_shared_v4(row + ((4u ^ sw) << 4), pack_bf16x2(__uint_as_float(b[0]), __uint_as_float(b[1])),
               pack_bf16x2(__uint_as_float(b[2]), __uint_as_float(b[3])),
               pack_bf16x2(__uint_as_float(b[4]), __uint_as_float(b[5])),
               pack_bf16x2(__uint_as_float(b[6]), __uint_as_float(b[7])));
  st_shared_v4(row + ((5u ^ sw) << 4), ones_col40 ? 0x00003F80u : 0u, 0u, 0u, 0u);
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

__global__ void __launch_bounds__(TF_THREADS, 1)
temporal_qkv_attn_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, TfParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full[TF_KB], a_free[TF_KB];
  __shared__ uint64_t w_full[TF_W_STAGES], w_empty[TF_W_STAGES];
  __shared__ uint64_t g_full[2], g_free[2];
  __shared__ uint64_t qk_ready, s_full, p_ready, v_ready;
  __shared__ uint64_t o_full[2], o_free[2];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sK = smem_base + TF_OFF_K;
  const uint32_t sV = smem_base + TF_OFF_V;
  const uint32_t sW = smem_base + TF_OFF_W;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Work items are (tile, head) pairs in tile-major order; every CTA takes one contiguous range, so the load is even
  // to 1 / heads of a tile (640 tiles on 148 SMs would otherwise cost 5 rounds for 4.3 rounds of work).  A range may
  // start or end in the middle of a tile: two CTAs then load the same input tile and write different heads of it.
  const int total_items = p.B * p.tiles_per_b * TF_HEADS;
  const int per_cta = (total_items + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int g0 = static_cast<int>(blockIdx.x) * per_cta;
  const int my_items = max(0, min(per_cta, total_items - g0));
  const int tile0 = g0 / TF_HEADS;
  const int my_tiles = my_items > 0 ? (g0 + my_items - 1) / TF_HEADS - tile0 + 1 : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TF_KB; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_free[s], 1);
    }
    for (int s = 0; s < TF_W_STAGES; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&g_full[s], 1);
      mbar_init(&g_free[s], 8);
      mbar_init(&o_full[s], 1);
      mbar_init(&o_free[s], 4);
    }
    mbar_init(&qk_ready, 4);
    mbar_init(&s_full, 1);
    mbar_init(&p_ready, 4);
    mbar_init(&v_ready, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();  // everything above (barriers, TMEM, descriptor prefetch) overlaps the previous kernel's tail

  // local item m -> global item g0 + m = (tile, head); `it` = index of the tile among this CTA's tiles
  auto head_of = [&](uint32_t m) -> int { return (g0 + static_cast<int>(m)) & (TF_HEADS - 1); };
  auto it_of = [&](uint32_t m) -> int { return (g0 + static_cast<int>(m)) / TF_HEADS - tile0; };
  auto first_of_tile = [&](uint32_t m) -> bool { return m == 0 || head_of(m) == 0; };
  auto last_of_tile = [&](uint32_t m) -> bool { return static_cast<int>(m) == my_items - 1 || head_of(m) == TF_HEADS - 1; };
  auto tile_coords = [&](int it, int& b, int& hw0) {
    const int tile = tile0 + it;
    b = tile / p.tiles_per_b;
    hw0 = (tile % p.tiles_per_b) * p.seqs_per_tile;
  };

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    if (elect_one()) {
      uint32_t nw = 0;
      for (int it = 0; it < my_tiles; ++it) {
        int b, hw0;
        tile_coords(it, b, hw0);
        for (int kb = 0; kb < TF_KB; ++kb) {
          // k-block kb of the previous tile is free once the projection of its last head has read it
          mbar_wait(&a_free[kb], (static_cast<uint32_t>(it) & 1u) ^ 1u);
          if (kb == 0) TF_MARK(3, it * TF_HEADS, 0);
          mbar_arrive_expect_tx(&a_full[kb], TF_KBLK_BYTES);
          for (int g = 0; g < p.seqs_per_tile; ++g) {
            // box = (64 columns, 1 position, F frames): the F rows of sequence (b, hw0 + g); positions past HW are
            // zero-filled by TMA
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(sA + kb * TF_KBLK_BYTES + g * p.F * 128), "l"(reinterpret_cast<uint64_t>(&tmX)),
                "r"(smem_u32(&a_full[kb])), "r"(kb * 64), "r"(hw0 + g), "r"(b * p.F)
                : "memory");
          }
        }
        const int h_begin = it == 0 ? (g0 & (TF_HEADS - 1)) : 0;
        const int h_end = it == my_tiles - 1 ? ((g0 + my_items - 1) & (TF_HEADS - 1)) + 1 : TF_HEADS;
        for (int h = h_begin; h < h_end; ++h) {
          for (int kb = 0; kb < TF_KB; ++kb, ++nw) {
            const uint32_t st = nw % TF_W_STAGES;
            mbar_wait(&w_empty[st], ((nw / TF_W_STAGES) & 1u) ^ 1u);
            mbar_arrive_expect_tx(&w_full[st], TF_W_STAGE_BYTES);
            tma_load_2d_a(sW + st * TF_W_STAGE_BYTES, &tmW, &w_full[st], kb * 64, h * TF_HEAD_ROWS);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ------------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, TF_HEAD_ROWS);
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
      constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(128, TF_DK);
      uint32_t nw = 0;
      // projection of item m into accumulator buffer m & 1
      auto issue_g = [&](uint32_t m) {
        const uint32_t buf = m & 1u;
        const bool first = first_of_tile(m), last = last_of_tile(m);
        const uint32_t it = static_cast<uint32_t>(it_of(m));
        TF_MARK(0, m, 0);
        mbar_wait(&g_free[buf], ((m >> 1) & 1u) ^ 1u);
        tc_fence_after_sync();
        for (int kb = 0; kb < TF_KB; ++kb, ++nw) {
          const uint32_t st = nw % TF_W_STAGES;
          if (first) mbar_wait(&a_full[kb], it & 1u);
          mbar_wait(&w_full[st], (nw / TF_W_STAGES) & 1u);
          tc_fence_after_sync();
          const uint64_t da = umma_desc_k_sw128(sA + kb * TF_KBLK_BYTES);
          const uint64_t db = umma_desc_k_sw128(sW + st * TF_W_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tmem_base + TF_COL_G + buf * TF_HEAD_ROWS, da + static_cast<uint64_t>(2 * k),
                         db + static_cast<uint64_t>(2 * k), idesc_g, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&w_empty[st]);
          if (last) umma_commit(&a_free[kb]);  // last head of the tile (for this CTA): k-block kb may be reloaded
        }
        umma_commit(&g_full[buf]);
        TF_MARK(0, m, 1);
      };
      // S(m) = q k^T: A = packed bf16 q in TMEM (in place over the first 24 accumulator columns), B = k in smem
      auto issue_s = [&](uint32_t m) {
        TF_MARK(0, m, 2);
        mbar_wait(&qk_ready, m & 1u);
        tc_fence_after_sync();
        const uint32_t tq = tmem_base + TF_COL_G + (m & 1u) * TF_HEAD_ROWS;
#pragma unroll
        for (int k = 0; k < TF_DK / 16; ++k)
          umma_bf16_ts(tmem_base + TF_COL_S, tq + 8 * k, umma_desc_k_sw128(sK + k * 32), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&s_full);
        TF_MARK(0, m, 3);
      };
      // O(m) = P v: A = packed bf16 P in TMEM (in place over the first 64 score columns), B = v in smem (MN-major)
      auto issue_pv = [&](uint32_t m) {
        const uint32_t buf = m & 1u;
        TF_MARK(0, m, 4);
        mbar_wait(&p_ready, m & 1u);
        mbar_wait(&v_ready, m & 1u);
        mbar_wait(&o_free[buf], ((m >> 1) & 1u) ^ 1u);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16_ts(tmem_base + TF_COL_O + buf * TF_DK, tmem_base + TF_COL_S + 8 * k,
                       umma_desc_mn_sw128(sV + k * (16 * 128), TF_KBLK_BYTES, 1024), idesc_o, k > 0 ? 1u : 0u);
        umma_commit(&o_full[buf]);
        TF_MARK(0, m, 5);
      };
      const uint32_t items = static_cast<uint32_t>(my_items);
      if (items > 0) issue_g(0);
      if (items > 1) issue_g(1);
      for (uint32_t m = 0; m < items; ++m) {
        issue_s(m);
        if (m + 2 < items) issue_g(m + 2);
        issue_pv(m);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------ WG-A: q, k conversion + softmax ------------------------------------
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c = p.scale_log2e;
    const int my_group = lane >> p.F_log2;  // sequence of this row inside the warp's 32-row block
    const uint32_t zeros[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t m = 0; m < static_cast<uint32_t>(my_items); ++m) {
      const uint32_t buf = m & 1u;
      const bool mark = warp == 4 && lane == 0;
      const uint32_t tg = lane_addr + TF_COL_G + buf * TF_HEAD_ROWS;
      if (mark) TF_MARK(1, m, 0);
      mbar_wait(&g_full[buf], (m >> 1) & 1u);
      if (mark) TF_MARK(1, m, 1);
      tc_fence_after_sync();
      {
        // q: 40 fp32 columns -> 20 packed bf16 columns (+ 4 zero columns: K padded to 48), in place
        uint32_t a[32], b[8];
        tmem_ld_x32(tg, a);
        tmem_ld_x8(tg + 32, b);
        tmem_ld_wait();
        uint32_t lo[16], hi[8];
#pragma unroll
        for (int i = 0; i < 16; ++i) lo[i] = pack_bf16x2(__uint_as_float(a[2 * i]), __uint_as_float(a[2 * i + 1]));
#pragma unroll
        for (int i = 0; i < 4; ++i) hi[i] = pack_bf16x2(__uint_as_float(b[2 * i]), __uint_as_float(b[2 * i + 1]));
#pragma unroll
        for (int i = 4; i < 8; ++i) hi[i] = 0u;
        tmem_st_x16(tg, lo);
        tmem_st_x8(tg + 16, hi);
      }
      {
        // k -> smem operand tile; S(m-1) has completed (this warp consumed it), so the tile is free
        uint32_t a[32], b[8];
        tmem_ld_x32(tg + TF_D, a);
        tmem_ld_x8(tg + TF_D + 32, b);
        tmem_ld_wait();
        tf_store40(a, b, sK, r, false);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&g_free[buf]);  // G(m+2) is issued behind S(m) by the same thread, so q (in place) is safe
        mbar_arrive(&qk_ready);
      }
      if (mark) TF_MARK(1, m, 2);
      mbar_wait(&s_full, m & 1u);
      if (mark) TF_MARK(1, m, 3);
      tc_fence_after_sync();
      uint32_t v[32];
      tmem_ld_x32(lane_addr + TF_COL_S + q * 32, v);
      tmem_ld_wait();
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float s = ((i >> p.F_log2) == my_group) ? __uint_as_float(v[i]) * c : -INFINITY;
        v[i] = __float_as_uint(s);
        mx = fmaxf(mx, s);
      }
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2)
        pk[i >> 1] = pack_bf16x2(tf_exp2(__uint_as_float(v[i]) - mx), tf_exp2(__uint_as_float(v[i + 1]) - mx));
      // P (bf16, 64 packed columns) over the scores: this warp's rows are non-zero only in its own 32-key block
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j == q) tmem_st_x16(lane_addr + TF_COL_S + 16 * j, pk);
        else tmem_st_x16(lane_addr + TF_COL_S + 16 * j, zeros);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready);
      if (mark) TF_MARK(1, m, 4);
    }
  } else if (warp >= 8) {
    // ------------------------------------ WG-B: v conversion + output epilogue ------------------------------------
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int seq = r >> p.F_log2, frame = r & (p.F - 1);
    auto epilogue = [&](uint32_t m) {
      // o_full[m & 1] has been observed by the caller
      const uint32_t buf = m & 1u;
      int b, hw0;
      tile_coords(it_of(m), b, hw0);
      const int head = head_of(m);
      const int hw = hw0 + seq;
      uint32_t o[32], o2[16];
      tmem_ld_x32(lane_addr + TF_COL_O + buf * TF_DK, o);
      tmem_ld_x16(lane_addr + TF_COL_O + buf * TF_DK + 32, o2);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[buf]);
      if (hw < p.HW) {
        const float inv = rcp_fast(__uint_as_float(o2[8]));  // column 40: sum_j P_ij (the ones column of V)
        __nv_bfloat16* orow = p.O + ((static_cast<long long>(b) * p.F + frame) * p.HW + hw) * p.ldo + head * TF_D;
        uint4* dst = reinterpret_cast<uint4*>(orow);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_uint4(pack_bf16x2(__uint_as_float(o[8 * j]) * inv, __uint_as_float(o[8 * j + 1]) * inv),
                              pack_bf16x2(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv),
                              pack_bf16x2(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv),
                              pack_bf16x2(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv));
        dst[4] = make_uint4(pack_bf16x2(__uint_as_float(o2[0]) * inv, __uint_as_float(o2[1]) * inv),
                            pack_bf16x2(__uint_as_float(o2[2]) * inv, __uint_as_float(o2[3]) * inv),
                            pack_bf16x2(__uint_as_float(o2[4]) * inv, __uint_as_float(o2[5]) * inv),
                            pack_bf16x2(__uint_as_float(o2[6]) * inv, __uint_as_float(o2[7]) * inv));
      }
    };
    for (uint32_t m = 0; m < static_cast<uint32_t>(my_items); ++m) {
      const uint32_t buf = m & 1u;
      const bool mark = warp == 8 && lane == 0;
      if (mark) TF_MARK(2, m, 0);
      mbar_wait(&g_full[buf], (m >> 1) & 1u);
      if (mark) TF_MARK(2, m, 1);
      tc_fence_after_sync();
      uint32_t a[32], b2[8];
      tmem_ld_x32(lane_addr + TF_COL_G + buf * TF_HEAD_ROWS + 2 * TF_D, a);
      tmem_ld_x8(lane_addr + TF_COL_G + buf * TF_HEAD_ROWS + 2 * TF_D + 32, b2);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&g_free[buf]);
      // the V tile is free (and O(m-1) is complete) once PV(m-1) has completed
      if (m > 0) {
        mbar_wait(&o_full[(m - 1) & 1u], ((m - 1) >> 1) & 1u);
        tc_fence_after_sync();
      }
      if (mark) TF_MARK(2, m, 2);
      // columns 40..47 of the V tile are padding: column 40 carries 1.0 so that O[:, 40] = sum_j P_ij
      tf_store40(a, b2, sV, r, true);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&v_ready);
      if (mark) TF_MARK(2, m, 3);
      if (m > 0) epilogue(m - 1);
      if (mark) TF_MARK(2, m, 4);
    }
    if (my_items > 0) {
      const uint32_t m = static_cast<uint32_t>(my_items) - 1;
      mbar_wait(&o_full[m & 1u], (m >> 1) & 1u);
      tc_fence_after_sync();
      epilogue(m);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace fmc

using namespace fmc;

// Diagnostics: device buffer of 4 * 64 * 8 int64 receiving clock64() stamps of CTA 0's pipeline events (nullptr = off).
extern "C" int fmc_debug_set_timeline(void* device_buffer) {
  g_tf_timeline = static_cast<long long*>(device_buffer);
  return FMC_OK;
}

extern "C" int fmc_temporal_qkv_attn_bf16(const void* X, long long ldx, const void* Wqkv, long long ldw, void* O,
                                          long long ldo, int B, int F, int HW, int channels, int heads, float scale,
                                          void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(X && Wqkv && O, FMC_ERR_ARG, "fmc_temporal_qkv_attn_bf16: null operand");
  FMC_REQUIRE(channels == TF_C && heads == TF_HEADS, FMC_ERR_SHAPE,
              "fmc_temporal_qkv_attn_bf16: only %d channels x %d heads is implemented (got %d x %d); use "
              "fmc_gemm_bf16 + fmc_temporal_attn_bf16", TF_C, TF_HEADS, channels, heads);
  FMC_REQUIRE(F == 4 || F == 8 || F == 16 || F == 32, FMC_ERR_SHAPE,
              "fmc_temporal_qkv_attn_bf16: frame count %d not in {4, 8, 16, 32}", F);
  FMC_REQUIRE(ldx % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0 && ldx >= TF_C && ldw >= TF_C && ldo >= TF_C, FMC_ERR_SHAPE,
              "fmc_temporal_qkv_attn_bf16: row strides must be multiples of 8 elements and >= %d", TF_C);
  FMC_REQUIRE((reinterpret_cast<uintptr_t>(O) & 15) == 0, FMC_ERR_SHAPE, "fmc_temporal_qkv_attn_bf16: O not 16-byte aligned");
  FMC_REQUIRE(B > 0 && HW > 0, FMC_ERR_SHAPE, "fmc_temporal_qkv_attn_bf16: empty problem");

  CUtensorMap tmX, tmW;
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(TF_C), static_cast<uint64_t>(HW), static_cast<uint64_t>(B) * F};
    const uint64_t strides[2] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(HW) * ldx * 2};
    const uint32_t box[3] = {64, 1, static_cast<uint32_t>(F)};
    int rc = make_tmap_bf16(&tmX, X, 3, dims, strides, box, true);
    if (rc != FMC_OK) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(TF_C), static_cast<uint64_t>(TF_HEADS * TF_HEAD_ROWS)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    const uint32_t box[2] = {64, TF_HEAD_ROWS};
    int rc = make_tmap_bf16(&tmW, Wqkv, 2, dims, strides, box, true);
    if (rc != FMC_OK) return rc;
  }
  TfParams p{};
  p.B = B; p.F = F; p.HW = HW;
  p.F_log2 = (F == 4) ? 2 : (F == 8 ? 3 : (F == 16 ? 4 : 5));
  p.seqs_per_tile = 128 / F;
  p.tiles_per_b = ceil_div(HW, p.seqs_per_tile);
  p.scale_log2e = scale * 1.4426950408889634f;
  p.O = static_cast<__nv_bfloat16*>(O);
  p.ldo = ldo;
  p.timeline = g_tf_timeline;
  static unsigned long long attr_devs = 0;  // per device: the attribute belongs to the (device, function) pair
  if (first_use_on_this_device(&attr_devs)) {
    FMC_CUDA_OK(cudaFuncSetAttribute(temporal_qkv_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM_BYTES));
  }
  const int items = B * p.tiles_per_b * TF_HEADS;
  const int grid = items < device_sm_count() ? items : device_sm_count();
  launch_k(temporal_qkv_attn_kernel, dim3(grid), dim3(TF_THREADS), TF_SMEM_BYTES, stream, tmX, tmW, p);
  return check_launch("temporal_qkv_attn_kernel");
}
