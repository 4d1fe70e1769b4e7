// Spatial self-attention (and, dQ only, text cross-attention) BACKWARD on the tensor cores at head_dim 40 / 80 / 160 --
// levels 0 and 1 dominate the training step: 16 images x 8 heads x 2560 x 2560 resp. 640 x 640 scores per attention.  Flash style, probabilities recomputed; two kernels, both built like
// the forward kernel of attn_spatial.cu (TMA producer warp, one tcgen05.mma issuer thread, TMEM allocator warp, four
// softmax warps with one TMEM lane = one row per thread) and arranged so that EVERY MMA has the forward's operand forms --
// A K-major from shared memory, B K-major for the score-type products, B MN-major for the accumulate-type products:
//
//   dQ kernel   per (image, head, 128 queries), two sweeps over the key tiles:
//                 sweep 1:  S = Q K^T                          -> row log-sum-exp L (log2 units), D = sum_c dO O
//                 sweep 2:  S = Q K^T,  dP = dO V^T             (score-type, TMEM columns 0 / 128)
//                           dS = exp2(S c - L) (dP - D) scale   -> bf16 -> smem (the forward's P layout)
//                           dQ += dS K                           (accumulate-type: B = K tile, MN-major)
//   dK/dV kernel per (image, head, 128 keys), one sweep over the query tiles, on TRANSPOSED scores:
//                 S^T = K Q^T,  dP^T = V dO^T                    (TMEM lane = key, column = query)
//                 P^T = exp2(S^T c - L[query]),  dS^T = P^T (dP^T - D[query]) scale   -> bf16 -> smem
//                 dV += P^T dO,  dK += dS^T Q                    (B = dO / Q tile, MN-major)
//
// The dO tiles come through a 3-D tensor map (d, head, row) whose 64-wide boxes are zero-filled past the head's last
// column, so the contraction over the padded head width (48 for head_dim 40) of dP = dO V^T sees zeros there whatever V
// holds in those columns; Q and K heads are zero-padded by the projection itself.  Operand tiles wider than 64 columns
// (head_dim 80) are two SWIZZLE_128B chunks, AB_TILE bytes apart (K-major: chunk = k / 4; MN-major: LBO).  The dK/dV
// kernel takes 64-query tiles at head_dim 80 so that its resident K / V tiles and a two-deep (Q, dO) ring fit.  L and D travel from the first kernel to the second through the
// [row, head] fp32 scratch buffers of the entry point.
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fmc {

constexpr int AB_BM = 128;
constexpr int AB_THREADS = 256;
constexpr int AB_TILE = AB_BM * 128;        // [128 rows x 64 bf16] SWIZZLE_128B tile
constexpr int AB_PT = 2 * AB_BM * 128;      // [128 x 128] bf16 probability-type tile (two 64-column chunks)

struct AbParams {
  int heads, n, images, blocks;  // n query tokens per image, blocks = ceil(n / 128)
  int nk, key_blocks, kv_div, kv_stride;  // keys per group; image i reads keys of group i / kv_div at row group * kv_stride
  int have_lse;  // lse already holds the forward's row log-sum-exp (log2 units): the ping-pong dQ kernel skips its first sweep
  int head_stride, q_col0, k_col0, v_col0;
  float scale, scale_log2e;
  const __nv_bfloat16* O;
  const __nv_bfloat16* dO;
  long long ldo, lddo;
  __nv_bfloat16* dQ; long long lddq; int dq_col0;
  __nv_bfloat16* dK; long long lddk; int dk_col0;
  __nv_bfloat16* dV; long long lddv; int dv_col0;
  float* lse;   // [rows, heads], log2 units
  float* dsum;  // [rows, heads]
};

__device__ __forceinline__ float ab_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ab_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void ab_tma_load_3d(uint32_t smem_dst, const void* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// row r of a [128 x 128] bf16 tile in the forward's P layout: 16-byte piece j (8 columns) of the row
__device__ __forceinline__ uint32_t ab_piece_addr(uint32_t tile, int r, int piece) {
  return tile + static_cast<uint32_t>(r) * 128u + static_cast<uint32_t>(piece >> 3) * AB_TILE +
         (static_cast<uint32_t>((piece & 7) ^ (r & 7)) << 4);
}

// =====================================================================================================================
// dQ kernel
// =====================================================================================================================
template <int D, int ABQ_STAGES>
struct AbqCfg {
  static constexpr int DK = (D + 15) / 16 * 16;
  static constexpr int CH = (DK + 63) / 64;
  static constexpr int OT = CH * AB_TILE;  // one operand tile: 128 rows x DK columns
  static constexpr int SMEM = 2 * OT + AB_PT + ABQ_STAGES * 2 * OT + 1024;
};

template <int AB_D, int ABQ_STAGES>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmG, AbParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full, q_empty, s_full, s_free, ds_ready, ds_free, dq_final, dq_free;
  __shared__ uint64_t kv_full[ABQ_STAGES], kv_empty[ABQ_STAGES];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  using Cfg = AbqCfg<AB_D, ABQ_STAGES>;
  constexpr int AB_DK = Cfg::DK, CH = Cfg::CH, OT = Cfg::OT;
  const uint32_t sQ = smem_base, sG = sQ + OT, sDS = sG + OT, sKV = sDS + AB_PT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.images * p.heads * p.blocks;
  const int T = p.key_blocks;  // key tiles per sweep (self-attention: = blocks; text cross-attention: 1)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmG);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&q_full, 1); mbar_init(&q_empty, 1);
    mbar_init(&s_full, 1); mbar_init(&s_free, 4);
    mbar_init(&ds_ready, 4); mbar_init(&ds_free, 1);
    mbar_init(&dq_final, 1); mbar_init(&dq_free, 4);
    for (int s = 0; s < ABQ_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      uint32_t t = 0, it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int qb = item % p.blocks;
        const int head = (item / p.blocks) % p.heads;
        const int img = item / (p.blocks * p.heads);
        const int row0 = img * p.n;
        const int kv_row0 = (img / p.kv_div) * p.kv_stride;
        mbar_wait(&q_empty, (it & 1u) ^ 1u);
        mbar_arrive_expect_tx(&q_full, 2 * OT);
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
          tma_load_2d_a(sQ + ch * AB_TILE, &tmQ, &q_full, p.q_col0 + head * p.head_stride + ch * 64, row0 + qb * AB_BM);
          ab_tma_load_3d(sG + ch * AB_TILE, &tmG, &q_full, ch * 64, head, row0 + qb * AB_BM);
        }
        for (int u = 0; u < 2 * T; ++u, ++t) {
          const int j = u < T ? u : u - T;
          const int st = t % ABQ_STAGES;
          mbar_wait(&kv_empty[st], ((t / ABQ_STAGES) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&kv_full[st], 2 * OT);
          const uint32_t sK = sKV + st * 2 * OT;
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
            tma_load_2d_a(sK + ch * AB_TILE, &tmK, &kv_full[st], p.k_col0 + head * p.head_stride + ch * 64, kv_row0 + j * AB_BM);
            tma_load_2d_a(sK + OT + ch * AB_TILE, &tmV, &kv_full[st], p.v_col0 + head * AB_D + ch * 64, kv_row0 + j * AB_BM);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(AB_BM, AB_BM);
      constexpr uint32_t idesc_a = umma_idesc_bf16_bmn(AB_BM, AB_DK);
      uint32_t t = 0, it = 0, nds = 0;
      // dQ += dS(tile u) K(tile u): the K tile sits in stage `st`
      auto issue_dq = [&](int st, bool first) {
        mbar_wait(&ds_ready, nds & 1u);
        tc_fence_after_sync();
        const uint32_t sK = sKV + st * 2 * OT;
#pragma unroll
        for (int k = 0; k < AB_BM / 16; ++k) {
          const uint64_t da = umma_desc_k_sw128(sDS + (k >> 2) * AB_TILE + (k & 3) * 32);
          const uint64_t db = umma_desc_mn_sw128(sK + k * (16 * 128), AB_TILE, 1024);
          umma_bf16_ss(tmem_base + 256u, da, db, idesc_a, (!first || k > 0) ? 1u : 0u);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(&ds_free);
        ++nds;
      };
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        mbar_wait(&q_full, it & 1u);
        tc_fence_after_sync();
        int prev_st = -1;
        for (int u = 0; u < 2 * T; ++u, ++t) {
          const int st = t % ABQ_STAGES;
          if constexpr (ABQ_STAGES == 1) {
            // one K / V stage (head_dim 160): the pending dQ MMA is what frees it, so it goes first (no overlap between
            // the score MMAs of this tile and the softmax of the previous one; these shapes have two key tiles)
            if (u > T) {
              issue_dq(prev_st, u == T + 1);
              prev_st = -2;
            }
          }
          mbar_wait(&kv_full[st], (t / ABQ_STAGES) & 1u);
          mbar_wait(&s_free, (t & 1u) ^ 1u);
          tc_fence_after_sync();
          const uint32_t sK = sKV + st * 2 * OT;
#pragma unroll
          for (int k = 0; k < AB_DK / 16; ++k) {
            const uint32_t off = (k >> 2) * AB_TILE + (k & 3) * 32;
            umma_bf16_ss(tmem_base, umma_desc_k_sw128(sQ + off), umma_desc_k_sw128(sK + off), idesc_s, k > 0 ? 1u : 0u);
          }
          if (u >= T) {
#pragma unroll
            for (int k = 0; k < AB_DK / 16; ++k) {
              const uint32_t off = (k >> 2) * AB_TILE + (k & 3) * 32;
              umma_bf16_ss(tmem_base + 128u, umma_desc_k_sw128(sG + off), umma_desc_k_sw128(sK + OT + off), idesc_s,
                           k > 0 ? 1u : 0u);
            }
          }
          umma_commit(&s_full);
          if (u < T) {
            umma_commit(&kv_empty[st]);
          } else {
            if (prev_st >= 0) {
              issue_dq(prev_st, u == T + 1);
            } else if (prev_st == -1) {
              mbar_wait(&dq_free, (it & 1u) ^ 1u);  // the previous item's epilogue has read the dQ accumulator
              tc_fence_after_sync();
            }
            prev_st = st;
          }
        }
        issue_dq(prev_st, T == 1);
        umma_commit(&q_empty);
        umma_commit(&dq_final);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c = p.scale_log2e;
    uint32_t t = 0, it = 0, nds = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int qb = item % p.blocks;
      const int head = (item / p.blocks) % p.heads;
      const int img = item / (p.blocks * p.heads);
      const int q_in_img = qb * AB_BM + r;
      const bool row_ok = q_in_img < p.n;
      const long long row = static_cast<long long>(img) * p.n + q_in_img;
      // D = sum_c dO[row, c] O[row, c] straight from global memory (one row per thread, 80 bytes each)
      float dsum = 0.f;
      if (row_ok) {
        const uint4* op = reinterpret_cast<const uint4*>(p.O + row * p.ldo + head * AB_D);
        const uint4* gp = reinterpret_cast<const uint4*>(p.dO + row * p.lddo + head * AB_D);
#pragma unroll
        for (int v = 0; v < AB_D / 8; ++v) {
          const uint4 a = __ldg(op + v), b = __ldg(gp + v);
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) dsum += bf16_lo(aw[e]) * bf16_lo(bw[e]) + bf16_hi(aw[e]) * bf16_hi(bw[e]);
        }
      }
      float m = -INFINITY, l = 0.f, L2 = 0.f;
      for (int u = 0; u < 2 * T; ++u, ++t) {
        const int j = u < T ? u : u - T;
        const int valid = p.nk - j * AB_BM;
        mbar_wait(&s_full, t & 1u);
        tc_fence_after_sync();
        if (u < T) {
          // ---- sweep 1: online row maximum / sum in log2 units
          float mx = -INFINITY;
          uint32_t v[32];
#pragma unroll 1
          for (int c4 = 0; c4 < 4; ++c4) {
            tmem_ld_x32(lane_addr + c4 * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float s = (c4 * 32 + i < valid) ? __uint_as_float(v[i]) * c : -INFINITY;
              v[i] = __float_as_uint(s);
              mx = fmaxf(mx, s);
            }
            const float m_new = fmaxf(m, mx);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) sum += ab_exp2(__uint_as_float(v[i]) - m_new);
            l = l * ab_exp2(m - m_new) + sum;
            m = m_new;
          }
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free);
          if (u == T - 1) {
            L2 = m + log2f(l);
            if (row_ok) {
              p.lse[row * p.heads + head] = L2;
              p.dsum[row * p.heads + head] = dsum;
            }
          }
        } else {
          // ---- sweep 2: dS = exp2(S c - L) (dP - D) scale  -> bf16 -> smem, 32 columns at a time
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint32_t sv[32], dv[32], pk[16];
            tmem_ld_x32(lane_addr + c4 * 32, sv);
            tmem_ld_x32(lane_addr + 128u + c4 * 32, dv);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const bool ok0 = c4 * 32 + i < valid, ok1 = c4 * 32 + i + 1 < valid;
              const float p0 = ok0 ? ab_exp2(fmaf(__uint_as_float(sv[i]), c, -L2)) : 0.f;
              const float p1 = ok1 ? ab_exp2(fmaf(__uint_as_float(sv[i + 1]), c, -L2)) : 0.f;
              const float d0 = ok0 ? p0 * (__uint_as_float(dv[i]) - dsum) * p.scale : 0.f;
              const float d1 = ok1 ? p1 * (__uint_as_float(dv[i + 1]) - dsum) * p.scale : 0.f;
              pk[i >> 1] = pack_bf16x2(d0, d1);
            }
            if (c4 == 0) mbar_wait(&ds_free, (nds & 1u) ^ 1u);  // dQ MMA of the previous tile has read the dS tile
#pragma unroll
            for (int g = 0; g < 4; ++g)
              st_shared_v4(ab_piece_addr(sDS, r, c4 * 4 + g), pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          }
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free);  // S / dP consumed: the next tile's score MMAs may overwrite them
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ds_ready);
          ++nds;
        }
      }
      // ---- epilogue: dQ accumulator -> bf16 -> global (40 of the 48 columns)
      mbar_wait(&dq_final, it & 1u);
      tc_fence_after_sync();
      __nv_bfloat16* dst_row = p.dQ + row * p.lddq + p.dq_col0 + head * p.head_stride;
#pragma unroll 1
      for (int cc = 0; cc < AB_DK / 16; ++cc) {
        uint32_t o[16];
        tmem_ld_x16(lane_addr + 256u + cc * 16, o);
        tmem_ld_wait();
        if (row_ok) {
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(__uint_as_float(o[2 * i]), __uint_as_float(o[2 * i + 1]));
          uint4* dst = reinterpret_cast<uint4*>(dst_row + cc * 16);
          if (cc * 16 + 8 <= AB_D) dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
          if (cc * 16 + 16 <= AB_D) dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dq_free);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================================
// dQ kernel, ping-pong form: 64-key tiles, TWO softmax warpgroups on alternate tiles (each thread still owns one query row;
// group g = tile parity).  The single-group kernel above has one warp per scheduler doing the exponentials and cannot
// overlap a tile's MMAs with its softmax (one S / dP buffer): ncu showed tensor pipe 12 %, XU 34 %, issue 42 % at level 0.
// Here S / dP are double-buffered in TMEM (2 x (64 + 64) columns + dQ), each group has its own dS tile in shared memory,
// and the MMA thread alternates between the groups.  The row log-sum-exp of the first sweep is folded across the two
// groups through shared memory.
// =====================================================================================================================
constexpr int ABP_BN = 64;
constexpr int ABP_THREADS = 384;
template <int D, int STAGES>
struct AbpCfg {
  static constexpr int DK = (D + 15) / 16 * 16;
  static constexpr int CH = (DK + 63) / 64;
  static constexpr int OT = CH * AB_TILE;            // Q / dO tile: 128 rows x DK columns
  static constexpr int KC = ABP_BN * 128;            // one 64-column chunk of a 64-key tile
  static constexpr int KT = CH * KC;                 // K / V tile: 64 keys x DK columns
  static constexpr int DS = AB_BM * 128;             // dS tile of one group: 128 queries x 64 keys
  static constexpr int SMEM = 2 * OT + 2 * DS + STAGES * 2 * KT + 1024;
  static_assert(256 + DK <= 512, "TMEM: 2 x (S, dP) + dQ");
};

template <int AB_D, int STAGES>
__global__ void __launch_bounds__(ABP_THREADS, 1)
attn_bwd_dq_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmG, AbParams p) {
  pdl_launch_dependents();
  using Cfg = AbpCfg<AB_D, STAGES>;
  constexpr int AB_DK = Cfg::DK, CH = Cfg::CH, OT = Cfg::OT, KC = Cfg::KC, KT = Cfg::KT, DS = Cfg::DS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full, q_empty, dq_final, dq_free;
  __shared__ uint64_t s_full[2], s_free[2], ds_ready[2], ds_free[2];
  __shared__ uint64_t kv_full[STAGES], kv_empty[STAGES];
  __shared__ float fold_m[2][AB_BM], fold_l[2][AB_BM];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base, sG = sQ + OT, sDS = sG + OT, sKV = sDS + 2 * DS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.images * p.heads * p.blocks;
  const int T = (p.nk + ABP_BN - 1) / ABP_BN;  // 64-key tiles per sweep
  const int U0 = p.have_lse ? T : 0;           // first step of an item: the log-sum-exp sweep is skipped when the forward supplied it

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmG);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&q_full, 1); mbar_init(&q_empty, 1);
    mbar_init(&dq_final, 1); mbar_init(&dq_free, 4);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1); mbar_init(&s_free[g], 4);
      mbar_init(&ds_ready[g], 4); mbar_init(&ds_free[g], 1);
    }
    for (int s = 0; s < STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      uint32_t t = 0, it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int qb = item % p.blocks;
        const int head = (item / p.blocks) % p.heads;
        const int img = item / (p.blocks * p.heads);
        const int row0 = img * p.n;
        const int kv_row0 = (img / p.kv_div) * p.kv_stride;
        mbar_wait(&q_empty, (it & 1u) ^ 1u);
        mbar_arrive_expect_tx(&q_full, 2 * OT);
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
          tma_load_2d_a(sQ + ch * AB_TILE, &tmQ, &q_full, p.q_col0 + head * p.head_stride + ch * 64, row0 + qb * AB_BM);
          ab_tma_load_3d(sG + ch * AB_TILE, &tmG, &q_full, ch * 64, head, row0 + qb * AB_BM);
        }
        for (int u = U0; u < 2 * T; ++u, ++t) {
          const int j = u < T ? u : u - T;
          const int st = t % STAGES;
          mbar_wait(&kv_empty[st], ((t / STAGES) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&kv_full[st], 2 * KT);
          const uint32_t sK = sKV + st * 2 * KT;
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
            tma_load_2d_a(sK + ch * KC, &tmK, &kv_full[st], p.k_col0 + head * p.head_stride + ch * 64, kv_row0 + j * ABP_BN);
            tma_load_2d_a(sK + KT + ch * KC, &tmV, &kv_full[st], p.v_col0 + head * AB_D + ch * 64, kv_row0 + j * ABP_BN);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(AB_BM, ABP_BN);
      constexpr uint32_t idesc_a = umma_idesc_bf16_bmn(AB_BM, AB_DK);
      uint32_t t = 0, it = 0;
      uint32_t n_tile[2] = {0, 0};  // tiles handed to each group so far (barrier phases)
      uint32_t n_ds[2] = {0, 0};    // dS tiles consumed from each group so far
      // dQ += dS_g K(stage st)
      auto issue_dq = [&](int g, int st, bool first) {
        mbar_wait(&ds_ready[g], n_ds[g] & 1u);
        tc_fence_after_sync();
        const uint32_t sK = sKV + st * 2 * KT;
#pragma unroll
        for (int k = 0; k < ABP_BN / 16; ++k) {
          const uint64_t da = umma_desc_k_sw128(sDS + g * DS + k * 32);
          const uint64_t db = umma_desc_mn_sw128(sK + k * (16 * 128), KC, 1024);
          umma_bf16_ss(tmem_base + 256u, da, db, idesc_a, (!first || k > 0) ? 1u : 0u);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(&ds_free[g]);
        ++n_ds[g];
      };
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        mbar_wait(&q_full, it & 1u);
        tc_fence_after_sync();
        int prev_st = -1, prev_g = 0;
        for (int u = U0; u < 2 * T; ++u, ++t) {
          const int st = t % STAGES;
          const int g = u & 1;
          mbar_wait(&kv_full[st], (t / STAGES) & 1u);
          mbar_wait(&s_free[g], (n_tile[g] & 1u) ^ 1u);
          tc_fence_after_sync();
          const uint32_t sK = sKV + st * 2 * KT;
          const uint32_t col = tmem_base + static_cast<uint32_t>(g) * 128u;
#pragma unroll
          for (int k = 0; k < AB_DK / 16; ++k)
            umma_bf16_ss(col, umma_desc_k_sw128(sQ + (k >> 2) * AB_TILE + (k & 3) * 32),
                         umma_desc_k_sw128(sK + (k >> 2) * KC + (k & 3) * 32), idesc_s, k > 0 ? 1u : 0u);
          if (u >= T) {
#pragma unroll
            for (int k = 0; k < AB_DK / 16; ++k)
              umma_bf16_ss(col + 64u, umma_desc_k_sw128(sG + (k >> 2) * AB_TILE + (k & 3) * 32),
                           umma_desc_k_sw128(sK + KT + (k >> 2) * KC + (k & 3) * 32), idesc_s, k > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[g]);
          ++n_tile[g];
          if (u < T) {
            umma_commit(&kv_empty[st]);
          } else {
            if (prev_st >= 0) {
              issue_dq(prev_g, prev_st, u == T + 1);
            } else {
              mbar_wait(&dq_free, (it & 1u) ^ 1u);  // the previous item's epilogue has read the dQ accumulator
              tc_fence_after_sync();
            }
            prev_st = st;
            prev_g = g;
          }
        }
        issue_dq(prev_g, prev_st, T == 1);
        umma_commit(&q_empty);
        umma_commit(&dq_final);
      }
    }
  } else if (warp >= 4) {
    const int g = (warp - 4) >> 2;  // softmax group: tiles with u & 1 == g
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g) * 128u;
    const uint32_t dq_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 256u;
    const uint32_t my_ds = sDS + g * DS;
    const float c = p.scale_log2e;
    uint32_t n_tile = 0, n_ds = 0, it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int qb = item % p.blocks;
      const int head = (item / p.blocks) % p.heads;
      const int img = item / (p.blocks * p.heads);
      const int q_in_img = qb * AB_BM + r;
      const bool row_ok = q_in_img < p.n;
      const long long row = static_cast<long long>(img) * p.n + q_in_img;
      float dsum = 0.f;
      if (row_ok) {
        const uint4* op = reinterpret_cast<const uint4*>(p.O + row * p.ldo + head * AB_D);
        const uint4* gp = reinterpret_cast<const uint4*>(p.dO + row * p.lddo + head * AB_D);
#pragma unroll
        for (int v = 0; v < AB_D / 8; ++v) {
          const uint4 a = __ldg(op + v), b = __ldg(gp + v);
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) dsum += bf16_lo(aw[e]) * bf16_lo(bw[e]) + bf16_hi(aw[e]) * bf16_hi(bw[e]);
        }
      }
      float m = -INFINITY, l = 0.f;
      // ---- sweep 1 (this group's tiles): online row maximum / sum in log2 units
      for (int u = g + U0; u < T; u += 2) {
        const int valid = p.nk - u * ABP_BN;
        mbar_wait(&s_full[g], n_tile & 1u);
        ++n_tile;
        tc_fence_after_sync();
        float mx = -INFINITY;
        uint32_t va[32], vb[32];
        auto sweep1_chunk = [&](uint32_t (&v)[32], int c2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float s = (c2 * 32 + i < valid) ? __uint_as_float(v[i]) * c : -INFINITY;
            v[i] = __float_as_uint(s);
            mx = fmaxf(mx, s);
          }
          const float m_new = fmaxf(m, mx);
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) sum += ab_exp2(__uint_as_float(v[i]) - m_new);
          l = l * ab_exp2(m - m_new) + sum;
          m = m_new;
        };
        tmem_ld_x32(lane_addr, va);
        tmem_ld_wait();
        tmem_ld_x32(lane_addr + 32, vb);  // in flight while the first chunk is reduced
        sweep1_chunk(va, 0);
        tmem_ld_wait();
        sweep1_chunk(vb, 1);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[g]);
      }
      // ---- fold (m, l) of the two groups (a group without a sweep-1 tile contributes (-inf, 0)); group 0 always has tile 0
      float L2;
      if (p.have_lse) {  // uniform over the CTA: no fold barriers
        L2 = row_ok ? p.lse[row * p.heads + head] : 0.f;
        if (g == 0 && row_ok) p.dsum[row * p.heads + head] = dsum;
      } else {
        fold_m[g][r] = m;
        fold_l[g][r] = l;
        ab_bar_sync(1, 256);
        const float m0 = fold_m[0][r], m1 = fold_m[1][r];
        const float mm = fmaxf(m0, m1);
        const float lsum = fold_l[0][r] * ab_exp2(m0 - mm) + fold_l[1][r] * ab_exp2(m1 - mm);
        L2 = mm + log2f(lsum);
        if (g == 0 && row_ok) {
          p.lse[row * p.heads + head] = L2;
          p.dsum[row * p.heads + head] = dsum;
        }
        ab_bar_sync(2, 256);  // fold_* are rewritten by the next item only after both groups have read them
      }
      // ---- sweep 2 (this group's tiles): dS = exp2(S c - L) (dP - D) scale  -> bf16 -> this group's smem tile
      for (int u = T + ((T & 1) == g ? 0 : 1); u < 2 * T; u += 2) {
        const int valid = p.nk - (u - T) * ABP_BN;
        mbar_wait(&s_full[g], n_tile & 1u);
        ++n_tile;
        tc_fence_after_sync();
        uint32_t sa[16], da[16], sb[16], db[16];
        auto sweep2_chunk = [&](const uint32_t (&sv)[16], const uint32_t (&dv)[16], int c4) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const bool ok0 = c4 * 16 + i < valid, ok1 = c4 * 16 + i + 1 < valid;
            const float p0 = ok0 ? ab_exp2(fmaf(__uint_as_float(sv[i]), c, -L2)) : 0.f;
            const float p1 = ok1 ? ab_exp2(fmaf(__uint_as_float(sv[i + 1]), c, -L2)) : 0.f;
            const float d0 = ok0 ? p0 * (__uint_as_float(dv[i]) - dsum) * p.scale : 0.f;
            const float d1 = ok1 ? p1 * (__uint_as_float(dv[i + 1]) - dsum) * p.scale : 0.f;
            pk[i >> 1] = pack_bf16x2(d0, d1);
          }
          if (c4 == 0) mbar_wait(&ds_free[g], (n_ds & 1u) ^ 1u);  // the dQ MMA of this group's previous tile has read dS
          st_shared_v4(ab_piece_addr(my_ds, r, c4 * 2), pk[0], pk[1], pk[2], pk[3]);
          st_shared_v4(ab_piece_addr(my_ds, r, c4 * 2 + 1), pk[4], pk[5], pk[6], pk[7]);
        };
        // 16-column chunks; the loads of chunk k + 1 are in flight while chunk k is computed
        tmem_ld_x16(lane_addr, sa);
        tmem_ld_x16(lane_addr + 64u, da);
        tmem_ld_wait();
        tmem_ld_x16(lane_addr + 16, sb);
        tmem_ld_x16(lane_addr + 64u + 16, db);
        sweep2_chunk(sa, da, 0);
        tmem_ld_wait();
        tmem_ld_x16(lane_addr + 32, sa);
        tmem_ld_x16(lane_addr + 64u + 32, da);
        sweep2_chunk(sb, db, 1);
        tmem_ld_wait();
        tmem_ld_x16(lane_addr + 48, sb);
        tmem_ld_x16(lane_addr + 64u + 48, db);
        sweep2_chunk(sa, da, 2);
        tmem_ld_wait();
        sweep2_chunk(sb, db, 3);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[g]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ds_ready[g]);
        ++n_ds;
      }
      if (g == 0) {
        // ---- epilogue (group 0): dQ accumulator -> bf16 -> global
        mbar_wait(&dq_final, it & 1u);
        tc_fence_after_sync();
        __nv_bfloat16* dst_row = p.dQ + row * p.lddq + p.dq_col0 + head * p.head_stride;
#pragma unroll 1
        for (int cc = 0; cc < AB_DK / 16; ++cc) {
          uint32_t o[16];
          tmem_ld_x16(dq_addr + cc * 16, o);
          tmem_ld_wait();
          if (row_ok) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(__uint_as_float(o[2 * i]), __uint_as_float(o[2 * i + 1]));
            uint4* dst = reinterpret_cast<uint4*>(dst_row + cc * 16);
            if (cc * 16 + 8 <= AB_D) dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            if (cc * 16 + 16 <= AB_D) dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dq_free);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================================
// dK / dV kernel (transposed scores: TMEM lane = key, column = query)
// =====================================================================================================================
template <int D, int BQ, int STAGES>
struct AbkCfg {
  static constexpr int DK = (D + 15) / 16 * 16;
  static constexpr int CH = (DK + 63) / 64;
  static constexpr int OT = CH * AB_TILE;    // resident K / V tile: 128 rows x DK columns
  static constexpr int QC = BQ * 128;        // one 64-column chunk of a BQ-row (Q / dO) tile
  static constexpr int QT = CH * QC;         // Q / dO tile: BQ rows x DK columns
  static constexpr int PT = (BQ / 64) * AB_TILE;  // P^T / dS^T tile: 128 keys x BQ queries
  static constexpr int SMEM = 2 * OT + 2 * PT + STAGES * 2 * QT + 1024;
  static constexpr uint32_t COL_DP = BQ;        // TMEM columns: S^T [0, BQ) | dP^T [BQ, 2 BQ) | dV | dK
  static constexpr uint32_t COL_DV = 2 * BQ;
  static constexpr uint32_t COL_DK = 2 * BQ + DK;
  static_assert(2 * BQ + 2 * DK <= 512, "TMEM: S^T, dP^T, dV, dK");
};

template <int AB_D, int BQ, int ABK_STAGES>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmG, AbParams p) {
  pdl_launch_dependents();
  using Cfg = AbkCfg<AB_D, BQ, ABK_STAGES>;
  constexpr int AB_DK = Cfg::DK, CH = Cfg::CH, OT = Cfg::OT, QC = Cfg::QC, QT = Cfg::QT, PT = Cfg::PT;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t kv_full, kv_empty, s_full, s_free, pds_ready, pds_free, acc_final, acc_free;
  __shared__ uint64_t q_full[ABK_STAGES], q_empty[ABK_STAGES];
  __shared__ float vec_l[2][BQ], vec_d[2][BQ];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = smem_base, sV = sK + OT, sP = sV + OT, sDS = sP + PT, sQG = sDS + PT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.images * p.heads * p.blocks;
  const int T = (p.n + BQ - 1) / BQ;  // query tiles per image

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmG);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&kv_full, 1); mbar_init(&kv_empty, 1);
    mbar_init(&s_full, 1); mbar_init(&s_free, 4);
    mbar_init(&pds_ready, 4); mbar_init(&pds_free, 1);
    mbar_init(&acc_final, 1); mbar_init(&acc_free, 4);
    for (int s = 0; s < ABK_STAGES; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      uint32_t t = 0, it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int kb = item % p.blocks;
        const int head = (item / p.blocks) % p.heads;
        const int img = item / (p.blocks * p.heads);
        const int row0 = img * p.n;
        mbar_wait(&kv_empty, (it & 1u) ^ 1u);
        mbar_arrive_expect_tx(&kv_full, 2 * OT);
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
          tma_load_2d_a(sK + ch * AB_TILE, &tmK, &kv_full, p.k_col0 + head * p.head_stride + ch * 64, row0 + kb * AB_BM);
          tma_load_2d_a(sV + ch * AB_TILE, &tmV, &kv_full, p.v_col0 + head * AB_D + ch * 64, row0 + kb * AB_BM);
        }
        for (int i = 0; i < T; ++i, ++t) {
          const int st = t % ABK_STAGES;
          mbar_wait(&q_empty[st], ((t / ABK_STAGES) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&q_full[st], 2 * QT);
          const uint32_t sQ = sQG + st * 2 * QT;
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
            tma_load_2d_a(sQ + ch * QC, &tmQ, &q_full[st], p.q_col0 + head * p.head_stride + ch * 64, row0 + i * BQ);
            ab_tma_load_3d(sQ + QT + ch * QC, &tmG, &q_full[st], ch * 64, head, row0 + i * BQ);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(AB_BM, BQ);
      constexpr uint32_t idesc_a = umma_idesc_bf16_bmn(AB_BM, AB_DK);
      uint32_t t = 0, it = 0, npd = 0;
      // dV += P^T dO, dK += dS^T Q with the (Q, dO) tiles of stage `st`
      auto issue_acc = [&](int st, bool first) {
        mbar_wait(&pds_ready, npd & 1u);
        tc_fence_after_sync();
        const uint32_t sQ = sQG + st * 2 * QT, sG = sQ + QT;
#pragma unroll
        for (int k = 0; k < BQ / 16; ++k) {
          const uint64_t da = umma_desc_k_sw128(sP + (k >> 2) * AB_TILE + (k & 3) * 32);
          const uint64_t db = umma_desc_mn_sw128(sG + k * (16 * 128), QC, 1024);
          umma_bf16_ss(tmem_base + Cfg::COL_DV, da, db, idesc_a, (!first || k > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < BQ / 16; ++k) {
          const uint64_t da = umma_desc_k_sw128(sDS + (k >> 2) * AB_TILE + (k & 3) * 32);
          const uint64_t db = umma_desc_mn_sw128(sQ + k * (16 * 128), QC, 1024);
          umma_bf16_ss(tmem_base + Cfg::COL_DK, da, db, idesc_a, (!first || k > 0) ? 1u : 0u);
        }
        umma_commit(&q_empty[st]);
        umma_commit(&pds_free);
        ++npd;
      };
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        mbar_wait(&kv_full, it & 1u);
        tc_fence_after_sync();
        int prev_st = -1;
        for (int i = 0; i < T; ++i, ++t) {
          const int st = t % ABK_STAGES;
          mbar_wait(&q_full[st], (t / ABK_STAGES) & 1u);
          mbar_wait(&s_free, (t & 1u) ^ 1u);
          tc_fence_after_sync();
          const uint32_t sQ = sQG + st * 2 * QT, sG = sQ + QT;
#pragma unroll
          for (int k = 0; k < AB_DK / 16; ++k)  // S^T = K Q^T
            umma_bf16_ss(tmem_base, umma_desc_k_sw128(sK + (k >> 2) * AB_TILE + (k & 3) * 32),
                         umma_desc_k_sw128(sQ + (k >> 2) * QC + (k & 3) * 32), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < AB_DK / 16; ++k)  // dP^T = V dO^T (dO zero beyond the head's last column)
            umma_bf16_ss(tmem_base + Cfg::COL_DP, umma_desc_k_sw128(sV + (k >> 2) * AB_TILE + (k & 3) * 32),
                         umma_desc_k_sw128(sG + (k >> 2) * QC + (k & 3) * 32), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&s_full);
          if (prev_st >= 0) {
            issue_acc(prev_st, i == 1);
          } else {
            mbar_wait(&acc_free, (it & 1u) ^ 1u);
            tc_fence_after_sync();
          }
          prev_st = st;
        }
        issue_acc(prev_st, T == 1);
        umma_commit(&kv_empty);
        umma_commit(&acc_final);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int r = q * 32 + lane;  // key row of the tile
    const int tid = threadIdx.x - 128;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c = p.scale_log2e;
    uint32_t t = 0, it = 0, npd = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int kb = item % p.blocks;
      const int head = (item / p.blocks) % p.heads;
      const int img = item / (p.blocks * p.heads);
      const long long img_row0 = static_cast<long long>(img) * p.n;
      for (int i = 0; i < T; ++i, ++t) {
        // per-query log-sum-exp / D of this query tile (written by the dQ kernel): +inf masks queries past the image
        const int buf = t & 1u;
        if (tid < BQ) {
          const int qi = i * BQ + tid;
          const bool ok = qi < p.n;
          vec_l[buf][tid] = ok ? p.lse[(img_row0 + qi) * p.heads + head] : INFINITY;
          vec_d[buf][tid] = ok ? p.dsum[(img_row0 + qi) * p.heads + head] : 0.f;
        }
        ab_bar_sync(1, 128);
        mbar_wait(&s_full, t & 1u);
        tc_fence_after_sync();
        uint32_t sa[16], da[16], sb[16], db[16];
        auto chunk16 = [&](const uint32_t (&sv)[16], const uint32_t (&dv)[16], int k) {
          uint32_t pp[8], pd[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const int col = k * 16 + e;
            const float p0 = ab_exp2(fmaf(__uint_as_float(sv[e]), c, -vec_l[buf][col]));
            const float p1 = ab_exp2(fmaf(__uint_as_float(sv[e + 1]), c, -vec_l[buf][col + 1]));
            const float d0 = p0 * (__uint_as_float(dv[e]) - vec_d[buf][col]) * p.scale;
            const float d1 = p1 * (__uint_as_float(dv[e + 1]) - vec_d[buf][col + 1]) * p.scale;
            pp[e >> 1] = pack_bf16x2(p0, p1);
            pd[e >> 1] = pack_bf16x2(d0, d1);
          }
          if (k == 0) mbar_wait(&pds_free, (npd & 1u) ^ 1u);  // the accumulate MMAs of the previous tile have read P^T / dS^T
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            st_shared_v4(ab_piece_addr(sP, r, k * 2 + h), pp[4 * h], pp[4 * h + 1], pp[4 * h + 2], pp[4 * h + 3]);
            st_shared_v4(ab_piece_addr(sDS, r, k * 2 + h), pd[4 * h], pd[4 * h + 1], pd[4 * h + 2], pd[4 * h + 3]);
          }
        };
        // 16-column chunks; the TMEM loads of chunk k + 1 are in flight while chunk k is computed
        tmem_ld_x16(lane_addr, sa);
        tmem_ld_x16(lane_addr + Cfg::COL_DP, da);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < BQ / 16; k += 2) {
          tmem_ld_x16(lane_addr + (k + 1) * 16, sb);
          tmem_ld_x16(lane_addr + Cfg::COL_DP + (k + 1) * 16, db);
          chunk16(sa, da, k);
          tmem_ld_wait();
          if (k + 2 < BQ / 16) {
            tmem_ld_x16(lane_addr + (k + 2) * 16, sa);
            tmem_ld_x16(lane_addr + Cfg::COL_DP + (k + 2) * 16, da);
          }
          chunk16(sb, db, k + 1);
          if (k + 2 < BQ / 16) tmem_ld_wait();
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pds_ready);
        ++npd;
      }
      // ---- epilogue: dV, dK accumulators -> bf16 -> global
      mbar_wait(&acc_final, it & 1u);
      tc_fence_after_sync();
      const int k_in_img = kb * AB_BM + r;
      const bool row_ok = k_in_img < p.n;
      const long long row = img_row0 + k_in_img;
      __nv_bfloat16* dv_row = p.dV + row * p.lddv + p.dv_col0 + head * AB_D;
      __nv_bfloat16* dk_row = p.dK + row * p.lddk + p.dk_col0 + head * p.head_stride;
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {
#pragma unroll 1
        for (int cc = 0; cc < AB_DK / 16; ++cc) {
          uint32_t o[16];
          tmem_ld_x16(lane_addr + (which == 0 ? Cfg::COL_DV : Cfg::COL_DK) + cc * 16, o);
          tmem_ld_wait();
          if (row_ok) {
            uint32_t w[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) w[e] = pack_bf16x2(__uint_as_float(o[2 * e]), __uint_as_float(o[2 * e + 1]));
            uint4* dst = reinterpret_cast<uint4*>((which == 0 ? dv_row : dk_row) + cc * 16);
            if (cc * 16 + 8 <= AB_D) dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            if (cc * 16 + 16 <= AB_D) dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_free);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// host side: tensor maps + launches
// QSTAGES: K / V ring depth of the dQ kernel; PP_STAGES > 0 selects the ping-pong dQ kernel (64-key tiles, two softmax
// groups) with that ring depth -- FMC_ATTN_BWD_DQ1=1 in the environment keeps the single-group kernel for A/B runs
template <int D, int QSTAGES, int BQ, int KSTAGES, int PP_STAGES>
static int attention_bwd_tc_launch(const void* Q, long long ldq, int q_col0, const void* K, long long ldk, int k_col0,
                                   const void* V, long long ldv, int v_col0, int head_stride, const void* O, long long ldo,
                                   const void* dO, long long lddo, void* dQ, long long lddq, int dq_col0, void* dK,
                                   long long lddk, int dk_col0, void* dV, long long lddv, int dv_col0, float* lse, float* dsum,
                                   int images, int heads, int n, int nk, int kv_div, int kv_stride, float scale,
                                   bool have_lse, cudaStream_t stream) {
  using CfgQ = AbqCfg<D, QSTAGES>;
  using CfgK = AbkCfg<D, BQ, KSTAGES>;
  const long long rows = static_cast<long long>(images) * n;
  const bool want_dkv = dK != nullptr;
  const long long kv_rows = want_dkv ? rows : static_cast<long long>((images + kv_div - 1) / kv_div) * kv_stride;
  CUtensorMap tmQ, tmK, tmV, tmG, tmQb, tmGb;
  bool ping_pong = PP_STAGES > 0;
  if (const char* e = getenv("FMC_ATTN_BWD_DQ1")) ping_pong = ping_pong && e[0] != '1';
  auto map2d = [&](CUtensorMap* m, const void* base, long long ld, int box_rows, long long nrows) {
    const uint64_t dims[2] = {static_cast<uint64_t>(ld), static_cast<uint64_t>(nrows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ld) * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(box_rows)};
    return make_tmap_bf16(m, base, 2, dims, strides, box, true);
  };
  // dO as (d, head, row): 64-wide boxes are zero-filled past the head's last column
  auto map_do = [&](CUtensorMap* m, int box_rows) {
    const uint64_t dims[3] = {static_cast<uint64_t>(D), static_cast<uint64_t>(heads), static_cast<uint64_t>(rows)};
    const uint64_t strides[2] = {static_cast<uint64_t>(D) * 2, static_cast<uint64_t>(lddo) * 2};
    const uint32_t box3[3] = {64, 1, static_cast<uint32_t>(box_rows)};
    return make_tmap_bf16(m, dO, 3, dims, strides, box3, true);
  };
  int rc = map2d(&tmQ, Q, ldq, AB_BM, rows);
  if (rc == FMC_OK) rc = map2d(&tmK, K, ldk, AB_BM, kv_rows);
  if (rc == FMC_OK) rc = map2d(&tmV, V, ldv, AB_BM, kv_rows);
  CUtensorMap tmK64, tmV64;  // 64-key boxes of the ping-pong dQ kernel (the dK / dV kernel keeps 128-key tiles)
  if (rc == FMC_OK && ping_pong) rc = map2d(&tmK64, K, ldk, ABP_BN, kv_rows);
  if (rc == FMC_OK && ping_pong) rc = map2d(&tmV64, V, ldv, ABP_BN, kv_rows);
  if (rc == FMC_OK) rc = map_do(&tmG, AB_BM);
  if (rc == FMC_OK && want_dkv) rc = map2d(&tmQb, Q, ldq, BQ, rows);
  if (rc == FMC_OK && want_dkv) rc = map_do(&tmGb, BQ);
  if (rc != FMC_OK) return rc;
  AbParams p{};
  p.heads = heads; p.n = n; p.images = images; p.blocks = ceil_div(n, AB_BM);
  p.nk = nk; p.key_blocks = ceil_div(nk, AB_BM); p.kv_div = kv_div; p.kv_stride = kv_stride;
  p.have_lse = have_lse && ping_pong ? 1 : 0;  // only the ping-pong dQ kernel has the short form; the others recompute
  p.head_stride = head_stride; p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.O = static_cast<const __nv_bfloat16*>(O); p.dO = static_cast<const __nv_bfloat16*>(dO); p.ldo = ldo; p.lddo = lddo;
  p.dQ = static_cast<__nv_bfloat16*>(dQ); p.lddq = lddq; p.dq_col0 = dq_col0;
  p.dK = static_cast<__nv_bfloat16*>(dK); p.lddk = lddk; p.dk_col0 = dk_col0;
  p.dV = static_cast<__nv_bfloat16*>(dV); p.lddv = lddv; p.dv_col0 = dv_col0;
  p.lse = lse; p.dsum = dsum;
  static unsigned long long attr_devs = 0;
  if (first_use_on_this_device(&attr_devs)) {
    FMC_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<D, QSTAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgQ::SMEM));
    FMC_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<D, BQ, KSTAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     CfgK::SMEM));
  }
  const int items = images * heads * p.blocks;
  const int grid = items < device_sm_count() ? items : device_sm_count();
  if constexpr (PP_STAGES > 0) {
    if (ping_pong) {
      using CfgP = AbpCfg<D, PP_STAGES>;
      static unsigned long long pp_devs = 0;
      if (first_use_on_this_device(&pp_devs))
        FMC_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_pp_kernel<D, PP_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         CfgP::SMEM));
      launch_k(attn_bwd_dq_pp_kernel<D, PP_STAGES>, dim3(grid), dim3(ABP_THREADS), CfgP::SMEM, stream, tmQ, tmK64, tmV64, tmG, p);
    }
  }
  if (!ping_pong)
    launch_k(attn_bwd_dq_tc_kernel<D, QSTAGES>, dim3(grid), dim3(AB_THREADS), CfgQ::SMEM, stream, tmQ, tmK, tmV, tmG, p);
  rc = check_launch("attn_bwd_dq kernel");
  if (rc != FMC_OK || !want_dkv) return rc;
  launch_k(attn_bwd_dkv_tc_kernel<D, BQ, KSTAGES>, dim3(grid), dim3(AB_THREADS), CfgK::SMEM, stream, tmQb, tmK, tmV, tmGb, p);
  return check_launch("attn_bwd_dkv_tc_kernel");
}

// head_dim 40 (heads padded to 48), 80 or 160; anything else is the caller's SIMT path.  dK = dV = NULL: dQ only, keys / values of
// group image / kv_div at row group * kv_stride (text cross-attention: nk = 77 inside 80-row groups); else self-attention
int attention_bwd_tc(int head_dim, const void* Q, long long ldq, int q_col0, const void* K, long long ldk, int k_col0,
                     const void* V, long long ldv, int v_col0, int head_stride, const void* O, long long ldo, const void* dO,
                     long long lddo, void* dQ, long long lddq, int dq_col0, void* dK, long long lddk, int dk_col0, void* dV,
                     long long lddv, int dv_col0, float* lse, float* dsum, int images, int heads, int n, int nk, int kv_div,
                     int kv_stride, float scale, bool have_lse, cudaStream_t stream) {
  if (head_dim == 40)
    return attention_bwd_tc_launch<40, 3, 128, 2, 4>(Q, ldq, q_col0, K, ldk, k_col0, V, ldv, v_col0, head_stride, O, ldo, dO, lddo,
                                                  dQ, lddq, dq_col0, dK, lddk, dk_col0, dV, lddv, dv_col0, lse, dsum, images,
                                                  heads, n, nk, kv_div, kv_stride, scale, have_lse, stream);
  if (head_dim == 80)
    return attention_bwd_tc_launch<80, 2, 64, 2, 3>(Q, ldq, q_col0, K, ldk, k_col0, V, ldv, v_col0, head_stride, O, ldo, dO, lddo,
                                                 dQ, lddq, dq_col0, dK, lddk, dk_col0, dV, lddv, dv_col0, lse, dsum, images,
                                                 heads, n, nk, kv_div, kv_stride, scale, have_lse, stream);
  if (head_dim == 160)
    return attention_bwd_tc_launch<160, 1, 64, 2, 0>(Q, ldq, q_col0, K, ldk, k_col0, V, ldv, v_col0, head_stride, O, ldo, dO, lddo,
                                                  dQ, lddq, dq_col0, dK, lddk, dk_col0, dV, lddv, dv_col0, lse, dsum, images,
                                                  heads, n, nk, kv_div, kv_stride, scale, have_lse, stream);
  set_error("attention_bwd_tc: head_dim %d not in {40, 80, 160}", head_dim);
  return FMC_ERR_SHAPE;
}

}  // namespace fmc
