// Spatial attention core (self over N_l latent tokens, cross over 77 text tokens):
//     O[i, :, h] = softmax(Q[i, :, h] K[kv(i), :, h]^T * scale) V[kv(i), :, h]
// Replaces the materialised baddbmm -> softmax -> bmm of the reference
// (fmc/models/attention_processor.py:148-154, diffusers Attention.get_attention_scores): the score matrix
// ([256, 2560, 2560] fp32 = 6.7 GB at level 0) never leaves the SM.
//
// sm_100a design, one CTA per SM, persistent over (image, head, 128-query block) work items:
//   warp 0     TMA producer: Q tile once per item, K / V^T tiles through a ring of smem stages
//   warp 1     MMA issuer  : S = Q K^T (tcgen05, fp32 in TMEM, two S buffers), O += P V (P staged as bf16 in smem)
//   warp 2     TMEM allocator
//   warps 4-7  softmax     : one query row per thread (TMEM lane == row, no shuffles), exp2 with the scale folded in,
//                            lazy rescaling of the TMEM-resident O accumulator (only when the row max grows by > 2^8),
//                            epilogue O / l -> bf16 -> global
// Q, K and V are read exactly as the fused projection GEMM wrote them ([token, q | k | v] rows): Q and K are K-major
// SWIZZLE_128B operands of S = Q K^T, V is the MN-major B operand of O += P V -- no transposes anywhere.
#include <cuda_bf16.h>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace fmc {

constexpr int FA_BM = 128;  // query rows per item
constexpr int FA_BN = 128;  // keys per KV tile
constexpr int FA_THREADS = 256;
constexpr float FA_RESCALE_THRESHOLD = 8.0f;  // log2 units

struct FaParams {
  int heads;
  int nq;             // query tokens per image
  int nk;             // valid keys per kv group
  int kv_div;         // kv group of image i = i / kv_div   (1 for self-attention, f for text cross-attention)
  int kv_stride;      // rows between kv groups in K (and columns in V^T)
  int head_stride;    // elements between heads inside a Q / K row (48 for d = 40: zero padded by the projection)
  int q_col0;         // column of head 0 inside the Q rows
  int k_col0;         // column of head 0 inside the K rows
  int v_col0;         // column of head 0 inside the V rows (V heads are never padded)
  int images;
  int q_blocks;       // ceil(nq / 128)
  int kv_tiles;       // ceil(nk / 128)
  float scale_log2e;  // softmax scale * log2(e)
  __nv_bfloat16* O;
  long long ldo;
  float* lse;         // optional [q_rows, heads]: row log-sum-exp in log2 units (m + log2 l), for the training backward
};

template <int D>
struct FaCfg {
  static constexpr int DK = (D + 15) / 16 * 16;      // 48 / 80 / 160 : K extent of QK^T, N extent of PV
  static constexpr int QCH = (DK + 63) / 64;         // 64-element (128 B) chunks per Q / K row
  static constexpr int Q_BYTES = QCH * FA_BM * 128;
  static constexpr int K_BYTES = QCH * FA_BN * 128;
  // V tile: QCH 64-channel chunks of 128 key rows exactly as the projection wrote them (MN-major B operand of PV)
  static constexpr int V_BYTES = QCH * FA_BN * 128;
  static constexpr int P_BYTES = 2 * FA_BM * 128;
  static constexpr int KV_BYTES = K_BYTES + V_BYTES;
  static constexpr int STAGES = (D <= 48) ? 4 : (D <= 80 ? 2 : 1);
  static constexpr int SMEM_BYTES = Q_BYTES + P_BYTES + STAGES * KV_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;
  static constexpr int O_COL = 256;                  // S buffers at columns 0 and 128
  static_assert(V_BYTES % 1024 == 0 && K_BYTES % 1024 == 0, "tiles must stay 1024-byte aligned");
  static_assert(DK % 16 == 0 && DK <= 256, "UMMA N constraint");
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
__global__ void __launch_bounds__(FA_THREADS, 1)
spatial_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, FaParams p) {
  pdl_launch_dependents();
  using Cfg = FaCfg<D>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int DK = Cfg::DK;

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full, q_empty, p_ready, p_free, o_final, o_free;
  __shared__ uint64_t kv_full[STAGES], kv_empty[STAGES];
  __shared__ uint64_t s_full[2], s_free[2];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sP = sQ + Cfg::Q_BYTES;
  const uint32_t sKV = sP + Cfg::P_BYTES;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_items = p.images * p.heads * p.q_blocks;
  const int T = p.kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&q_full, 1);
    mbar_init(&q_empty, 1);
    mbar_init(&p_ready, 4);
    mbar_init(&p_free, 1);
    mbar_init(&o_final, 1);
    mbar_init(&o_free, 4);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_free[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_slot, Cfg::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();  // everything above (barriers, TMEM, descriptor prefetch) overlaps the previous kernel's tail

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    if (elect_one()) {
      uint32_t t = 0;   // global KV tile counter of this CTA
      uint32_t it = 0;  // item counter of this CTA
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int qb = item % p.q_blocks;
        const int head = (item / p.q_blocks) % p.heads;
        const int img = item / (p.q_blocks * p.heads);
        const int q_row0 = img * p.nq + qb * FA_BM;
        const int kv_row0 = (img / p.kv_div) * p.kv_stride;
        mbar_wait(&q_empty, (it & 1u) ^ 1u);
        mbar_arrive_expect_tx(&q_full, Cfg::Q_BYTES);
#pragma unroll
        for (int c = 0; c < Cfg::QCH; ++c)
          tma_load_2d_a(sQ + c * (FA_BM * 128), &tmQ, &q_full, p.q_col0 + head * p.head_stride + c * 64, q_row0);
        for (int j = 0; j < T; ++j, ++t) {
          const int st = t % STAGES;
          const uint32_t ph = (t / STAGES) & 1u;
          mbar_wait(&kv_empty[st], ph ^ 1u);
          mbar_arrive_expect_tx(&kv_full[st], Cfg::KV_BYTES);
          const uint32_t sK = sKV + st * Cfg::KV_BYTES;
          const uint32_t sV = sK + Cfg::K_BYTES;
#pragma unroll
          for (int c = 0; c < Cfg::QCH; ++c)
            tma_load_2d_a(sK + c * (FA_BN * 128), &tmK, &kv_full[st], p.k_col0 + head * p.head_stride + c * 64,
                          kv_row0 + j * FA_BN);
#pragma unroll
          for (int c = 0; c < Cfg::QCH; ++c)
            tma_load_2d_a(sV + c * (FA_BN * 128), &tmV, &kv_full[st], p.v_col0 + head * D + c * 64,
                          kv_row0 + j * FA_BN);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ------------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(FA_BM, FA_BN);
      constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(FA_BM, DK);
      uint32_t t = 0, it = 0;
      auto issue_pv = [&](uint32_t u, bool first_of_item) {
        mbar_wait(&p_ready, u & 1u);
        tc_fence_after_sync();
        const int st = u % STAGES;
        const uint32_t sV = sKV + st * Cfg::KV_BYTES + Cfg::K_BYTES;
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k) {
          const uint64_t da = umma_desc_k_sw128(sP + (k >> 2) * (FA_BM * 128) + (k & 3) * 32);
          const uint64_t db = umma_desc_mn_sw128(sV + k * (16 * 128), FA_BN * 128, 1024);
          umma_bf16_ss(tmem_base + Cfg::O_COL, da, db, idesc_o, (!first_of_item || k > 0) ? 1u : 0u);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(&p_free);
      };
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        mbar_wait(&q_full, it & 1u);
        mbar_wait(&o_free, (it & 1u) ^ 1u);
        tc_fence_after_sync();
        for (int j = 0; j < T; ++j, ++t) {
          const int st = t % STAGES;
          const uint32_t ph = (t / STAGES) & 1u;
          const uint32_t sb = t & 1u;
          const uint32_t sph = (t >> 1) & 1u;
          // with a single KV stage the previous PV must be issued first: its commit is what frees the stage
          if constexpr (STAGES == 1) {
            if (j >= 1) issue_pv(t - 1, j == 1);
          }
          mbar_wait(&kv_full[st], ph);
          mbar_wait(&s_free[sb], sph ^ 1u);
          tc_fence_after_sync();
          const uint32_t sK = sKV + st * Cfg::KV_BYTES;
#pragma unroll
          for (int k = 0; k < DK / 16; ++k) {
            const uint64_t da = umma_desc_k_sw128(sQ + (k >> 2) * (FA_BM * 128) + (k & 3) * 32);
            const uint64_t db = umma_desc_k_sw128(sK + (k >> 2) * (FA_BN * 128) + (k & 3) * 32);
            umma_bf16_ss(tmem_base + sb * 128u, da, db, idesc_s, k > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[sb]);
          if (j == T - 1) umma_commit(&q_empty);
          if constexpr (STAGES > 1) {
            if (j >= 1) issue_pv(t - 1, j == 1);
          }
        }
        issue_pv(t - 1, T == 1);
        umma_commit(&o_final);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------ softmax + epilogue ------------------------------------
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row inside the 128-query tile == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c = p.scale_log2e;
    uint32_t t = 0, it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int qb = item % p.q_blocks;
      const int head = (item / p.q_blocks) % p.heads;
      const int img = item / (p.q_blocks * p.heads);
      float m_used = -INFINITY;
      float l = 0.f;
      for (int j = 0; j < T; ++j, ++t) {
        const uint32_t sb = t & 1u;
        const uint32_t sph = (t >> 1) & 1u;
        const int valid = p.nk - j * FA_BN;  // keys of this tile with index < valid are real
        mbar_wait(&s_full[sb], sph);
        tc_fence_after_sync();
        const uint32_t s_addr = lane_addr + sb * 128u;
        // pass A: row maximum (scaled to log2 units); the 32-column TMEM loads are software-pipelined and full tiles
        // skip the per-key predicates (c > 0, so the maximum commutes with the scale)
        float mx = -INFINITY;
        {
          auto max_chunk = [&](const uint32_t (&v)[32], int c4) {
            if (valid >= (c4 + 1) * 32) {
              float mr = -INFINITY;
#pragma unroll
              for (int i = 0; i < 32; i += 2) mr = fmaxf(mr, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
              mx = fmaxf(mx, mr * c);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                mx = fmaxf(mx, (c4 * 32 + i < valid) ? __uint_as_float(v[i]) * c : -INFINITY);
            }
          };
          uint32_t va[32], vb[32];
          tmem_ld_x32(s_addr, va);
          tmem_ld_wait();
          tmem_ld_x32(s_addr + 32, vb);
          max_chunk(va, 0);
          tmem_ld_wait();
          tmem_ld_x32(s_addr + 64, va);
          max_chunk(vb, 1);
          tmem_ld_wait();
          tmem_ld_x32(s_addr + 96, vb);
          max_chunk(va, 2);
          tmem_ld_wait();
          max_chunk(vb, 3);
        }
        const bool need = mx > m_used + FA_RESCALE_THRESHOLD;
        const float m_new = need ? mx : m_used;
        const float alpha = fast_exp2(m_used - m_new);  // 1 when unchanged, 0 on the first tile
        // PV of the previous tile must have finished: O is stable and the P buffer is free
        mbar_wait(&p_free, (t & 1u) ^ 1u);
        tc_fence_after_sync();
        if (j > 0 && __any_sync(0xffffffffu, need)) {
#pragma unroll 1
          for (int cc = 0; cc < DK / 16; ++cc) {
            uint32_t o[16];
            tmem_ld_x16(lane_addr + Cfg::O_COL + cc * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x16(lane_addr + Cfg::O_COL + cc * 16, o);
          }
          tmem_st_wait();
        }
        l *= alpha;
        m_used = m_new;
        // pass B: P = exp2(S * c - m) -> bf16 -> smem (SWIZZLE_128B K-major: row r, 16-byte piece j at (j ^ (r & 7)))
        const uint32_t p_row = sP + r * 128;
        {
          auto exp_chunk = [&](const uint32_t (&v)[32], int c4) {
            uint32_t pk[16];
            if (valid >= (c4 + 1) * 32) {
              const uint64_t c2 = f2_pack(c, c), nm2 = f2_pack(-m_used, -m_used);
              uint64_t ls2 = f2_pack(0.f, 0.f);
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float x0, x1;
                f2_unpack(f2_fma(f2_pack(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), c2, nm2), x0, x1);
                const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
                ls2 = f2_add(ls2, f2_pack(p0, p1));
                pk[i >> 1] = pack_bf16x2(p0, p1);
              }
              float la, lb;
              f2_unpack(ls2, la, lb);
              l += la + lb;
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float p0 = (c4 * 32 + i < valid) ? fast_exp2(fmaf(__uint_as_float(v[i]), c, -m_used)) : 0.f;
                const float p1 = (c4 * 32 + i + 1 < valid) ? fast_exp2(fmaf(__uint_as_float(v[i + 1]), c, -m_used)) : 0.f;
                l += p0 + p1;
                pk[i >> 1] = pack_bf16x2(p0, p1);
              }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int piece = c4 * 4 + g;  // 16-byte piece index inside the 256-byte P row
              const uint32_t addr = p_row + (piece >> 3) * (FA_BM * 128) + (((piece & 7) ^ (r & 7)) << 4);
              st_shared_v4(addr, pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
            }
          };
          uint32_t va[32], vb[32];
          tmem_ld_x32(s_addr, va);
          tmem_ld_wait();
          tmem_ld_x32(s_addr + 32, vb);
          exp_chunk(va, 0);
          tmem_ld_wait();
          tmem_ld_x32(s_addr + 64, va);
          exp_chunk(vb, 1);
          tmem_ld_wait();
          tmem_ld_x32(s_addr + 96, vb);
          exp_chunk(va, 2);
          tmem_ld_wait();
          exp_chunk(vb, 3);
        }
        tc_fence_before_sync();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&s_free[sb]);
          mbar_arrive(&p_ready);
        }
      }
      // epilogue: O / l -> bf16 -> global
      mbar_wait(&o_final, it & 1u);
      tc_fence_after_sync();
      const float inv = 1.0f / l;
      const int q_in_img = qb * FA_BM + r;
      const bool row_ok = q_in_img < p.nq;
      if (p.lse != nullptr && row_ok)
        p.lse[(static_cast<long long>(img) * p.nq + q_in_img) * p.heads + head] = m_used + log2f(l);
      __nv_bfloat16* orow = p.O + (static_cast<long long>(img) * p.nq + q_in_img) * p.ldo + head * D;
#pragma unroll 1
      for (int cc = 0; cc < DK / 16; ++cc) {
        uint32_t o[16];
        tmem_ld_x16(lane_addr + Cfg::O_COL + cc * 16, o);
        tmem_ld_wait();
        if (row_ok) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            pk[i] = pack_bf16x2(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv);
          uint4* dst = reinterpret_cast<uint4*>(orow + cc * 16);
          if (cc * 16 + 8 <= D) dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          if (cc * 16 + 16 <= D) dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// =====================================================================================================================
// Two-query-tile variant (head_dim 40): 256 query rows per work item, one softmax warpgroup per 128-row tile.
//   warp 0      TMA producer (both Q tiles once per item; K / V tiles through a 3-stage ring)
//   warp 1      MMA issuer: per KV tile  PV_A(j-1), S_A = Q_A K^T(j), PV_B(j-1), S_B = Q_B K^T(j)
//   warp 2      TMEM allocator
//   warps 4-7   softmax warpgroup A (rows 0-127),  warps 8-11  softmax warpgroup B (rows 128-255)
// While warpgroup A turns S_A(j) into P_A(j), the tensor pipe computes S_B(j) and warpgroup B is still busy with the
// previous tile, so MUFU / TMEM reads of one group hide the MMA round trip of the other.
// The softmax is single pass for every tile after the first: P = exp2(S c - m) with the running reference m; the row
// maximum is tracked on the side and a growth of more than 2^8 schedules a rescale of (O, l) for the NEXT tile, when
// PV of this tile has certainly finished (S_full(j+1) is committed after PV(j) in the same in-order MMA stream).  A
// growth beyond 2^64 (overflow territory) falls back to an immediate rescale + second pass.  O / l is exact for any
// reference m, so a pending rescale at the end of the row is simply dropped.
// =====================================================================================================================
constexpr int FA2_THREADS = 384;
constexpr float FA2_REDO_THRESHOLD = 64.0f;

template <int D>
struct Fa2Cfg {
  static constexpr int DK = (D + 15) / 16 * 16;
  static constexpr int QCH = (DK + 63) / 64;
  static constexpr int Q_BYTES = QCH * FA_BM * 128;     // one 128-row Q tile
  static constexpr int K_BYTES = QCH * FA_BN * 128;
  static constexpr int V_BYTES = QCH * FA_BN * 128;
  static constexpr int P_BYTES = 2 * FA_BM * 128;       // one 128 x 128 bf16 P tile
  static constexpr int KV_BYTES = K_BYTES + V_BYTES;
  static constexpr int STAGES = 3;
  static constexpr int SMEM_BYTES = 2 * Q_BYTES + 2 * P_BYTES + STAGES * KV_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;
  static constexpr int O_COL = 256;                      // S_A at 0, S_B at 128, O_A at 256, O_B at 256 + DK
  static_assert(QCH == 1, "two-tile variant is built for head_dim <= 64");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// VF16: V (and therefore P) are IEEE fp16 instead of bf16 -- the projection GEMM writes its V columns as fp16
// (FMC_GEMM_F16_TAIL).  The probabilities then come from ex2.approx.f16x2, TWO exponentials per MUFU operation: at
// head_dim 40 this kernel is MUFU bound, and fp16 P carries three more mantissa bits than bf16 P.
template <int D, bool VF16>
__global__ void __launch_bounds__(FA2_THREADS, 1)
spatial_attn2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, FaParams p) {
  pdl_launch_dependents();
  using Cfg = Fa2Cfg<D>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int DK = Cfg::DK;

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full, q_empty;
  __shared__ uint64_t kv_full[STAGES], kv_empty[STAGES];
  __shared__ uint64_t s_full[2], p_ready[2], o_final[2], o_free[2];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;                        // two Q tiles
  const uint32_t sP = sQ + 2 * Cfg::Q_BYTES;            // two P tiles
  const uint32_t sKV = sP + 2 * Cfg::P_BYTES;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_pairs = (p.nq + 2 * FA_BM - 1) / (2 * FA_BM);
  const int num_items = p.images * p.heads * q_pairs;
  const int T = p.kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&q_full, 1);
    mbar_init(&q_empty, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int w = 0; w < 2; ++w) {
      mbar_init(&s_full[w], 1);
      mbar_init(&p_ready[w], 4);
      mbar_init(&o_final[w], 1);
      mbar_init(&o_free[w], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_slot, Cfg::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();  // everything above (barriers, TMEM, descriptor prefetch) overlaps the previous kernel's tail

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    if (elect_one()) {
      uint32_t t = 0, it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int qp = item % q_pairs;
        const int head = (item / q_pairs) % p.heads;
        const int img = item / (q_pairs * p.heads);
        const int q_row0 = img * p.nq + qp * 2 * FA_BM;
        const int kv_row0 = (img / p.kv_div) * p.kv_stride;
        mbar_wait(&q_empty, (it & 1u) ^ 1u);
        mbar_arrive_expect_tx(&q_full, 2 * Cfg::Q_BYTES);
        tma_load_2d_a(sQ, &tmQ, &q_full, p.q_col0 + head * p.head_stride, q_row0);
        tma_load_2d_a(sQ + Cfg::Q_BYTES, &tmQ, &q_full, p.q_col0 + head * p.head_stride, q_row0 + FA_BM);
        for (int j = 0; j < T; ++j, ++t) {
          const int st = t % STAGES;
          const uint32_t ph = (t / STAGES) & 1u;
          mbar_wait(&kv_empty[st], ph ^ 1u);
          mbar_arrive_expect_tx(&kv_full[st], Cfg::KV_BYTES);
          const uint32_t sK = sKV + st * Cfg::KV_BYTES;
          tma_load_2d_a(sK, &tmK, &kv_full[st], p.k_col0 + head * p.head_stride, kv_row0 + j * FA_BN);
          tma_load_2d_a(sK + Cfg::K_BYTES, &tmV, &kv_full[st], p.v_col0 + head * D, kv_row0 + j * FA_BN);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ------------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(FA_BM, FA_BN);
      constexpr uint32_t idesc_o = VF16 ? umma_idesc_f16_bmn(FA_BM, DK) : umma_idesc_bf16_bmn(FA_BM, DK);
      uint32_t t = 0, it = 0;
      // O_w += P_w(tile u) V(tile u)
      auto issue_pv = [&](int w, uint32_t u, bool first_of_item) {
        mbar_wait(&p_ready[w], u & 1u);
        if (first_of_item) mbar_wait(&o_free[w], (it & 1u) ^ 1u);  // previous item's epilogue has read O_w
        tc_fence_after_sync();
        const uint32_t sV = sKV + (u % STAGES) * Cfg::KV_BYTES + Cfg::K_BYTES;
        const uint32_t sPw = sP + w * Cfg::P_BYTES;
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k) {
          const uint64_t da = umma_desc_k_sw128(sPw + (k >> 2) * (FA_BM * 128) + (k & 3) * 32);
          const uint64_t db = umma_desc_mn_sw128(sV + k * (16 * 128), FA_BN * 128, 1024);
          umma_bf16_ss(tmem_base + Cfg::O_COL + w * DK, da, db, idesc_o, (!first_of_item || k > 0) ? 1u : 0u);
        }
      };
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        mbar_wait(&q_full, it & 1u);
        tc_fence_after_sync();
        for (int j = 0; j < T; ++j, ++t) {
          const int st = t % STAGES;
          const uint32_t ph = (t / STAGES) & 1u;
          mbar_wait(&kv_full[st], ph);
          tc_fence_after_sync();
          const uint32_t sK = sKV + st * Cfg::KV_BYTES;
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            if (j >= 1) {
              issue_pv(w, t - 1, j == 1);
              if (w == 1) umma_commit(&kv_empty[(t - 1) % STAGES]);
            }
            // S_w is free: P_w(j-1) ready means group w has finished reading S_w(j-1) (for j = 0 the wait happened
            // when the last PV of the previous item was issued)
#pragma unroll
            for (int k = 0; k < DK / 16; ++k) {
              const uint64_t da = umma_desc_k_sw128(sQ + w * Cfg::Q_BYTES + k * 32);
              const uint64_t db = umma_desc_k_sw128(sK + k * 32);
              umma_bf16_ss(tmem_base + w * 128u, da, db, idesc_s, k > 0 ? 1u : 0u);
            }
            umma_commit(&s_full[w]);
          }
          if (j == T - 1) umma_commit(&q_empty);
        }
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          issue_pv(w, t - 1, T == 1);
          umma_commit(&o_final[w]);
        }
        umma_commit(&kv_empty[(t - 1) % STAGES]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------ softmax warpgroups + epilogue ------------------------------------
    const int w = (warp - 4) >> 2;  // 0: rows 0-127 (A), 1: rows 128-255 (B)
    const int q = warp & 3;
    const int r = q * 32 + lane;    // row inside this group's 128-row tile == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t s_addr = lane_addr + static_cast<uint32_t>(w) * 128u;
    const uint32_t o_addr = lane_addr + Cfg::O_COL + static_cast<uint32_t>(w) * DK;
    const uint32_t p_row = sP + w * Cfg::P_BYTES + r * 128;
    const float c = p.scale_log2e;
    uint32_t t = 0, it = 0;

    // P = exp2(S c - m) for the 128 keys of the tile -> bf16 -> smem; returns the row sum, tracks the scaled row max.
    // The four 32-column TMEM loads are software-pipelined: chunk c4 + 1 is in flight while chunk c4 is exponentiated
    // (-10 % at level 0).  Measured dead ends, kept out of the code (profiles/r01_spatial_attention_experiments.md):
    // splitting max / sum into independent chains (+6 %: more instructions, the warps are MUFU / issue bound), 64-key
    // half tiles that ping-pong inside a group (+11 %) and serving the two groups on demand instead of A, B (+14 %):
    // the two warps sharing a scheduler overlap best -- one in its FFMA phase, one in its MUFU phase -- when the
    // groups stay in step.
    auto exp_chunk = [&](const uint32_t (&v)[32], int c4, float m, int valid, float& mx, float& lsum) {
      uint32_t pk[16];
      if (valid >= FA_BN) {
        float mr = -INFINITY;  // raw (unscaled) maximum: c > 0, so max commutes with the scale
        // scale-and-shift and the row sum run on packed fp32 pairs (fma / add .f32x2): two fewer issue slots per pair
        const uint64_t c2 = f2_pack(c, c), nm2 = f2_pack(-m, -m);
        if constexpr (VF16) {
          uint32_t ha = 0u, hb = 0u;  // two fp16x2 partial sums (<= 16 terms of <= 1 each)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float v0 = __uint_as_float(v[i]), v1 = __uint_as_float(v[i + 1]);
            mr = fmaxf(mr, fmaxf(v0, v1));
            float x0, x1;
            f2_unpack(f2_fma(f2_pack(v0, v1), c2, nm2), x0, x1);
            const uint32_t ph = ex2_f16x2(pack_f16x2(x0, x1));
            if (i & 2) hb = add_f16x2(hb, ph);
            else ha = add_f16x2(ha, ph);
            pk[i >> 1] = ph;
          }
          lsum += sum_f16x2(ha) + sum_f16x2(hb);
        } else {
          uint64_t ls2 = f2_pack(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float v0 = __uint_as_float(v[i]), v1 = __uint_as_float(v[i + 1]);
            mr = fmaxf(mr, fmaxf(v0, v1));
            float x0, x1;
            f2_unpack(f2_fma(f2_pack(v0, v1), c2, nm2), x0, x1);
            const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
            ls2 = f2_add(ls2, f2_pack(p0, p1));
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          float la, lb;
          f2_unpack(ls2, la, lb);
          lsum += la + lb;
        }
        mx = fmaxf(mx, mr * c);
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const bool ok0 = c4 * 32 + i < valid, ok1 = c4 * 32 + i + 1 < valid;
          const float s0 = ok0 ? __uint_as_float(v[i]) * c : -INFINITY;
          const float s1 = ok1 ? __uint_as_float(v[i + 1]) * c : -INFINITY;
          mx = fmaxf(mx, fmaxf(s0, s1));
          const float p0 = ok0 ? fast_exp2(s0 - m) : 0.f, p1 = ok1 ? fast_exp2(s1 - m) : 0.f;
          lsum += p0 + p1;
          pk[i >> 1] = VF16 ? pack_f16x2(p0, p1) : pack_bf16x2(p0, p1);
        }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int piece = c4 * 4 + g;
        const uint32_t addr = p_row + (piece >> 3) * (FA_BM * 128) + (((piece & 7) ^ (r & 7)) << 4);
        st_shared_v4(addr, pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
      }
    };
    auto exp_pass = [&](float m, int valid, float& mx) -> float {
      float lsum = 0.f;
      uint32_t va[32], vb[32];
      tmem_ld_x32(s_addr, va);
      tmem_ld_wait();
      tmem_ld_x32(s_addr + 32, vb);
      exp_chunk(va, 0, m, valid, mx, lsum);
      tmem_ld_wait();
      tmem_ld_x32(s_addr + 64, va);
      exp_chunk(vb, 1, m, valid, mx, lsum);
      tmem_ld_wait();
      tmem_ld_x32(s_addr + 96, vb);
      exp_chunk(va, 2, m, valid, mx, lsum);
      tmem_ld_wait();
      exp_chunk(vb, 3, m, valid, mx, lsum);
      return lsum;
    };
    auto rescale_o = [&](float alpha) {
#pragma unroll 1
      for (int cc = 0; cc < DK / 16; ++cc) {
        uint32_t o[16];
        tmem_ld_x16(o_addr + cc * 16, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st_x16(o_addr + cc * 16, o);
      }
      tmem_st_wait();
    };

    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int qp = item % q_pairs;
      const int head = (item / q_pairs) % p.heads;
      const int img = item / (q_pairs * p.heads);
      float m_used = -INFINITY, l = 0.f;
      float pend_alpha = 1.0f, pend_m = 0.f;
      bool pend = false;
      for (int j = 0; j < T; ++j, ++t) {
        const int valid = p.nk - j * FA_BN;
        mbar_wait(&s_full[w], t & 1u);  // S_w(j) complete, hence PV_w(j-1) complete as well (in-order MMA stream)
        tc_fence_after_sync();
        if (j == 0) {
          // first tile: exact row maximum first
          float mx = -INFINITY;
#pragma unroll 1
          for (int c4 = 0; c4 < 4; ++c4) {
            uint32_t v[32];
            tmem_ld_x32(s_addr + c4 * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              mx = fmaxf(mx, (c4 * 32 + i < valid) ? __uint_as_float(v[i]) * c : -INFINITY);
          }
          m_used = mx;
          float dummy = -INFINITY;
          l = exp_pass(m_used, valid, dummy);
        } else {
          if (__any_sync(0xffffffffu, pend)) {
            rescale_o(pend_alpha);  // alpha = 1 for rows without a pending change
            l *= pend_alpha;
            if (pend) m_used = pend_m;
            pend = false;
            pend_alpha = 1.0f;
          }
          float mx = -INFINITY;
          float lt = exp_pass(m_used, valid, mx);
          if (__any_sync(0xffffffffu, mx > m_used + FA2_REDO_THRESHOLD)) {
            // overflow territory: rescale now (PV(j-1) is complete) and redo the tile against the new maximum
            const float m_new = fmaxf(m_used, mx);
            const float alpha = fast_exp2(m_used - m_new);
            rescale_o(alpha);
            l *= alpha;
            m_used = m_new;
            float dummy = -INFINITY;
            lt = exp_pass(m_used, valid, dummy);
          } else if (mx > m_used + FA_RESCALE_THRESHOLD) {
            pend = true;
            pend_m = mx;
            pend_alpha = fast_exp2(m_used - mx);
          }
          l += lt;
        }
        tc_fence_before_sync();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[w]);
      }
      // epilogue: O_w / l -> bf16 -> global  (O and l share the reference m_used; a pending rescale is irrelevant)
      mbar_wait(&o_final[w], it & 1u);
      tc_fence_after_sync();
      const float inv = 1.0f / l;
      const int q_in_img = qp * 2 * FA_BM + w * FA_BM + r;
      const bool row_ok = q_in_img < p.nq;
      if (p.lse != nullptr && row_ok)
        p.lse[(static_cast<long long>(img) * p.nq + q_in_img) * p.heads + head] = m_used + log2f(l);
      __nv_bfloat16* orow = p.O + (static_cast<long long>(img) * p.nq + q_in_img) * p.ldo + head * D;
#pragma unroll 1
      for (int cc = 0; cc < DK / 16; ++cc) {
        uint32_t o[16];
        tmem_ld_x16(o_addr + cc * 16, o);
        tmem_ld_wait();
        if (row_ok) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            pk[i] = pack_bf16x2(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv);
          uint4* dst = reinterpret_cast<uint4*>(orow + cc * 16);
          if (cc * 16 + 8 <= D) dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          if (cc * 16 + 16 <= D) dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[w]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int D, bool VF16>
static int launch_fa2(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const FaParams& p,
                      cudaStream_t stream) {
  using Cfg = Fa2Cfg<D>;
  static unsigned long long attr_devs = 0;  // per device: the attribute belongs to the (device, function) pair
  if (first_use_on_this_device(&attr_devs)) {
    FMC_CUDA_OK(cudaFuncSetAttribute(spatial_attn2_kernel<D, VF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  const int items = p.images * p.heads * ((p.nq + 2 * FA_BM - 1) / (2 * FA_BM));
  const int grid = items < device_sm_count() ? items : device_sm_count();
  launch_k(spatial_attn2_kernel<D, VF16>, dim3(grid), dim3(FA2_THREADS), Cfg::SMEM_BYTES, stream, tmQ, tmK, tmV, p);
  return check_launch("spatial_attn2_kernel");
}

template <int D>
static int launch_fa(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const FaParams& p,
                     cudaStream_t stream) {
  using Cfg = FaCfg<D>;
  static unsigned long long attr_devs = 0;  // per device: the attribute belongs to the (device, function) pair
  if (first_use_on_this_device(&attr_devs)) {
    FMC_CUDA_OK(cudaFuncSetAttribute(spatial_attn_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  const int items = p.images * p.heads * p.q_blocks;
  const int grid = items < device_sm_count() ? items : device_sm_count();
  launch_k(spatial_attn_kernel<D>, dim3(grid), dim3(FA_THREADS), Cfg::SMEM_BYTES, stream, tmQ, tmK, tmV, p);
  return check_launch("spatial_attn_kernel");
}

}  // namespace fmc

namespace fmc {
int cross_attention_short_keys(const void* Q, long long ldq, int q_col0, long long q_rows, const void* K, long long ldk,
                               int k_col0, const void* V, long long ldv, int v_col0, long long kv_rows, int head_stride,
                               void* O, long long ldo, int images, int heads, int head_dim, int nq, int nk, int kv_div,
                               int kv_stride, float scale, cudaStream_t stream);
}

using namespace fmc;

static int spatial_attn_impl(const void* Q, long long ldq, int q_col0, long long q_rows, const void* K, long long ldk,
                             int k_col0, const void* V, long long ldv, int v_col0, long long kv_rows, int head_stride,
                             void* O, long long ldo, int images, int heads, int head_dim, int nq, int nk, int kv_div,
                             int kv_stride, float scale, bool v_f16, float* lse, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(Q && K && V && O, FMC_ERR_ARG, "fmc_spatial_attn_bf16: null operand");
  FMC_REQUIRE(head_dim == 40 || head_dim == 80 || head_dim == 160, FMC_ERR_SHAPE,
              "fmc_spatial_attn_bf16: head_dim %d not in {40, 80, 160}", head_dim);
  const int dk = (head_dim + 15) / 16 * 16;
  FMC_REQUIRE(head_stride >= dk, FMC_ERR_SHAPE,
              "fmc_spatial_attn_bf16: head_stride %d must cover the zero-padded head width %d", head_stride, dk);
  FMC_REQUIRE(images > 0 && heads > 0 && nq > 0 && nk > 0 && kv_div > 0, FMC_ERR_SHAPE,
              "fmc_spatial_attn_bf16: empty problem");
  FMC_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && q_col0 % 8 == 0 && k_col0 % 8 == 0 &&
                  v_col0 % 8 == 0,
              FMC_ERR_SHAPE, "fmc_spatial_attn_bf16: strides / column offsets must be multiples of 8 elements");
  FMC_REQUIRE((reinterpret_cast<uintptr_t>(O) & 15) == 0, FMC_ERR_SHAPE, "fmc_spatial_attn_bf16: O not 16-byte aligned");

  // short-key cross-attention (text: 77 keys): K / V stay resident per (kv group, head), the query tiles stream
  // (attn_cross.cu); FMC_CROSS_GENERIC=1 keeps the generic flash kernel
  static const bool cross_generic = getenv("FMC_CROSS_GENERIC") != nullptr;
  if (!cross_generic && !v_f16 && lse == nullptr && nk <= 80 && kv_stride >= 80 && images % kv_div == 0 &&
      (images / kv_div) * static_cast<long long>(kv_stride) <= kv_rows)
    return cross_attention_short_keys(Q, ldq, q_col0, q_rows, K, ldk, k_col0, V, ldv, v_col0, kv_rows, head_stride, O, ldo,
                                      images, heads, head_dim, nq, nk, kv_div, kv_stride, scale, stream);

  CUtensorMap tmQ, tmK, tmV;
  const uint32_t box[2] = {64, 128};
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(ldq), static_cast<uint64_t>(q_rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldq) * 2};
    int rc = make_tmap_bf16(&tmQ, Q, 2, dims, strides, box, true);
    if (rc != FMC_OK) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(ldk), static_cast<uint64_t>(kv_rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldk) * 2};
    int rc = make_tmap_bf16(&tmK, K, 2, dims, strides, box, true);
    if (rc != FMC_OK) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(ldv), static_cast<uint64_t>(kv_rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldv) * 2};
    int rc = make_tmap_bf16(&tmV, V, 2, dims, strides, box, true);
    if (rc != FMC_OK) return rc;
  }
  FaParams p{};
  p.heads = heads;
  p.nq = nq;
  p.nk = nk;
  p.kv_div = kv_div;
  p.kv_stride = kv_stride;
  p.head_stride = head_stride;
  p.q_col0 = q_col0;
  p.k_col0 = k_col0;
  p.v_col0 = v_col0;
  p.images = images;
  p.q_blocks = ceil_div(nq, FA_BM);
  p.kv_tiles = ceil_div(nk, FA_BN);
  p.scale_log2e = scale * 1.4426950408889634f;
  p.O = static_cast<__nv_bfloat16*>(O);
  p.ldo = ldo;
  p.lse = lse;
  FMC_REQUIRE(!v_f16 || head_dim == 40, FMC_ERR_SHAPE, "fp16 V is implemented for head_dim 40 only (got %d)", head_dim);
  switch (head_dim) {
    case 40: return v_f16 ? launch_fa2<40, true>(tmQ, tmK, tmV, p, stream) : launch_fa2<40, false>(tmQ, tmK, tmV, p, stream);
    case 80: return launch_fa<80>(tmQ, tmK, tmV, p, stream);
    default: return launch_fa<160>(tmQ, tmK, tmV, p, stream);
  }
}

extern "C" int fmc_spatial_attn_bf16(const void* Q, long long ldq, int q_col0, long long q_rows, const void* K,
                                     long long ldk, int k_col0, const void* V, long long ldv, int v_col0,
                                     long long kv_rows, int head_stride, void* O, long long ldo, int images, int heads,
                                     int head_dim, int nq, int nk, int kv_div, int kv_stride, float scale,
                                     void* stream_) {
  return spatial_attn_impl(Q, ldq, q_col0, q_rows, K, ldk, k_col0, V, ldv, v_col0, kv_rows, head_stride, O, ldo, images,
                           heads, head_dim, nq, nk, kv_div, kv_stride, scale, false, nullptr, stream_);
}

extern "C" int fmc_spatial_attn_vf16(const void* Q, long long ldq, int q_col0, long long q_rows, const void* K,
                                     long long ldk, int k_col0, const void* V, long long ldv, int v_col0,
                                     long long kv_rows, int head_stride, void* O, long long ldo, int images, int heads,
                                     int head_dim, int nq, int nk, int kv_div, int kv_stride, float scale,
                                     void* stream_) {
  return spatial_attn_impl(Q, ldq, q_col0, q_rows, K, ldk, k_col0, V, ldv, v_col0, kv_rows, head_stride, O, ldo, images,
                           heads, head_dim, nq, nk, kv_div, kv_stride, scale, true, nullptr, stream_);
}

extern "C" int fmc_spatial_attn_lse_bf16(const void* Q, long long ldq, int q_col0, long long q_rows, const void* K,
                                         long long ldk, int k_col0, const void* V, long long ldv, int v_col0,
                                         long long kv_rows, int head_stride, void* O, long long ldo, float* lse, int images,
                                         int heads, int head_dim, int nq, int nk, int kv_div, int kv_stride, float scale,
                                         void* stream_) {
  FMC_REQUIRE(lse != nullptr, FMC_ERR_ARG, "fmc_spatial_attn_lse_bf16: null lse");
  return spatial_attn_impl(Q, ldq, q_col0, q_rows, K, ldk, k_col0, V, ldv, v_col0, kv_rows, head_stride, O, ldo, images,
                           heads, head_dim, nq, nk, kv_div, kv_stride, scale, false, lse, stream_);
}
