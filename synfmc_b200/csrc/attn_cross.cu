// Text cross-attention (attn2 of diffusers BasicTransformerBlock with LoRAAttnProcessor,
// fmc/models/attention_processor.py:108-169): O = softmax(Q K_text^T * scale) V_text with at most 80 keys.
//
// The generic flash kernel (attn_spatial.cu) walks KV tiles per (image, head, query block); with 77 keys there is one
// KV tile, so every work item is a serial chain Q load -> S -> softmax -> PV -> epilogue with nothing to overlap
// (2.2 us per 128 queries).  Here the roles are turned around: the keys / values of one (clip, head) stay in shared
// memory and the QUERY tiles stream past them -- all frames of a clip share the same text, so one (clip, head) has
// f * nq / 128 query tiles (320 at level 0).  Two softmax warpgroups take alternate tiles, so the S MMA of tile i+2, the
// softmax of tile i+1 and the PV / epilogue of tile i overlap.
//   warp 0       TMA producer: K, V once per work item, Q tiles through a ring
//   warp 1       tcgen05.mma issuer: S(i) = Q_i K^T (N = 80), O(i) = P_i V with P read from TENSOR MEMORY (packed bf16,
//                written by the softmax warps in place over the scores)
//   warp 2       TMEM allocator
//   warps 4-7    softmax / epilogue group 0 (even tiles),  warps 8-11  group 1 (odd tiles)
// TMEM (512 columns): S / P 2 x 96 | O 2 x 160.
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fmc {

constexpr int CA_THREADS = 384;
constexpr int CA_KEYS = 80;  // key rows staged per (clip, head): 77 text tokens padded to a multiple of 16

struct CaParams {
  int heads, nk;
  int clip_rows;       // query rows per clip = frames * nq (contiguous)
  int tiles_per_clip;  // ceil(clip_rows / 128)
  int chunks;          // query-tile chunks per (clip, head)
  int chunk_tiles;     // tiles per chunk
  int clips;
  int kv_stride;       // rows between clips in K / V
  int head_stride, q_col0, k_col0, v_col0;
  float scale_log2e;
  __nv_bfloat16* O;
  long long ldo;
};

template <int D>
struct CaCfg {
  static constexpr int DK = (D + 15) / 16 * 16;
  static constexpr int QCH = (DK + 63) / 64;
  static constexpr int Q_BYTES = QCH * 128 * 128;
  static constexpr int KV_CH_BYTES = CA_KEYS * 128;  // one 64-column chunk of K (or V): 80 rows x 128 B
  static constexpr int KV_BYTES = 2 * QCH * KV_CH_BYTES;
  static constexpr int STAGES = QCH == 1 ? 6 : (QCH == 2 ? 4 : 3);
  static constexpr int SMEM_BYTES = KV_BYTES + STAGES * Q_BYTES + 1024;
  static constexpr uint32_t S_COL = 0;     // group g: columns 96 g .. 96 g + 79 (P packed bf16 over the first 40)
  static constexpr uint32_t O_COL = 192;   // group g: columns 192 + 160 g ..
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(KV_CH_BYTES % 1024 == 0, "SW128 tiles stay 1024-byte aligned");
};

__device__ __forceinline__ float ca_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ca_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
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
__device__ __forceinline__ void ca_tmem_st_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

template <int D>
__global__ void __launch_bounds__(CA_THREADS, 1)
cross_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, CaParams p) {
  pdl_launch_dependents();
  using Cfg = CaCfg<D>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int DK = Cfg::DK;
  constexpr int QCH = Cfg::QCH;

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t kv_full, kv_free;
  __shared__ uint64_t q_full[STAGES], q_empty[STAGES];
  __shared__ uint64_t s_full[2], p_ready[2], o_full[2], o_free[2];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = smem_base;
  const uint32_t sV = sK + QCH * Cfg::KV_CH_BYTES;
  const uint32_t sQ = smem_base + Cfg::KV_BYTES;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_items = p.clips * p.heads * p.chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&kv_full, 1);
    mbar_init(&kv_free, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_ready[g], 4);
      mbar_init(&o_full[g], 1);
      mbar_init(&o_free[g], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();

  // item -> (clip, head, first tile, number of tiles)
  auto decode = [&](int item, int& clip, int& head, int& tile0, int& ntiles) {
    const int chunk = item % p.chunks;
    head = (item / p.chunks) % p.heads;
    clip = item / (p.chunks * p.heads);
    tile0 = chunk * p.chunk_tiles;
    ntiles = min(p.chunk_tiles, p.tiles_per_clip - tile0);
    if (ntiles < 0) ntiles = 0;
  };

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    if (elect_one()) {
      uint32_t it = 0, t = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        int clip, head, tile0, ntiles;
        decode(item, clip, head, tile0, ntiles);
        mbar_wait(&kv_free, (it & 1u) ^ 1u);  // every MMA of the previous item has completed
        mbar_arrive_expect_tx(&kv_full, Cfg::KV_BYTES);
        for (int c = 0; c < QCH; ++c) {
          tma_load_2d_a(sK + c * Cfg::KV_CH_BYTES, &tmK, &kv_full, p.k_col0 + head * p.head_stride + c * 64,
                        clip * p.kv_stride);
          tma_load_2d_a(sV + c * Cfg::KV_CH_BYTES, &tmV, &kv_full, p.v_col0 + head * D + c * 64, clip * p.kv_stride);
        }
        for (int i = 0; i < ntiles; ++i, ++t) {
          const uint32_t st = t % STAGES;
          mbar_wait(&q_empty[st], ((t / STAGES) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&q_full[st], Cfg::Q_BYTES);
          const int row = clip * p.clip_rows + (tile0 + i) * 128;
          for (int c = 0; c < QCH; ++c)
            tma_load_2d_a(sQ + st * Cfg::Q_BYTES + c * (128 * 128), &tmQ, &q_full[st],
                          p.q_col0 + head * p.head_stride + c * 64, row);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ------------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, CA_KEYS);
      constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(128, DK);
      uint32_t it = 0, t = 0;  // t: global tile counter of this CTA (ring position); per group: n = tiles of that group
      uint32_t ng[2] = {0, 0};   // tiles issued so far per group (S side)
      uint32_t npv[2] = {0, 0};  // PVs issued so far per group
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        int clip, head, tile0, ntiles;
        decode(item, clip, head, tile0, ntiles);
        mbar_wait(&kv_full, it & 1u);
        tc_fence_after_sync();
        auto issue_s = [&](int i) {
          const uint32_t g = static_cast<uint32_t>(i) & 1u;
          const uint32_t st = (t + i) % STAGES;
          mbar_wait(&q_full[st], ((t + i) / STAGES) & 1u);
          tc_fence_after_sync();
          const uint32_t q = sQ + st * Cfg::Q_BYTES;
#pragma unroll
          for (int k = 0; k < DK / 16; ++k) {
            const uint64_t da = umma_desc_k_sw128(q + (k >> 2) * (128 * 128) + (k & 3) * 32);
            const uint64_t db = umma_desc_k_sw128(sK + (k >> 2) * Cfg::KV_CH_BYTES + (k & 3) * 32);
            umma_bf16_ss(tmem_base + Cfg::S_COL + g * 96u, da, db, idesc_s, k > 0 ? 1u : 0u);
          }
          umma_commit(&q_empty[st]);
          umma_commit(&s_full[g]);
          ++ng[g];
        };
        auto issue_pv = [&](int i) {
          const uint32_t g = static_cast<uint32_t>(i) & 1u;
          mbar_wait(&p_ready[g], npv[g] & 1u);
          mbar_wait(&o_free[g], (npv[g] & 1u) ^ 1u);  // the epilogue of this group's previous tile has read O_g
          tc_fence_after_sync();
#pragma unroll
          for (int k = 0; k < CA_KEYS / 16; ++k)
            ca_umma_ts(tmem_base + Cfg::O_COL + g * 160u, tmem_base + Cfg::S_COL + g * 96u + 8 * k,
                       umma_desc_mn_sw128(sV + k * (16 * 128), Cfg::KV_CH_BYTES, 1024), idesc_o, k > 0 ? 1u : 0u);
          umma_commit(&o_full[g]);
          ++npv[g];
        };
        // tile parity inside the item decides the group; items always start with group 0
        if (ntiles > 0) issue_s(0);
        if (ntiles > 1) issue_s(1);
        for (int i = 0; i < ntiles; ++i) {
          issue_pv(i);                       // reads P_g(i) in place of S_g(i)
          if (i + 2 < ntiles) issue_s(i + 2);  // overwrites S_g: issued behind PV_g(i) by the same thread
        }
        umma_commit(&kv_free);
        t += ntiles;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------ softmax + epilogue groups ------------------------------------
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t s_addr = lane_addr + Cfg::S_COL + static_cast<uint32_t>(g) * 96u;
    const uint32_t o_addr = lane_addr + Cfg::O_COL + static_cast<uint32_t>(g) * 160u;
    const float c = p.scale_log2e;
    uint32_t n = 0;  // tiles handled by this group so far (barrier phases)
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int clip, head, tile0, ntiles;
      decode(item, clip, head, tile0, ntiles);
      for (int i = g; i < ntiles; i += 2, ++n) {
        mbar_wait(&s_full[g], n & 1u);
        tc_fence_after_sync();
        uint32_t v[80];
        {
          uint32_t a[32], b[32], d[16];
          tmem_ld_x32(s_addr, a);
          tmem_ld_x32(s_addr + 32, b);
          tmem_ld_x16(s_addr + 64, d);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] = a[j]; v[32 + j] = b[j]; }
#pragma unroll
          for (int j = 0; j < 16; ++j) v[64 + j] = d[j];
        }
        // four independent chains for the row maximum and the row sum (a single dependent chain of 80 is ~650 clk)
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 80; ++j) {
          const float s = (j < p.nk) ? __uint_as_float(v[j]) * c : -INFINITY;
          v[j] = __float_as_uint(s);
          m4[j & 3] = fmaxf(m4[j & 3], s);
        }
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pk[40];
#pragma unroll
        for (int j = 0; j < 80; j += 2) {
          const float p0 = ca_exp2(__uint_as_float(v[j]) - mx), p1 = ca_exp2(__uint_as_float(v[j + 1]) - mx);
          l4[(j >> 1) & 3] += p0 + p1;
          pk[j >> 1] = pack_bf16x2(p0, p1);
        }
        const float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        {
          uint32_t w0[16], w1[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { w0[j] = pk[j]; w1[j] = pk[16 + j]; }
          tmem_st_x16(s_addr, w0);
          tmem_st_x16(s_addr + 16, w1);
          ca_tmem_st_x8(s_addr + 32, &pk[32]);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[g]);
        // epilogue of this tile: O / l -> bf16 -> global
        mbar_wait(&o_full[g], n & 1u);
        tc_fence_after_sync();
        const float inv = 1.0f / l;
        const int row_in_clip = (tile0 + i) * 128 + r;
        const bool row_ok = row_in_clip < p.clip_rows;
        __nv_bfloat16* orow = p.O + (static_cast<long long>(clip) * p.clip_rows + row_in_clip) * p.ldo + head * D;
#pragma unroll 1
        for (int cc = 0; cc < DK / 16; ++cc) {
          uint32_t o[16];
          tmem_ld_x16(o_addr + cc * 16, o);
          tmem_ld_wait();
          if (row_ok) {
            uint32_t ob[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              ob[j] = pack_bf16x2(__uint_as_float(o[2 * j]) * inv, __uint_as_float(o[2 * j + 1]) * inv);
            uint4* dst = reinterpret_cast<uint4*>(orow + cc * 16);
            if (cc * 16 + 8 <= D) dst[0] = make_uint4(ob[0], ob[1], ob[2], ob[3]);
            if (cc * 16 + 16 <= D) dst[1] = make_uint4(ob[4], ob[5], ob[6], ob[7]);
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[g]);
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

template <int D>
static int launch_ca(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CaParams& p,
                     cudaStream_t stream) {
  using Cfg = CaCfg<D>;
  static unsigned long long attr_devs = 0;  // per device: the attribute belongs to the (device, function) pair
  if (first_use_on_this_device(&attr_devs)) {
    FMC_CUDA_OK(cudaFuncSetAttribute(cross_attn_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  const int items = p.clips * p.heads * p.chunks;
  const int grid = items < device_sm_count() ? items : device_sm_count();
  FMC_CUDA_OK(launch_k(cross_attn_kernel<D>, dim3(grid), dim3(CA_THREADS), Cfg::SMEM_BYTES, stream, tmQ, tmK, tmV, p));
  return check_launch("cross_attn_kernel");
}

// Called by fmc_spatial_attn_bf16 when the problem is a short-key cross-attention (nk <= 80, all images of a kv group
// contiguous).  Same argument meaning as there.
int cross_attention_short_keys(const void* Q, long long ldq, int q_col0, long long q_rows, const void* K, long long ldk,
                               int k_col0, const void* V, long long ldv, int v_col0, long long kv_rows, int head_stride,
                               void* O, long long ldo, int images, int heads, int head_dim, int nq, int nk, int kv_div,
                               int kv_stride, float scale, cudaStream_t stream) {
  CUtensorMap tmQ, tmK, tmV;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(ldq), static_cast<uint64_t>(q_rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldq) * 2};
    const uint32_t box[2] = {64, 128};
    int rc = make_tmap_bf16(&tmQ, Q, 2, dims, strides, box, true);
    if (rc != FMC_OK) return rc;
  }
  const uint32_t kbox[2] = {64, CA_KEYS};
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(ldk), static_cast<uint64_t>(kv_rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldk) * 2};
    int rc = make_tmap_bf16(&tmK, K, 2, dims, strides, kbox, true);
    if (rc != FMC_OK) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(ldv), static_cast<uint64_t>(kv_rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldv) * 2};
    int rc = make_tmap_bf16(&tmV, V, 2, dims, strides, kbox, true);
    if (rc != FMC_OK) return rc;
  }
  CaParams p{};
  p.heads = heads;
  p.nk = nk;
  p.clips = images / kv_div;
  p.clip_rows = kv_div * nq;
  p.tiles_per_clip = ceil_div(p.clip_rows, 128);
  // one work item per CTA where possible: every extra item costs a K / V reload behind a drained MMA pipeline
  const int want = device_sm_count() / (p.clips * heads);
  p.chunks = want < 1 ? 1 : (want > p.tiles_per_clip ? p.tiles_per_clip : want);
  p.chunk_tiles = ceil_div(p.tiles_per_clip, p.chunks);
  p.chunks = ceil_div(p.tiles_per_clip, p.chunk_tiles);
  p.kv_stride = kv_stride;
  p.head_stride = head_stride;
  p.q_col0 = q_col0;
  p.k_col0 = k_col0;
  p.v_col0 = v_col0;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.O = static_cast<__nv_bfloat16*>(O);
  p.ldo = ldo;
  switch (head_dim) {
    case 40: return launch_ca<40>(tmQ, tmK, tmV, p, stream);
    case 80: return launch_ca<80>(tmQ, tmK, tmV, p, stream);
    default: return launch_ca<160>(tmQ, tmK, tmV, p, stream);
  }
}

}  // namespace fmc
