// Temporal attention core: for every latent position (b, y, x) and head, softmax(q k^T * scale) v over the f frames.
// Replaces head_to_batch_dim -> baddbmm -> softmax -> bmm -> batch_to_head_dim of the reference temporal processors
// (fmc/models/attention_processor.py:61-67 AttnProcessor, :271-281 PoseAdaptorAttnProcessor, called from
// fmc/models/motion_module.py:349-389) including the three permute copies and the materialised
// [(b h w) * 8, f, f] score tensor.
//
// Activations stay channels-last [B, f, h, w, C]; a temporal sequence is a strided set of rows (frame stride = h*w
// rows), gathered by TMA.  f (4..32) is far below the 128-row tcgen05 tile, so 128/f sequences are packed into one
// tile, the 128x128 score tile is computed on the tensor cores and only its block diagonal is kept: each softmax
// thread (= query row = TMEM lane) reads the 32-column block of its warp and masks the other sequences.
//   warp 0 TMA (3-D maps, one box of f rows per sequence), warp 1 MMA issuer, warp 2 TMEM alloc,
//   warps 4-7 softmax + epilogue, software-pipelined over (tile, head) items with S / O double buffered.
// P never touches shared memory: it is written back to tensor memory as packed bf16 in place over the scores and read
// as the A operand of the PV MMA (tcgen05.mma with A in TMEM); the 64 KB this frees let head width 80 run with two
// Q|K|V stages (loads of item n+1 behind the MMAs of item n) instead of one.
// V is used as stored ([token, channel], MN-major B operand of the PV MMA).
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fmc {

constexpr int TA_THREADS = 256;

struct TaParams {
  int B, F, F_log2, HW, heads;
  int seqs_per_tile;  // 128 / F
  int tiles_per_b;    // ceil(HW / seqs_per_tile)
  int head_stride;    // between heads in Q / K columns (48 when d = 40: zero padded)
  int q_col0, k_col0, v_col0;
  float scale_log2e;
  __nv_bfloat16* O;
  long long ldo;
};

template <int D>
struct TaCfg {
  static constexpr int DK = (D + 15) / 16 * 16;
  static constexpr int QCH = (DK + 63) / 64;
  static constexpr int T_BYTES = QCH * 128 * 128;   // one of Q / K / V for one (tile, head)
  static constexpr int ITEM_BYTES = 3 * T_BYTES;
  static constexpr int STAGES = (QCH == 1) ? 4 : (QCH == 2 ? 2 : 1);  // 227 KB smem: 144 KB per item at d = 160
  static constexpr int SMEM_BYTES = STAGES * ITEM_BYTES + 1024;
  static constexpr bool O_DOUBLE = (256 + 2 * DK) <= 512;
  static constexpr int O_COL0 = 256;
  static constexpr int O_COL1 = O_DOUBLE ? 256 + DK : 256;
};

__device__ __forceinline__ float ta_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
__global__ void __launch_bounds__(TA_THREADS, 1)
temporal_attn_kernel(const __grid_constant__ CUtensorMap tmQKV, TaParams p) {
  pdl_launch_dependents();
  using Cfg = TaCfg<D>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int DK = Cfg::DK;
  constexpr int QCH = Cfg::QCH;

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t in_full[STAGES], in_empty[STAGES];
  __shared__ uint64_t s_full[2], p_ready[2], o_full[2], o_free[2];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sIn = smem_base;                   // STAGES x (Q | K | V)
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_items = p.B * p.tiles_per_b * p.heads;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&in_full[s], 1);
      mbar_init(&in_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 4);
      mbar_init(&o_full[s], 1);
      mbar_init(&o_free[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();  // everything above (barriers, TMEM, descriptor prefetch) overlaps the previous kernel's tail

  // item -> (b, tile, head): heads innermost so that neighbouring CTAs touch the same rows
  auto decode = [&](int item, int& b, int& hw0, int& head) {
    head = item % p.heads;
    const int tile = (item / p.heads) % p.tiles_per_b;
    b = item / (p.heads * p.tiles_per_b);
    hw0 = tile * p.seqs_per_tile;
  };
  auto o_index = [](uint32_t n) -> uint32_t { return Cfg::O_DOUBLE ? (n & 1u) : 0u; };
  auto o_phase = [](uint32_t n) -> uint32_t { return Cfg::O_DOUBLE ? ((n >> 1) & 1u) : (n & 1u); };

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    if (elect_one()) {
      uint32_t n = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        int b, hw0, head;
        decode(item, b, hw0, head);
        const int st = n % STAGES;
        const uint32_t ph = (n / STAGES) & 1u;
        mbar_wait(&in_empty[st], ph ^ 1u);
        mbar_arrive_expect_tx(&in_full[st], Cfg::ITEM_BYTES);
        const uint32_t base = sIn + st * Cfg::ITEM_BYTES;
        const int cols[3] = {p.q_col0 + head * p.head_stride, p.k_col0 + head * p.head_stride, p.v_col0 + head * D};
#pragma unroll
        for (int which = 0; which < 3; ++which) {
          for (int c = 0; c < QCH; ++c) {
            const uint32_t dst = base + which * Cfg::T_BYTES + c * (128 * 128);
            for (int g = 0; g < p.seqs_per_tile; ++g) {
              // box = (64 columns, 1 position, F frames): the F rows of sequence (b, hw0 + g); out-of-range positions
              // are zero-filled by TMA
              asm volatile(
                  "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                  ::"r"(dst + g * p.F * 128), "l"(reinterpret_cast<uint64_t>(&tmQKV)), "r"(smem_u32(&in_full[st])),
                  "r"(cols[which] + c * 64), "r"(hw0 + g), "r"(b * p.F)
                  : "memory");
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ------------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
      constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(128, DK);
      auto issue_s = [&](uint32_t n) {
        const int st = n % STAGES;
        const uint32_t ph = (n / STAGES) & 1u;
        mbar_wait(&in_full[st], ph);
        tc_fence_after_sync();
        const uint32_t sQ = sIn + st * Cfg::ITEM_BYTES;
        const uint32_t sK = sQ + Cfg::T_BYTES;
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) {
          const uint64_t da = umma_desc_k_sw128(sQ + (k >> 2) * (128 * 128) + (k & 3) * 32);
          const uint64_t db = umma_desc_k_sw128(sK + (k >> 2) * (128 * 128) + (k & 3) * 32);
          umma_bf16_ss(tmem_base + (n & 1u) * 128u, da, db, idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[n & 1u]);
      };
      auto issue_pv = [&](uint32_t n) {
        const int st = n % STAGES;
        mbar_wait(&p_ready[n & 1u], (n >> 1) & 1u);
        mbar_wait(&o_free[o_index(n)], o_phase(n) ^ 1u);
        tc_fence_after_sync();
        const uint32_t sV = sIn + st * Cfg::ITEM_BYTES + 2 * Cfg::T_BYTES;
        const uint32_t o_col = o_index(n) ? Cfg::O_COL1 : Cfg::O_COL0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A = P(n): packed bf16 in the first 64 columns of score buffer n & 1 (S(n+2), the next writer of that
          // buffer, is issued behind this MMA by the same thread)
          const uint64_t db = umma_desc_mn_sw128(sV + k * (16 * 128), 128 * 128, 1024);
          asm volatile(
              "{\n\t"
              ".reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
              "}\n" ::"r"(tmem_base + o_col),
              "r"(tmem_base + (n & 1u) * 128u + 8u * k), "l"(db), "r"(idesc_o), "r"(k > 0 ? 1u : 0u)
              : "memory");
        }
        umma_commit(&o_full[o_index(n)]);
        umma_commit(&in_empty[st]);
      };
      uint32_t n = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        if constexpr (STAGES >= 2) {
          issue_s(n);
          if (n >= 1) issue_pv(n - 1);
        } else {
          if (n >= 1) issue_pv(n - 1);
          issue_s(n);
        }
      }
      if (n >= 1) issue_pv(n - 1);
    }
  } else if (warp >= 4) {
    // ------------------------------------ softmax + epilogue ------------------------------------
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c = p.scale_log2e;
    const int my_group = lane >> p.F_log2;  // sequence of this row inside the warp's 32-row block
    float inv_prev = 0.f;
    int prev_item = -1;
    auto epilogue = [&](uint32_t n, int item, float inv) {
      int b, hw0, head;
      decode(item, b, hw0, head);
      mbar_wait(&o_full[o_index(n)], o_phase(n));
      tc_fence_after_sync();
      const int seq = r >> p.F_log2, frame = r & (p.F - 1);
      const int hw = hw0 + seq;
      const bool ok = hw < p.HW;
      const uint32_t o_col = o_index(n) ? Cfg::O_COL1 : Cfg::O_COL0;
      __nv_bfloat16* orow = p.O + ((static_cast<long long>(b) * p.F + frame) * p.HW + hw) * p.ldo + head * D;
#pragma unroll 1
      for (int cc = 0; cc < DK / 16; ++cc) {
        uint32_t o[16];
        tmem_ld_x16(lane_addr + o_col + cc * 16, o);
        tmem_ld_wait();
        if (ok) {
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
      if (lane == 0) mbar_arrive(&o_free[o_index(n)]);
    };
    uint32_t n = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
      mbar_wait(&s_full[n & 1u], (n >> 1) & 1u);
      tc_fence_after_sync();
      uint32_t v[32];
      tmem_ld_x32(lane_addr + (n & 1u) * 128u + q * 32, v);
      tmem_ld_wait();
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float s = ((i >> p.F_log2) == my_group) ? __uint_as_float(v[i]) * c : -INFINITY;
        v[i] = __float_as_uint(s);
        mx = fmaxf(mx, s);
      }
      float l = 0.f;
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = ta_exp2(__uint_as_float(v[i]) - mx);
        const float p1 = ta_exp2(__uint_as_float(v[i + 1]) - mx);
        l += p0 + p1;
        pk[i >> 1] = pack_bf16x2(p0, p1);
      }
      // P (64 packed columns) over the scores: this warp's rows are non-zero only in its own 32-key block
      {
        const uint32_t zeros[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        const uint32_t p_addr = lane_addr + (n & 1u) * 128u;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (g == q) tmem_st_x16(p_addr + 16 * g, pk);
          else tmem_st_x16(p_addr + 16 * g, zeros);
        }
        tmem_st_wait();
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[n & 1u]);
      if (prev_item >= 0) epilogue(n - 1, prev_item, inv_prev);
      inv_prev = 1.0f / l;
      prev_item = item;
    }
    if (prev_item >= 0) epilogue(n - 1, prev_item, inv_prev);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int D>
static int launch_ta(const CUtensorMap& tm, const TaParams& p, cudaStream_t stream) {
  using Cfg = TaCfg<D>;
  static unsigned long long attr_devs = 0;  // per device: the attribute belongs to the (device, function) pair
  if (first_use_on_this_device(&attr_devs)) {
    FMC_CUDA_OK(cudaFuncSetAttribute(temporal_attn_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  const int items = p.B * p.tiles_per_b * p.heads;
  const int grid = items < device_sm_count() ? items : device_sm_count();
  launch_k(temporal_attn_kernel<D>, dim3(grid), dim3(TA_THREADS), Cfg::SMEM_BYTES, stream, tm, p);
  return check_launch("temporal_attn_kernel");
}

}  // namespace fmc

using namespace fmc;

extern "C" int fmc_temporal_attn_bf16(const void* QKV, long long ld, int q_col0, int k_col0, int v_col0,
                                      int head_stride, void* O, long long ldo, int B, int F, int HW, int heads,
                                      int head_dim, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(QKV && O, FMC_ERR_ARG, "fmc_temporal_attn_bf16: null operand");
  FMC_REQUIRE(head_dim == 40 || head_dim == 80 || head_dim == 160, FMC_ERR_SHAPE,
              "fmc_temporal_attn_bf16: head_dim %d not in {40, 80, 160}", head_dim);
  FMC_REQUIRE(F == 4 || F == 8 || F == 16 || F == 32, FMC_ERR_SHAPE,
              "fmc_temporal_attn_bf16: frame count %d not in {4, 8, 16, 32}", F);
  const int dk = (head_dim + 15) / 16 * 16;
  FMC_REQUIRE(head_stride >= dk, FMC_ERR_SHAPE, "fmc_temporal_attn_bf16: head_stride %d < padded head width %d",
              head_stride, dk);
  FMC_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && q_col0 % 8 == 0 && k_col0 % 8 == 0 && v_col0 % 8 == 0, FMC_ERR_SHAPE,
              "fmc_temporal_attn_bf16: strides / column offsets must be multiples of 8 elements");
  FMC_REQUIRE(B > 0 && HW > 0 && heads > 0, FMC_ERR_SHAPE, "fmc_temporal_attn_bf16: empty problem");

  CUtensorMap tm;
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(ld), static_cast<uint64_t>(HW), static_cast<uint64_t>(B) * F};
    const uint64_t strides[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(HW) * ld * 2};
    const uint32_t box[3] = {64, 1, static_cast<uint32_t>(F)};
    int rc = make_tmap_bf16(&tm, QKV, 3, dims, strides, box, true);
    if (rc != FMC_OK) return rc;
  }
  TaParams p{};
  p.B = B; p.F = F; p.HW = HW; p.heads = heads;
  p.F_log2 = (F == 4) ? 2 : (F == 8 ? 3 : (F == 16 ? 4 : 5));
  p.seqs_per_tile = 128 / F;
  p.tiles_per_b = ceil_div(HW, p.seqs_per_tile);
  p.head_stride = head_stride;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.O = static_cast<__nv_bfloat16*>(O);
  p.ldo = ldo;
  switch (head_dim) {
    case 40: return launch_ta<40>(tm, p, stream);
    case 80: return launch_ta<80>(tm, p, stream);
    default: return launch_ta<160>(tm, p, stream);
  }
}
