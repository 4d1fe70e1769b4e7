/* libfmc_b200 -- C ABI of the B200-native FMC denoising hot path.
 *
 * The reference (FudanCVL/SynFMC) is pure Python on top of torch/diffusers and has no FFI of its own;
 * every entry point below replaces the torch-eager arithmetic of the cited reference lines and is what
 * the Python mirror of `fmc.models` / `fmc.pipelines` (synfmc_b200/fmc/...) binds through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends in `_host`
 *   - the caller owns every buffer; nothing is allocated, nothing synchronises
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - bf16 tensors are row-major, "channels-last": activations are [B, f, h, w, C] so that the spatial
 *     view [(B f), (h w), C] and the temporal view [(B h w), f, C] (frame stride h*w*C) need no copies
 *   - return value 0 = success, negative = error (see fmc_last_error_string); an unsupported shape is an
 *     error, never a fallback
 */
#ifndef FMC_B200_H_
#define FMC_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define FMC_B200_ABI_VERSION 1

/* flags for fmc_gemm_bf16 */
#define FMC_GEMM_GEGLU 1   /* W rows interleaved (16 value, 16 gate); C[M, N/2] = value * gelu_erf(gate) */
#define FMC_GEMM_OUT_F32 2 /* C is fp32 instead of bf16 */

int fmc_abi_version(void);
const char* fmc_last_error_string(void);

/* C[M,N] = A[M,K] * W[N,K]^T (+ bias[N]) (+ rowbias[row / rows_per_group, :]) (+ residual[M,N]);
 * bf16 operands, fp32 accumulation in tensor memory (tcgen05), bf16 or fp32 output.
 * Replaces: attn.to_q/to_k/to_v/to_out[0] + folded Domain-LoRA (fmc/models/attention_processor.py:138-157),
 *           PoseAdaptorAttnProcessor.qkv_merge (attention_processor.py:257),
 *           TemporalTransformer3DModel.proj_in/proj_out (fmc/models/motion_module.py:219,228),
 *           diffusers FeedForward GEGLU + Linear (called at motion_module.py:297),
 *           Transformer2DModel proj_in/proj_out 1x1 convs (called at fmc/models/unet_blocks.py:407).
 * lda/ldw/ldc/ldr/ldrb are row strides in elements.  tile_n = 0 lets the library pick the N tile. */
int fmc_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N,
                  int K, const float* bias, const void* residual, long long ldr, const float* rowbias,
                  int rows_per_group, long long ldrb, int flags, int tile_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FMC_B200_H_ */
