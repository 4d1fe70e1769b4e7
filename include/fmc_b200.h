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

#define FMC_B200_ABI_VERSION 3

/* flags for fmc_gemm_bf16 */
#define FMC_GEMM_GEGLU 1   /* W rows interleaved (16 value, 16 gate); C[M, N/2] = value * gelu_erf(gate) */
#define FMC_GEMM_OUT_F32 2 /* C is fp32 instead of bf16 */
#define FMC_GEMM_F16_TAIL 4 /* output columns >= 32 * (flags >> 8) are written as IEEE fp16 (V of a fused q|k|v
                               projection feeding fmc_spatial_attn_vf16); bf16 output path only, no GEGLU / residual */

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

/* C = LayerNorm(A) Wo^T + b with the normalisation folded into the GEMM: A holds the UN-normalised rows; W = Wo scaled
 * by gamma per input channel (bf16); colsum[n] = sum_k W[n, k] of the bf16 weights; bias[n] = b[n] + sum_k Wo[n, k]
 * beta[k]; rowstats = float2 (mean, rstd) per row from fmc_rowstats_bf16.  Epilogue: acc * rstd - mean * rstd *
 * colsum[n] + bias[n] (then GEGLU if flagged).  Saves the LayerNorm write pass and the GEMM's read of it.
 * Replaces nn.LayerNorm + the following Linear at diffusers BasicTransformerBlock norm1 -> attn1.to_q|k|v,
 * norm2 -> attn2.to_q, norm3 -> ff.net.0 and fmc/models/motion_module.py:297 (ff_norm -> ff). */
int fmc_gemm_ln_bf16(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N,
                     int K, const float* bias, const float* colsum, const void* rowstats, int flags, int tile_n,
                     void* stream);

/* stats[row] = (mean, rstd = 1 / sqrt(var + eps)) over the C channels of each bf16 row (float2 per row). */
int fmc_rowstats_bf16(const void* x, long long ldx, void* stats, long long rows, int C, float eps, void* stream);

/* 3x3 convolution, padding 1, stride 1 or 2, on channels-last bf16 images as an implicit GEMM on the same tcgen05
 * kernel as fmc_gemm_bf16 (CTA-pair MMA): Out[n, oh, ow, :] = sum_{ky, kx, c} X[n, oh*s + ky - 1, ow*s + kx - 1, c] *
 * W[:, ky, kx, c] (+ bias[Cout]) (+ residual[n, oh, ow, :]).  X [images, H, Wd, Cin], Out / residual
 * [images, H/s, Wd/s, Cout] contiguous; W bf16 [Cout, 3, 3, Cin] (= torch conv weight in channels_last memory format).
 * The shifted input windows are fetched by 4-D TMA boxes whose out-of-image part is zero-filled: no im2col buffer.
 * Needs Cin % 64 == 0, Cout % 32 == 0 and an output width that divides 128; anything else is an error.
 * Replaces the per-frame convolutions of diffusers ResnetBlock2D / Downsample2D / Upsample2D and InflatedConv3d as
 * called at fmc/models/unet_blocks.py:402-404,420-422,686-688,701-704, fmc/models/resnet.py:16-24, and the 3x3
 * convolutions of the CameraEncoder / ObjectEncoder (fmc/models/pose_adaptor.py:102-135, fmc/adapter.py:64-98). */
int fmc_conv3x3_bf16(const void* X, const void* W, void* Out, const float* bias, const void* residual, int images,
                     int H, int Wd, int Cin, int Cout, int stride, int tile_n, void* stream);

/* O[i] = softmax(Q[i] K[kv(i)]^T * scale) V[kv(i)] per head, flash-style on tcgen05 (scores never leave the SM).
 * Replaces head_to_batch_dim + get_attention_scores (baddbmm, softmax) + bmm + batch_to_head_dim of the spatial
 * processors: fmc/models/attention_processor.py:148-154 (LoRAAttnProcessor, attn1 self / attn2 text cross).
 * Q rows: image i owns rows [i*nq, (i+1)*nq); kv group g = i / kv_div owns K/V rows [g*kv_stride, g*kv_stride + nk).
 * Head h of Q / K starts at column q_col0 / k_col0 + h*head_stride (head_stride = 48 with zero padding when
 * head_dim = 40); head h of V at column v_col0 + h*head_dim.  Q, K, V may alias one fused [token, q|k|v] buffer. */
int fmc_spatial_attn_bf16(const void* Q, long long ldq, int q_col0, long long q_rows, const void* K, long long ldk,
                          int k_col0, const void* V, long long ldv, int v_col0, long long kv_rows, int head_stride,
                          void* O, long long ldo, int images, int heads, int head_dim, int nq, int nk, int kv_div,
                          int kv_stride, float scale, void* stream);

/* Same with V (row layout as above) holding IEEE fp16 instead of bf16 -- written by fmc_gemm_bf16 with
 * FMC_GEMM_F16_TAIL.  head_dim 40 only.  The probabilities are then fp16 too and come from ex2.approx.f16x2 (two
 * exponentials per MUFU operation): this case is MUFU bound.  Q, K, O stay bf16. */
int fmc_spatial_attn_vf16(const void* Q, long long ldq, int q_col0, long long q_rows, const void* K, long long ldk,
                          int k_col0, const void* V, long long ldv, int v_col0, long long kv_rows, int head_stride,
                          void* O, long long ldo, int images, int heads, int head_dim, int nq, int nk, int kv_div,
                          int kv_stride, float scale, void* stream);
/* fmc_spatial_attn_bf16 that also writes lse[q_row, head] = row log-sum-exp of the scaled scores in log2 units (fp32,
 * [q_rows, heads]) for the training backward (fmc_attention_bwd_bf16 with lse_given = 1); always the generic flash kernel. */
int fmc_spatial_attn_lse_bf16(const void* Q, long long ldq, int q_col0, long long q_rows, const void* K, long long ldk,
                              int k_col0, const void* V, long long ldv, int v_col0, long long kv_rows, int head_stride,
                              void* O, long long ldo, float* lse, int images, int heads, int head_dim, int nq, int nk,
                              int kv_div, int kv_stride, float scale, void* stream);

/* Temporal self-attention core over the f frames of every latent position, directly on channels-last activations
 * (row of token (b, frame, hw) = (b*F + frame)*HW + hw; QKV row = [q | k | v] column blocks).
 * Replaces fmc/models/attention_processor.py:61-67 and :271-281 as reached from
 * fmc/models/motion_module.py:349-389 (TemporalSelfAttention), incl. the '(b h w) f c' rearranges of :218,:230. */
int fmc_temporal_attn_bf16(const void* QKV, long long ld, int q_col0, int k_col0, int v_col0, int head_stride,
                           void* O, long long ldo, int B, int F, int HW, int heads, int head_dim, float scale,
                           void* stream);

/* Fused q|k|v projection + temporal attention (channels = 320, 8 heads x 40): O = softmax((X Wq^T)(X Wk^T)^T * scale)
 * (X Wv^T) over the f frames of every latent position, one kernel, the [token, q|k|v] tensor never reaches HBM.
 * X, O: channels-last rows as for fmc_temporal_attn_bf16.  Wqkv: bf16 [8 * 128, 320], per head h the rows
 * [to_q rows of h (40) | to_k rows of h (40) | to_v rows of h (40) | 8 zero rows].
 * Replaces attn.to_q / to_k / to_v (bias-free) + the attention core of fmc/models/attention_processor.py:46-67
 * (AttnProcessor) and :259-281 (PoseAdaptorAttnProcessor, where X is the merged tensor of :257) as reached from
 * TemporalSelfAttention.forward, fmc/models/motion_module.py:349-389.  Other widths: error (use the un-fused pair). */
int fmc_temporal_qkv_attn_bf16(const void* X, long long ldx, const void* Wqkv, long long ldw, void* O, long long ldo,
                               int B, int F, int HW, int channels, int heads, float scale, void* stream);

/* Diagnostics for fmc_temporal_qkv_attn_bf16: `device_buffer` (4 * 64 * 8 int64, or NULL to switch off) receives
 * clock64() stamps of the pipeline events of CTA 0 -- [role][item][event], see csrc/temporal_fused.cu. */
int fmc_debug_set_timeline(void* device_buffer);

/* out = LayerNorm(x) * gamma + beta (+ pe[frame(row)]) in bf16; optionally out2 = that (in fp32) + add.
 * frame(row) = (row / HW) % F for channels-last [B, F, HW, C] rows.  One warp per row, fp32 statistics.
 * Replaces nn.LayerNorm at fmc/models/motion_module.py:289,297 and diffusers BasicTransformerBlock norm1-3,
 * PositionalEncoding.forward motion_module.py:319-321 (x + pe[:, :f]) and the `hidden_states + pose_feature`
 * operand of qkv_merge, fmc/models/attention_processor.py:257. */
int fmc_layernorm_bf16(const void* x, long long ldx, const float* gamma, const float* beta, float eps, void* out,
                       long long ldo, const float* pe, int F, int HW, const void* add, long long ldadd, void* out2,
                       long long ldo2, long long rows, int C, void* stream);

/* Per-image GroupNorm on channels-last x[images, HW, C] (+ channel bias rowbias[image / rowbias_div] added first)
 * (+ SiLU).
 * stats_ws: fp32 workspace of 2*images*(groups*ceil(HW/64) + C) floats (per-chunk partial sums, folded in a fixed
 * order -- deterministic, no atomics -- then one scale/shift pair per image and channel).  Replaces InflatedGroupNorm fmc/models/resnet.py:27-37
 * (motion_module.py:217), diffusers ResnetBlock2D.norm1/norm2 + SiLU (+ the time-embedding add between them),
 * Transformer2DModel.norm, and conv_norm_out + conv_act fmc/models/unet.py:1288-1292. */
int fmc_groupnorm_bf16(const void* x, long long ldx, const float* gamma, const float* beta, float eps, void* out,
                       long long ldo, float* stats_ws, int images, int HW, int C, int groups, int silu,
                       const float* rowbias, long long ldrb, int rowbias_div, void* stream);

/* Launch accounting only (never fails): how many kernels one fmc_groupnorm_bf16 call of this shape launches -- 1 when
 * the single-pass cluster kernel takes it (x read once: a cluster of CTAs holds one image's channel chunk in
 * registers and exchanges per-group partial sums through distributed shared memory), 3 for the partial / finalize /
 * apply form. */
int fmc_groupnorm_launches(int HW, int C, int groups);

/* out = a (+ b) (+ rowbias[row / rows_per_group]) (ReLU optional).  Residual adds, the ObjectEncoder feature
 * injection fmc/modified_modules.py:115-117, time-embedding broadcast, nn.ReLU of fmc/adapter.py:93. */
int fmc_add_bf16(const void* a, long long lda, const void* b, long long ldb, const float* rowbias, int rows_per_group,
                 long long ldrb, void* out, long long ldo, long long rows, int C, int relu, void* stream);

/* torch 'nearest' resize of channels-last [N, h, w, C] -> [N, oh, ow, C] (diffusers Upsample2D, called at
 * fmc/models/unet_blocks.py:701-704). */
int fmc_resize_nearest_bf16(const void* x, void* out, int N, int h, int w, int oh, int ow, int C, void* stream);

/* AvgPool2d(2) on channels-last (Downsample with use_conv=False, fmc/models/pose_adaptor.py:95, fmc/adapter.py:57). */
int fmc_avgpool2_bf16(const void* x, void* out, int N, int h, int w, int C, void* stream);

/* dst[r, 0:cols] = src[r, 0:cols] with row strides: the skip-connection channel concat torch.cat(dim=1) of
 * fmc/models/unet_blocks.py:660,786 written straight into the concatenated buffer. */
int fmc_copy2d_bf16(const void* src, long long lds, void* dst, long long ldd, long long rows, int cols, void* stream);

/* Layout / dtype conversion between the reference tensor layout [B, C, F, H, W] fp32 and channels-last bf16
 * [B, F, H, W, Cpad] (replaces the einops rearranges around every op, e.g. unet_blocks.py:402-409). */
int fmc_ncfhw_f32_to_cl_bf16(const float* x, void* out, int B, int C, int F, long long HW, int Cpad, void* stream);
int fmc_cl_bf16_to_ncfhw_f32(const void* x, long long ldc, float* out, int B, int C, int F, long long HW, void* stream);

/* fp32 -> bf16 (optional SiLU): `nonlinearity(temb)` of diffusers ResnetBlock2D, TimestepEmbedding.act. */
int fmc_cast_act_bf16(const float* x, void* out, long long n, int silu, void* stream);

/* diffusers Timesteps(dim, flip_sin_to_cos=True, freq_shift=0) (fmc/models/unet.py:1090): [cos | sin] in bf16. */
int fmc_timestep_embedding_bf16(const float* t, void* out, int B, int dim, void* stream);

/* Pluecker-ray embedding (o x d, d): ray_condition fmc/data/dataset.py:930-972 via to_plucker_embedding
 * train_cam_ctrl.py:77-90.  K[BF,4] = (fx,fy,cx,cy), c2w[BF,3,4].  _f32: out[BF,H,W,6] (the reference's return
 * layout).  _unshuffle_bf16: out[BF,H/8,W/8,384], channel comp*64 + dy*8 + dx = PixelUnshuffle(8) of
 * fmc/models/pose_adaptor.py:227-228 fused in (write-only, HBM bound). */
int fmc_plucker_f32(const float* K, const float* c2w, float* out, int BF, int H, int W, void* stream);
int fmc_plucker_unshuffle_bf16(const float* K, const float* c2w, void* out, int BF, int H, int W, void* stream);

/* ObjectEncoder input build, get_traj_features_v2 fmc/util.py:147-203: last object with mask > 0 wins per pixel;
 * channels = (info*m)*m (12) and m*m (1).  info[BF,n_obj,12], masks[BF,n_obj,H,W] fp32.
 * _f32: feat[BF,13,H,W] + mask[BF,H,W] in the reference layout (bit-exact).  _unshuffle_bf16: feat
 * [BF,H/8,W/8,832] with channel c*64 + dy*8 + dx (PixelUnshuffle(8) of fmc/adapter.py:163 fused) + mask[BF,H,W]. */
int fmc_traj_scatter_f32(const float* info, const float* masks, float* feat, float* mask_out, int BF, int n_obj, int H,
                         int W, void* stream);
int fmc_traj_scatter_unshuffle_bf16(const float* info, const float* masks, void* feat, float* mask_out, int BF,
                                    int n_obj, int H, int W, void* stream);

/* Gaussian sphere masks on the device, fmc/data/dataset.py:5350-5403 (`use_sphere_mask`): circles[BF, n_obj, 3] = (cx, cy,
 * r) of each object's minimum enclosing circle (r <= 0: object absent) -> masks[BF, n_obj, H, W] = exp(-0.5 (dist /
 * (r/2))^2) / max * disc(int(cx), int(cy), int(r)).  _circles_unshuffle_bf16: fmc_traj_scatter_unshuffle_bf16 with these
 * masks generated on the fly (no mask tensor is read; bit-identical to the two-step form). */
int fmc_sphere_mask_f32(const float* circles, float* masks, int BF, int n_obj, int H, int W, void* stream);
int fmc_traj_scatter_circles_unshuffle_bf16(const float* info, const float* circles, void* feat, float* mask_out, int BF,
                                            int n_obj, int H, int W, void* stream);

/* x * nearest-resized mask, fmc/adapter.py:175-177.  row_index[h] / col_index[w] map a level pixel to the
 * full-resolution mask pixel (composition of the iterated F.interpolate(mode='nearest') calls). */
int fmc_mask_modulate_bf16(const void* x, const float* mask, const int* row_index, const int* col_index, void* out,
                           int N, int h, int w, int C, int H, int W, void* stream);

/* CFG combine + DDIM(eta=0) update on fp32 latents: fmc/pipelines/pipeline_animation_cm_om.py:711-720 and
 * diffusers DDIMScheduler.step.  eps_cond = NULL disables guidance.  eps_out (optional) receives the combined eps. */
int fmc_cfg_ddim_step_f32(const float* eps_uncond, const float* eps_cond, float guidance_scale, const float* latents,
                          float* latents_out, float* eps_out, float alpha_t, float alpha_prev, long long n,
                          void* stream);

/* Multidiff window average + DDIM update for long clips, fmc/pipelines/pipeline_animation.py:669-702: window k covers
 * frames [k*stride, k*stride + L) with stride = L - multidiff_overlaps; per frame the (guided) predictions of the windows
 * covering it are averaged in the reference's order (`noise_full[win] += noise_pred / count[win]`, window by window), then
 * one DDIM (eta = 0) step updates the whole clip.  eps_windows [n_windows, (cfg ? 2b : b), C, L, HW] fp32, unconditional
 * half first; latents / latents_out [b, C, F_total, HW] with F_total = (n_windows - 1) * stride + L. */
int fmc_window_combine_ddim_f32(const float* eps_windows, int n_windows, int cfg, float guidance_scale,
                                const float* latents, float* latents_out, int b, int C, int F_total, long long HW, int L,
                                int stride, float alpha_t, float alpha_prev, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Backward of the hot path (csrc/backward.cu; train_cam_ctrl.py:586-665 `scaler.scale(loss).backward()`): activation
 * gradients through the frozen U-Net, parameter gradients for the trainable subset.  bf16 tensors in the forward's
 * layouts, fp32 arithmetic.  Linear layers: dX = dY W is fmc_gemm_bf16 against a transposed weight copy, dW = dY^T X is
 * fmc_gemm_bf16 (FMC_GEMM_OUT_F32) on operands transposed by fmc_transpose_bf16, bias gradients are fmc_colsum_f32.
 * --------------------------------------------------------------------------------------------------------------- */
int fmc_transpose_bf16(const void* x, long long ldx, void* out, long long ldo, long long rows, int cols, void* stream);

/* out[n] (+)= sum_m x[m, n] (x bf16 or fp32), deterministic two-stage; workspace: fmc_colsum_workspace_floats(rows, cols). */
int fmc_colsum_workspace_floats(long long rows, int cols);
int fmc_colsum_f32(const void* x, long long ldx, int x_is_bf16, float* out, float* workspace, long long rows, int cols,
                   int accumulate, void* stream);

/* LayerNorm backward (nn.LayerNorm of fmc/models/motion_module.py:289,297 and diffusers BasicTransformerBlock): x = the
 * forward's input rows, dy = gradient of the LayerNorm output (a positional-encoding add passes it through unchanged).
 * param_partials (optional): [fmc_layernorm_bwd_blocks(rows), 2, C] fp32, per-block sums of dy * xhat (d gamma) and dy
 * (d beta); fold them with fmc_colsum_f32. */
int fmc_layernorm_bwd_blocks(long long rows);
int fmc_layernorm_bwd_bf16(const void* x, long long ldx, const void* dy, long long lddy, const float* gamma, float eps,
                           void* dx, long long lddx, float* param_partials, long long rows, int C, void* stream);

/* GroupNorm backward with the forward's options (per-image channel bias added before the norm, SiLU after it), frozen
 * affine parameters; stats_ws: fmc_groupnorm_bwd_workspace_floats(images, HW, groups) floats (per-chunk partial sums of
 * the three passes: statistics, the two group means of the gradient, dx). */
long long fmc_groupnorm_bwd_workspace_floats(int images, int HW, int groups);
int fmc_groupnorm_bwd_bf16(const void* x, long long ldx, const void* dy, long long lddy, const float* gamma,
                           const float* beta, float eps, void* dx, long long lddx, float* stats_ws, int images, int HW, int C,
                           int groups, int silu, const float* rowbias, long long ldrb, int rowbias_div, void* stream);

/* GEGLU on the interleaved projection layout of FMC_GEMM_GEGLU (16 value | 16 gate column blocks): the training forward
 * keeps the projection (fmc_gemm_bf16 without the flag) and applies y = value * gelu_erf(gate) here; backward returns the
 * gradient of the projection. */
int fmc_geglu_fwd_bf16(const void* proj, long long ldp, void* y, long long ldy, long long rows, int H, void* stream);
int fmc_geglu_bwd_bf16(const void* proj, long long ldp, const void* dy, long long lddy, void* dproj, long long lddp,
                       long long rows, int H, void* stream);

int fmc_relu_bwd_bf16(const void* y, const void* dy, void* dx, long long n, void* stream);
/* backward of fmc_resize_nearest_bf16 for integer upsampling factors (oh % h == 0, ow % w == 0) */
int fmc_resize_nearest_bwd_bf16(const void* dy, void* dx, int N, int h, int w, int oh, int ow, int C, void* stream);
int fmc_avgpool2_bwd_bf16(const void* dy, void* dx, int N, int h, int w, int C, void* stream);

/* Attention backward (spatial self / text cross / temporal; row addressing and kv groups as fmc_attention_f32, head
 * layout as the bf16 forward: Q / K heads at col0 + h * head_stride, V / O / dO heads at col0 + h * head_dim).
 * Probabilities are recomputed flash-style; lse / dsum: fp32 [q_rows, heads] scratch (row log-sum-exp and sum_c dO O).
 * dK = dV = NULL skips the key / value gradients (text cross-attention: the text is frozen); they are implemented for
 * self-attention (kv_div = 1, nq = nk).  dQ / dK / dV use the Q / K / V layouts, so they can alias one [token, q|k|v]
 * gradient buffer (zero-initialised: padding columns are not written).
 * One entry point, three kernel families chosen by shape (same results within bf16 rounding; FMC_ATTN_BWD_SIMT=1 in the
 * environment forces the last one for A/B checks): tcgen05 kernels (csrc/attn_bwd_tc.cu) for inner = 1 at head_dim 40
 * (head_stride 48) / 80 / 160 -- self-attention and the dQ-only cross form; one warp per (sequence, head) for nq = nk = 16
 * self-attention (the temporal attentions; lse / dsum are not written there); CUDA-core kernels for everything else.
 * lse_given = 1: `lse` already holds the row log-sum-exp of the forward in log2 units (fmc_spatial_attn_lse_bf16); the
 * tcgen05 self-attention path at head_dim 40 / 80 then skips its first sweep over the keys, every other path ignores it
 * and recomputes. */
int fmc_attention_bwd_bf16(const void* Q, long long ldq, int q_col0, const void* K, long long ldk, int k_col0,
                           const void* V, long long ldv, int v_col0, int head_stride, const void* O, long long ldo,
                           const void* dO, long long lddo, void* dQ, long long lddq, int dq_col0, void* dK, long long lddk,
                           int dk_col0, void* dV, long long lddv, int dv_col0, float* lse, float* dsum, int lse_given,
                           int images, int heads, int head_dim, int nq, int nk, int kv_div, int kv_stride, int inner,
                           float scale, void* stream);

/* Training-step tail on flat fp32 buffers (train_cam_ctrl.py:647-665, train_cam_obj_ctrl.py:843-862: scaler.unscale_ ->
 * clip_grad_norm_ -> AdamW.step -- three passes over the trainable set in the reference), after the gradient all-reduce.
 * fmc_grad_norm_f32: state[0] = L2 norm of grad * inv_scale, state[1] = inv_scale * min(1, max_norm / (norm + 1e-6)) (max_norm
 * <= 0: no clipping), state[2] = 1 if any gradient is inf / nan; deterministic (two-stage sum, no atomics); workspace:
 * fmc_grad_norm_workspace_floats() floats.  fmc_adamw_step_f32: torch.optim.AdamW (no amsgrad) on grad * state[1], skipped
 * entirely when state[2] != 0 (GradScaler.step); state = NULL: plain step on the gradients as they are.  Nothing returns
 * to the host, so both can sit in a captured graph: fmc_grad_norm_f32 also counts the steps that were not skipped in
 * state[3], and fmc_adamw_step_f32 with step = 0 takes the bias-correction step count from there (and, with lr < 0, the
 * learning rate from state[4]) instead of from host arguments a graph would freeze.  state: 8 floats. */
int fmc_grad_norm_workspace_floats(void);
int fmc_grad_norm_f32(const float* grad, long long n, float inv_scale, float max_norm, float* workspace, float* state,
                      void* stream);
int fmc_adamw_step_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                       float beta1, float beta2, float eps, float weight_decay, int step, const float* state, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Reference-precision mode (BASELINE config 1, "output parity vs reference" at 1e-3 rel): fp32 activations between
 * kernels, linears on tcgen05.mma.kind::tf32, attention / norms / glue in fp32 (csrc/precise.cu).  Each entry point
 * shadows the bf16 one of the same name and replaces the same reference lines.
 * --------------------------------------------------------------------------------------------------------------- */

/* C[M,N] = A[M,K] W[N,K]^T (+ bias) (+ rowbias[row / rows_per_group]) (+ residual), all fp32, tf32 tensor-core products,
 * fp32 accumulation.  split = 1: operands are read as tf32 (10-bit mantissa).  split = 3: A is [M, 2K] = [tf32(a) | a -
 * tf32(a)] (fmc_split_tf32), W is [N, 2K] likewise; the kernel accumulates a_lo w_hi + a_hi w_lo + a_hi w_hi (fp32-class
 * products).  flags: FMC_GEMM_GEGLU only.  Same reference lines as fmc_gemm_bf16. */
int fmc_gemm_tf32(const float* A, long long lda, const float* W, long long ldw, float* C, long long ldc, int M, int N,
                  int K, const float* bias, const float* residual, long long ldr, const float* rowbias,
                  int rows_per_group, long long ldrb, int flags, int split, void* stream);

/* out[r, 0:K] = x[r] rounded to tf32 (nearest), out[r, K:2K] = x[r] - that (exact): operand form of split = 3. */
int fmc_split_tf32(const float* x, long long ldx, float* out, long long ldo, long long rows, int K, void* stream);

/* O = softmax(Q K^T * scale) V per (sequence, head) in fp32 (SIMT flash attention, head_dim 40 / 80 / 160).
 * Row of element t of sequence i: (i / inner) * len * inner + (i % inner) + t * inner.  inner = 1: spatial tokens of an
 * image (fmc_spatial_attn_bf16 semantics incl. kv_div / kv_stride for the text keys); inner = HW: the frame axis of
 * channels-last [B, F, HW, C] rows (fmc_temporal_attn_bf16 semantics).  Head h at column col0 + h * head_dim.
 * Replaces fmc/models/attention_processor.py:61-67,148-154,271-281. */
int fmc_attention_f32(const float* Q, long long ldq, int q_col0, const float* K, long long ldk, int k_col0,
                      const float* V, long long ldv, int v_col0, float* O, long long ldo, int images, int heads,
                      int head_dim, int nq, int nk, int kv_div, int kv_stride, int inner, float scale, void* stream);

/* fp32 forms of fmc_layernorm_bf16 / fmc_groupnorm_bf16 (stats_ws: 2 * images * groups floats). */
int fmc_layernorm_f32(const float* x, long long ldx, const float* gamma, const float* beta, float eps, float* out,
                      long long ldo, const float* pe, int F, int HW, const float* add, long long ldadd, float* out2,
                      long long ldo2, long long rows, int C, void* stream);
int fmc_groupnorm_f32(const float* x, long long ldx, const float* gamma, const float* beta, float eps, float* out,
                      long long ldo, float* stats_ws, int images, int HW, int C, int groups, int silu,
                      const float* rowbias, long long ldrb, int rowbias_div, void* stream);

/* out[(n, oy, ox), (ky, kx, c)] = x[n, oy*s + ky - 1, ox*s + kx - 1, c] (zero outside): the A operand of a 3x3,
 * padding-1 convolution as a GEMM against W[Cout, (ky, kx, cin)] -- same reference lines as fmc_conv3x3_bf16. */
int fmc_im2col3x3_f32(const float* x, float* out, int N, int H, int W, int C, int stride, void* stream);

/* fp32 forms of the glue kernels (C % 4 == 0 instead of % 8). */
int fmc_add_f32(const float* a, long long lda, const float* b, long long ldb, const float* rowbias, int rows_per_group,
                long long ldrb, float* out, long long ldo, long long rows, int C, int relu, void* stream);
int fmc_resize_nearest_f32(const float* x, float* out, int N, int h, int w, int oh, int ow, int C, void* stream);
int fmc_avgpool2_f32(const float* x, float* out, int N, int h, int w, int C, void* stream);
int fmc_copy2d_f32(const float* src, long long lds, float* dst, long long ldd, long long rows, int cols, void* stream);
int fmc_ncfhw_f32_to_cl_f32(const float* x, float* out, int B, int C, int F, long long HW, int Cpad, void* stream);
int fmc_cl_f32_to_ncfhw_f32(const float* x, long long ldc, float* out, int B, int C, int F, long long HW, void* stream);
int fmc_silu_f32(const float* x, float* out, long long n, void* stream);
int fmc_timestep_embedding_f32(const float* t, float* out, int B, int dim, void* stream);
int fmc_mask_modulate_f32(const float* x, const float* mask, const int* row_index, const int* col_index, float* out,
                          int N, int h, int w, int C, int H, int W, void* stream);

/* Weight gradient of a linear layer y = x W^T: dW [M, N] fp32 (+)= dY^T X with dY [T, M], X [T, N] bf16 row-major (M, N and
 * the row strides multiples of 8).  Both operands go to tcgen05 MN-major straight from the row-major tensors (no
 * transposed copies), the token axis is split across CTAs, the fp32 partials are folded in a fixed order (deterministic).
 * workspace: fmc_wgrad_workspace_floats(T, M, N) floats (may be NULL when that many splits is 1 and accumulate = 0).
 * Replaces the parameter-gradient half of `scaler.scale(loss).backward()` for the trainable linears
 * (train_cam_ctrl.py:648, train_cam_obj_ctrl.py:857). */
long long fmc_wgrad_workspace_floats(long long T, int M, int N);
int fmc_wgrad_bf16(const void* dY, long long lddy, const void* X, long long ldx, float* dW, long long lddw, float* workspace,
                   long long T, int M, int N, int accumulate, void* stream);

/* ---- pipeline edges: VAE decode / encode and the CLIP text encoder -----------------------------------
 * Their convolutions, linears, GroupNorm and LayerNorm run on the entry points above; these are the remaining pieces.
 * `is_f32` / `out_is_f32`: activation dtype of the call (0 = bf16, the product mode; 1 = fp32, reference-precision mode). */

/* out[r, :] = softmax(scale * scores[r, :]) over n columns (n % 4 == 0, n <= 4096); scores fp32 (fmc_gemm_bf16 with
 * FMC_GEMM_OUT_F32).  Replaces get_attention_scores' softmax of the diffusers Attention in the VAE mid block (1 head of
 * 512 over h*w tokens), called from vae.decode at fmc/pipelines/pipeline_animation.py:472 and vae.encode at
 * train_cam_ctrl.py:544. */
int fmc_softmax_rows(const float* scores, long long lds, void* out, long long ldo, int out_is_f32, long long rows, int n,
                     float scale, void* stream);

/* Multi-head self-attention for short sequences (<= 128 tokens, head_dim in {32, 64, 96, 128}), causal or not, on the
 * fused projection buffer [tokens, ld] with q / k / v at their column offsets and heads head_dim apart; out [tokens, ldo].
 * Replaces transformers CLIPAttention (causal mask, 77 tokens, 12 heads of 64) behind self.text_encoder(...) at
 * fmc/pipelines/pipeline_animation.py:506-510, :546-550 and train_cam_ctrl.py:557-561. */
int fmc_small_mha(const void* qkv, long long ld, int q_col0, int k_col0, int v_col0, void* out, long long ldo, int is_f32,
                  int seqs, int tokens_per_seq, int heads, int head_dim, float scale, int causal, void* stream);

/* out = x * sigmoid(1.702 x): transformers `quick_gelu`, the activation of CLIPMLP. */
int fmc_quick_gelu(const void* x, long long ldx, void* out, long long ldo, int is_f32, long long rows, int cols, void* stream);

/* out[b t, :] = token_table[ids[b t], :] + position_table[t, :] (fp32 tables, ids int64 clamped into the vocabulary):
 * transformers CLIPTextEmbeddings.forward. */
int fmc_embed_tokens(const long long* ids, const float* token_table, const float* position_table, void* out, long long ldo,
                     int out_is_f32, long long tokens, int tokens_per_seq, int C, int vocab, void* stream);

/* diffusers DiagonalGaussianDistribution.sample on channels-last moment rows [N HW, ldm] (mean | logvar, z channels each):
 * out [N, z, HW] fp32 = (mean + exp(0.5 clamp(logvar, -30, 20)) * noise) * out_scale; noise in the output layout, NULL =
 * the distribution's mode.  `vae.encode(x).latent_dist.sample() * 0.18215` of train_cam_ctrl.py:544-545 in one pass. */
int fmc_vae_sample_f32(const void* moments, long long ldm, int is_f32, const float* noise, float* out, int N, int z,
                       long long HW, float out_scale, void* stream);

/* Channels-last image rows [(b f) HW, ldc] -> video [b, C, f, HW] fp32 = clamp(x * mul + add, lo, hi): the rearrange +
 * `(video / 2 + 0.5).clamp(0, 1)` + `.float()` tail of decode_latents, fmc/pipelines/pipeline_animation.py:474-477. */
int fmc_cl_to_video_f32(const void* x, long long ldc, int is_f32, float* out, int B, int C, int F, long long HW, float mul,
                        float add, float lo, float hi, void* stream);

/* ---- relative-pose algebra on the device (fp64 like the reference's numpy) ------------------------------
 * A pose is a row-major 3x4 block [R | T] at the head of a record of `stride` doubles (12: 3x4 storage, 16: 4x4).
 * fmc_pose_relative_to_first_f64: create_relative_matrix_of_cam_list, fmc/data/utils.py:148-163 -- out[clip, f] =
 *   [R_f^T R_0 | R_f^T (T_0 - T_f) / scale_T], out[clip, 0] = eye(3, 4) exactly.
 * fmc_pose_absolute_from_relative_f64: create_absolute_matrix_from_ref_cam_list, :167-183 -- out[clip, f] = rows 0..2 of
 *   first[clip] (4x4) @ inv([rel[clip, f] with T * scale_T; 0 0 0 1]), out[clip, 0] = first[clip] rows 0..2 (the reference
 *   asserts 16 frames; any count works here).
 * fmc_pose_objects_relative_f64: create_relative_matrix_of_two_torch_matrix, :185-200 -- n object poses per camera pose,
 *   out[set, i] = [R_i^T R_cam | R_i^T (T_cam - T_obj0) / scale_T]: the translation of OBJECT 0 of the set is the one
 *   subtracted for every object, which is what the reference's stacked np.dot(...)[..., 0, 0] evaluates to. */
int fmc_pose_relative_to_first_f64(const double* poses, long long pose_stride, double* out, int clips, int frames,
                                   double scale_T, void* stream);
int fmc_pose_absolute_from_relative_f64(const double* first, const double* rel, double* out, int clips, int frames,
                                        double scale_T, void* stream);
int fmc_pose_objects_relative_f64(const double* cam, long long cam_stride, const double* obj, long long obj_stride, double* out,
                                  int sets, int n, double scale_T, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FMC_B200_H_ */
