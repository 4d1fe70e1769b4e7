"""Tensor-level wrappers over the C ABI (include/fmc_b200.h).  torch is used for device memory and the current
stream only; every function launches hand-written sm_100a kernels through ctypes and raises if the library or a
CUDA device is missing (no CPU / eager fallback)."""
import torch

from . import _cabi

BF16 = torch.bfloat16
F32 = torch.float32
# Activation dtypes: bf16 = the production layout; fp32 = the reference-precision mode (csrc/precise.cu: tf32 tensor-core
# linears, fp32 everything else).  Every wrapper picks the entry point from the dtype of its activation operand and
# refuses mixed operands -- one backend, two precisions.


def _stream():
    return 0 if DRY_RUN else torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


# Host-logic tests only (tests/test_dry_run.py): with DRY_RUN set and _cabi.call replaced by a recorder the wrappers run on
# CPU tensors WITHOUT executing any kernel -- outputs are uninitialised memory.  Never set outside tests.
DRY_RUN = False


def require_cuda(t):
    if not (t.is_cuda or DRY_RUN):
        raise RuntimeError("synfmc_b200 runs on CUDA tensors only (no CPU fallback)")


def _check_cuda(*tensors):
    """Operands must live on the CURRENT CUDA device: kernels are launched on its current stream."""
    if DRY_RUN:
        return
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _cabi.FmcError("synfmc_b200 ops run on CUDA tensors only (no CPU fallback)")
        if t.device.index != torch.cuda.current_device():
            raise _cabi.FmcError(f"operand on {t.device} but the current device is cuda:{torch.cuda.current_device()}: "
                                 "wrap the call in torch.cuda.device(tensor.device)")


def _rows2d(t, dtype=BF16):
    assert t.dtype == dtype and t.ndim == 2 and t.stride(1) == 1, (t.dtype, t.shape, t.stride())
    return t


def _act(t):
    """dtype of an activation operand: bf16 or fp32 (anything else is a caller error)."""
    assert t.dtype in (BF16, F32), t.dtype
    return t.dtype


def split_tf32(x):
    """fp32 [rows, K] -> [rows, 2K] = [tf32(x) | x - tf32(x)]: the operand form of the three-pass tf32 GEMM."""
    _check_cuda(x)
    _rows2d(x, F32)
    rows, K = x.shape
    out = torch.empty((rows, 2 * K), device=x.device, dtype=F32)
    _cabi.call("fmc_split_tf32", x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, K, _stream())
    return out


def gemm_f32(a, w, bias=None, residual=None, out=None, geglu=False, rowbias=None, rows_per_group=0, split=1):
    """Reference-precision linear: fp32 a [M, K]; w fp32 [N, K] (split = 1) or [N, 2K] = [hi | lo] (split = 3, the
    activation is split here)."""
    _check_cuda(a, w)
    _rows2d(a, F32)
    _rows2d(w, F32)
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K * (2 if split == 3 else 1), (w.shape, K, split)
    n_out = N // 2 if geglu else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=F32)
    assert out.dtype == F32 and out.shape == (M, n_out) and out.stride(1) == 1
    if residual is not None:
        _rows2d(residual, F32)
        assert residual.shape == (M, n_out)
    if bias is not None:
        assert bias.dtype == F32 and bias.numel() == N and bias.is_contiguous()
    if rowbias is not None:
        assert rowbias.dtype == F32 and rowbias.stride(1) == 1 and rowbias.shape[1] == N
    a_op = split_tf32(a) if split == 3 else a
    _cabi.call("fmc_gemm_tf32", a_op.data_ptr(), a_op.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0),
               M, N, K, _ptr(bias), _ptr(residual), residual.stride(0) if residual is not None else 0, _ptr(rowbias),
               rows_per_group, rowbias.stride(0) if rowbias is not None else 0, 1 if geglu else 0, split, _stream())
    return out


def attention_f32(q, q_col0, k, k_col0, v, v_col0, out, images, heads, head_dim, nq, nk, kv_div, kv_stride, inner, scale):
    _check_cuda(q, k, v, out)
    for t in (q, k, v, out):
        _rows2d(t, F32)
    _cabi.call("fmc_attention_f32", q.data_ptr(), q.stride(0), q_col0, k.data_ptr(), k.stride(0), k_col0, v.data_ptr(),
               v.stride(0), v_col0, out.data_ptr(), out.stride(0), images, heads, head_dim, nq, nk, kv_div, kv_stride,
               inner, float(scale), _stream())
    return out


def im2col3x3(x, stride=1):
    """fp32 [N, H, W, C] -> [N * OH * OW, 9 * C] rows in (ky, kx, c) order (3x3, padding 1)."""
    _check_cuda(x)
    assert x.dtype == F32 and x.is_contiguous()
    N, H, W, C = x.shape
    OH, OW = (H - 1) // stride + 1, (W - 1) // stride + 1
    out = torch.empty((N * OH * OW, 9 * C), device=x.device, dtype=F32)
    _cabi.call("fmc_im2col3x3_f32", x.data_ptr(), out.data_ptr(), N, H, W, C, stride, _stream())
    return out, OH, OW


def rowstats(x, eps=1e-5):
    """(mean, rstd) per row of a bf16 [rows, C] tensor -> fp32 [rows, 2] (fmc_rowstats_bf16)."""
    _check_cuda(x)
    _rows2d(x)
    stats = torch.empty((x.shape[0], 2), device=x.device, dtype=torch.float32)
    _cabi.call("fmc_rowstats_bf16", x.data_ptr(), x.stride(0), stats.data_ptr(), x.shape[0], x.shape[1], float(eps),
               _stream())
    return stats


def gemm(a, w, bias=None, residual=None, out=None, geglu=False, out_f32=False, rowbias=None, rows_per_group=0,
         tile_n=0, f16_from_col=None, ln_stats=None, ln_colsum=None):
    """out[M, N(/2)] = epilogue(a[M, K] @ w[N, K]^T); see fmc_gemm_bf16.  `f16_from_col`: columns from there on are
    written as IEEE fp16 bit patterns into the bf16 output tensor (FMC_GEMM_F16_TAIL)."""
    _check_cuda(a, w)
    _rows2d(a)
    _rows2d(w)
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    n_out = N // 2 if geglu else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=torch.float32 if out_f32 else BF16)
    assert out.shape == (M, n_out) and out.stride(1) == 1
    flags = (1 if geglu else 0) | (2 if out_f32 else 0)
    if f16_from_col is not None:
        assert f16_from_col % 32 == 0 and not geglu and not out_f32 and residual is None
        flags |= 4 | ((f16_from_col // 32) << 8)
    if residual is not None:
        _rows2d(residual)
        assert residual.shape == (M, n_out)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
    if rowbias is not None:
        assert rowbias.dtype == torch.float32 and rowbias.stride(1) == 1 and rowbias.shape[1] == N
    if ln_stats is not None:
        # LayerNorm folded into the GEMM: `a` un-normalised, `w` scaled by gamma, bias carrying W beta
        assert residual is None and rowbias is None and not out_f32
        assert ln_stats.dtype == torch.float32 and ln_stats.shape == (M, 2) and ln_stats.is_contiguous()
        assert ln_colsum.dtype == torch.float32 and ln_colsum.numel() == N and ln_colsum.is_contiguous()
        _cabi.call("fmc_gemm_ln_bf16", a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0),
                   M, N, K, _ptr(bias), ln_colsum.data_ptr(), ln_stats.data_ptr(), flags, tile_n, _stream())
        return out
    _cabi.call("fmc_gemm_bf16", a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0), M, N,
               K, _ptr(bias), _ptr(residual), residual.stride(0) if residual is not None else 0, _ptr(rowbias),
               rows_per_group, rowbias.stride(0) if rowbias is not None else 0, flags, tile_n, _stream())
    return out


def spatial_attn(q, q_col0, k, k_col0, v, v_col0, head_stride, out, images, heads, head_dim, nq, nk, kv_div, kv_stride,
                 scale, v_f16=False, lse=None):
    """`v_f16`: the V columns hold IEEE fp16 bit patterns (gemm(..., f16_from_col=v_col0)); head_dim 40 only.
    `lse`: fp32 [q rows, heads] receiving the row log-sum-exp in log2 units (training forward: the backward skips its
    first sweep, fmc_attention_bwd_bf16 with lse_given)."""
    if _act(q) == F32:
        assert not v_f16 and head_stride == head_dim and lse is None
        return attention_f32(q, q_col0, k, k_col0, v, v_col0, out, images, heads, head_dim, nq, nk, kv_div, kv_stride, 1,
                             scale)
    _check_cuda(q, k, v, out)
    for t in (q, k, v, out):
        _rows2d(t)
    assert k.shape[0] == v.shape[0]
    if lse is not None:
        assert not v_f16 and lse.dtype == F32 and lse.shape == (q.shape[0], heads) and lse.is_contiguous()
        _cabi.call("fmc_spatial_attn_lse_bf16", q.data_ptr(), q.stride(0), q_col0, q.shape[0], k.data_ptr(), k.stride(0), k_col0,
                   v.data_ptr(), v.stride(0), v_col0, k.shape[0], head_stride, out.data_ptr(), out.stride(0), lse.data_ptr(),
                   images, heads, head_dim, nq, nk, kv_div, kv_stride, float(scale), _stream())
        return out
    _cabi.call("fmc_spatial_attn_vf16" if v_f16 else "fmc_spatial_attn_bf16", q.data_ptr(), q.stride(0), q_col0, q.shape[0], k.data_ptr(), k.stride(0), k_col0,
               v.data_ptr(), v.stride(0), v_col0, k.shape[0], head_stride, out.data_ptr(), out.stride(0), images, heads,
               head_dim, nq, nk, kv_div, kv_stride, float(scale), _stream())
    return out


def temporal_attn(qkv, q_col0, k_col0, v_col0, head_stride, out, B, F, HW, heads, head_dim, scale):
    if _act(qkv) == F32:
        assert head_stride == head_dim and qkv.shape[0] == B * F * HW == out.shape[0]
        return attention_f32(qkv, q_col0, qkv, k_col0, qkv, v_col0, out, B * HW, heads, head_dim, F, F, 1, F, HW, scale)
    _check_cuda(qkv, out)
    _rows2d(qkv)
    _rows2d(out)
    assert qkv.shape[0] == B * F * HW == out.shape[0]
    _cabi.call("fmc_temporal_attn_bf16", qkv.data_ptr(), qkv.stride(0), q_col0, k_col0, v_col0, head_stride,
               out.data_ptr(), out.stride(0), B, F, HW, heads, head_dim, float(scale), _stream())
    return out


def temporal_qkv_attn(x, w_qkv, out, B, F, HW, heads, scale):
    """Fused per-head q|k|v projection + attention over frames (fmc_temporal_qkv_attn_bf16); x, out: [(B F HW), 320]."""
    _check_cuda(x, w_qkv, out)
    _rows2d(x)
    _rows2d(w_qkv)
    _rows2d(out)
    assert x.shape[0] == B * F * HW == out.shape[0] and x.shape[1] == out.shape[1] == w_qkv.shape[1]
    _cabi.call("fmc_temporal_qkv_attn_bf16", x.data_ptr(), x.stride(0), w_qkv.data_ptr(), w_qkv.stride(0), out.data_ptr(),
               out.stride(0), B, F, HW, x.shape[1], heads, float(scale), _stream())
    return out


def layernorm(x, gamma, beta, eps=1e-5, out=None, pe=None, F=0, HW=0, add=None, out2=None):
    _check_cuda(x)
    dt = _act(x)
    _rows2d(x, dt)
    rows, C = x.shape
    if out is None:
        out = torch.empty((rows, C), device=x.device, dtype=dt)
    if add is not None and out2 is None:
        out2 = torch.empty((rows, C), device=x.device, dtype=dt)
    if add is not None:
        _rows2d(add, dt)
    _cabi.call("fmc_layernorm_f32" if dt == F32 else "fmc_layernorm_bf16", x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(), float(eps),
               out.data_ptr(), out.stride(0), _ptr(pe), F, HW, _ptr(add), add.stride(0) if add is not None else 0,
               _ptr(out2), out2.stride(0) if out2 is not None else 0, rows, C, _stream())
    return (out, out2) if add is not None else out


def groupnorm(x, gamma, beta, eps, images, HW, groups=32, silu=False, rowbias=None, rowbias_div=1, out=None):
    """x: [images*HW, C] rows (channels-last); statistics per (image, group)."""
    _check_cuda(x)
    dt = _act(x)
    _rows2d(x, dt)
    rows, C = x.shape
    assert rows == images * HW
    if out is None:
        out = torch.empty((rows, C), device=x.device, dtype=dt)
    stats = torch.empty((2 * images * (groups * ((HW + 63) // 64) + C),), device=x.device, dtype=torch.float32)
    if rowbias is not None:
        assert rowbias.dtype == torch.float32 and rowbias.shape == (images // rowbias_div, C) and rowbias.stride(1) == 1
    _cabi.call("fmc_groupnorm_f32" if dt == F32 else "fmc_groupnorm_bf16", x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(), float(eps),
               out.data_ptr(), out.stride(0), stats.data_ptr(), images, HW, C, groups, 1 if silu else 0, _ptr(rowbias),
               rowbias.stride(0) if rowbias is not None else 0, rowbias_div, _stream())
    return out


def add(a, b=None, rowbias=None, rows_per_group=0, relu=False, out=None):
    _check_cuda(a)
    dt = _act(a)
    _rows2d(a, dt)
    rows, C = a.shape
    if out is None:
        out = torch.empty((rows, C), device=a.device, dtype=dt)
    if b is not None:
        _rows2d(b, dt)
        assert b.shape == a.shape
    _cabi.call("fmc_add_f32" if dt == F32 else "fmc_add_bf16", a.data_ptr(), a.stride(0), _ptr(b), b.stride(0) if b is not None else 0, _ptr(rowbias),
               rows_per_group, rowbias.stride(0) if rowbias is not None else 0, out.data_ptr(), out.stride(0), rows, C,
               1 if relu else 0, _stream())
    return out


def resize_nearest(x, oh, ow):
    """x: [N, h, w, C] contiguous bf16 / fp32."""
    _check_cuda(x)
    dt = _act(x)
    assert x.is_contiguous()
    N, h, w, C = x.shape
    out = torch.empty((N, oh, ow, C), device=x.device, dtype=dt)
    _cabi.call("fmc_resize_nearest_f32" if dt == F32 else "fmc_resize_nearest_bf16", x.data_ptr(), out.data_ptr(), N, h, w, oh, ow, C, _stream())
    return out


def avgpool2(x):
    _check_cuda(x)
    dt = _act(x)
    assert x.is_contiguous()
    N, h, w, C = x.shape
    out = torch.empty((N, h // 2, w // 2, C), device=x.device, dtype=dt)
    _cabi.call("fmc_avgpool2_f32" if dt == F32 else "fmc_avgpool2_bf16", x.data_ptr(), out.data_ptr(), N, h, w, C, _stream())
    return out


def copy2d(src, dst):
    """dst[:, :cols] = src (both row-strided 2-D bf16 views)."""
    _check_cuda(src, dst)
    dt = _act(src)
    _rows2d(src, dt)
    _rows2d(dst, dt)
    assert src.shape == dst.shape
    _cabi.call("fmc_copy2d_f32" if dt == F32 else "fmc_copy2d_bf16", src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0), src.shape[0],
               src.shape[1], _stream())
    return dst


def to_channels_last(x, c_pad=None, dtype=BF16):
    """[B, C, F, H, W] fp32 -> [B, F, H, W, Cpad] bf16 (or fp32 in the reference-precision mode)."""
    _check_cuda(x)
    x = x.contiguous().float()
    B, C, F, H, W = x.shape
    c_pad = c_pad or C
    out = torch.empty((B, F, H, W, c_pad), device=x.device, dtype=dtype)
    _cabi.call("fmc_ncfhw_f32_to_cl_f32" if dtype == F32 else "fmc_ncfhw_f32_to_cl_bf16", x.data_ptr(), out.data_ptr(), B, C, F, H * W, c_pad, _stream())
    return out


def from_channels_last(x, C=None):
    """[B, F, H, W, ld] bf16 / fp32 -> [B, C, F, H, W] fp32."""
    _check_cuda(x)
    dt = _act(x)
    assert x.is_contiguous()
    B, F, H, W, ld = x.shape
    C = C or ld
    out = torch.empty((B, C, F, H, W), device=x.device, dtype=torch.float32)
    _cabi.call("fmc_cl_f32_to_ncfhw_f32" if dt == F32 else "fmc_cl_bf16_to_ncfhw_f32", x.data_ptr(), ld, out.data_ptr(), B, C, F, H * W, _stream())
    return out


def cast_act(x, silu=False, dtype=BF16):
    _check_cuda(x)
    x = x.contiguous()
    assert x.dtype == torch.float32
    if dtype == F32:
        if not silu:
            return x
        out = torch.empty_like(x)
        _cabi.call("fmc_silu_f32", x.data_ptr(), out.data_ptr(), x.numel(), _stream())
        return out
    out = torch.empty(x.shape, device=x.device, dtype=BF16)
    _cabi.call("fmc_cast_act_bf16", x.data_ptr(), out.data_ptr(), x.numel(), 1 if silu else 0, _stream())
    return out


def timestep_embedding(t, dim, dtype=BF16):
    _check_cuda(t)
    t = t.contiguous().float()
    out = torch.empty((t.numel(), dim), device=t.device, dtype=dtype)
    _cabi.call("fmc_timestep_embedding_f32" if dtype == F32 else "fmc_timestep_embedding_bf16", t.data_ptr(), out.data_ptr(), t.numel(), dim, _stream())
    return out


def plucker(K, c2w, H, W):
    """K [BF, 4], c2w [BF, 3, 4] fp32 -> [BF, H, W, 6] fp32."""
    _check_cuda(K, c2w)
    K = K.contiguous().float()
    c2w = c2w.contiguous().float()
    BF = K.shape[0]
    out = torch.empty((BF, H, W, 6), device=K.device, dtype=torch.float32)
    _cabi.call("fmc_plucker_f32", K.data_ptr(), c2w.data_ptr(), out.data_ptr(), BF, H, W, _stream())
    return out


def plucker_unshuffle(K, c2w, H, W):
    """-> [BF, H/8, W/8, 384] bf16 (PixelUnshuffle(8) fused)."""
    _check_cuda(K, c2w)
    K = K.contiguous().float()
    c2w = c2w.contiguous().float()
    BF = K.shape[0]
    out = torch.empty((BF, H // 8, W // 8, 384), device=K.device, dtype=BF16)
    _cabi.call("fmc_plucker_unshuffle_bf16", K.data_ptr(), c2w.data_ptr(), out.data_ptr(), BF, H, W, _stream())
    return out


def traj_scatter(info, masks):
    """info [BF, n_obj, 12], masks [BF, n_obj, H, W] fp32 -> (feat [BF, 13, H, W] fp32, mask [BF, H, W] fp32)."""
    _check_cuda(info, masks)
    info = info.contiguous().float()
    masks = masks.contiguous().float()
    BF, n_obj, H, W = masks.shape
    feat = torch.empty((BF, 13, H, W), device=masks.device, dtype=torch.float32)
    mask = torch.empty((BF, H, W), device=masks.device, dtype=torch.float32)
    _cabi.call("fmc_traj_scatter_f32", info.data_ptr(), masks.data_ptr(), feat.data_ptr(), mask.data_ptr(), BF, n_obj, H, W,
               _stream())
    return feat, mask


def traj_scatter_unshuffle(info, masks):
    """-> (feat [BF, H/8, W/8, 832] bf16, mask [BF, H, W] fp32)."""
    _check_cuda(info, masks)
    info = info.contiguous().float()
    masks = masks.contiguous().float()
    BF, n_obj, H, W = masks.shape
    feat = torch.empty((BF, H // 8, W // 8, 832), device=masks.device, dtype=BF16)
    mask = torch.empty((BF, H, W), device=masks.device, dtype=torch.float32)
    _cabi.call("fmc_traj_scatter_unshuffle_bf16", info.data_ptr(), masks.data_ptr(), feat.data_ptr(), mask.data_ptr(), BF,
               n_obj, H, W, _stream())
    return feat, mask


def sphere_masks(circles, H, W):
    """circles [BF, n_obj, 3] = (cx, cy, r) fp32 -> Gaussian sphere masks [BF, n_obj, H, W] fp32 (fmc_sphere_mask_f32)."""
    _check_cuda(circles)
    circles = circles.contiguous().float()
    BF, n_obj, _ = circles.shape
    out = torch.empty((BF, n_obj, H, W), device=circles.device, dtype=torch.float32)
    _cabi.call("fmc_sphere_mask_f32", circles.data_ptr(), out.data_ptr(), BF, n_obj, H, W, _stream())
    return out


def traj_scatter_circles_unshuffle(info, circles, H, W):
    """info [BF, n_obj, 12], circles [BF, n_obj, 3] -> (feat [BF, H/8, W/8, 832] bf16, mask [BF, H, W] fp32): the object
    scatter with the Gaussian sphere masks generated on the fly."""
    _check_cuda(info, circles)
    info = info.contiguous().float()
    circles = circles.contiguous().float()
    BF, n_obj, _ = circles.shape
    assert info.shape == (BF, n_obj, 12)
    feat = torch.empty((BF, H // 8, W // 8, 832), device=info.device, dtype=BF16)
    mask = torch.empty((BF, H, W), device=info.device, dtype=torch.float32)
    _cabi.call("fmc_traj_scatter_circles_unshuffle_bf16", info.data_ptr(), circles.data_ptr(), feat.data_ptr(),
               mask.data_ptr(), BF, n_obj, H, W, _stream())
    return feat, mask


def mask_modulate(x, mask, row_index, col_index):
    """x [N, h, w, C] bf16 / fp32, mask [N, H, W] fp32, index maps int32 [h], [w]."""
    _check_cuda(x, mask)
    dt = _act(x)
    assert x.is_contiguous() and mask.is_contiguous()
    N, h, w, C = x.shape
    out = torch.empty_like(x)
    _cabi.call("fmc_mask_modulate_f32" if dt == F32 else "fmc_mask_modulate_bf16", x.data_ptr(), mask.data_ptr(), row_index.data_ptr(), col_index.data_ptr(),
               out.data_ptr(), N, h, w, C, mask.shape[1], mask.shape[2], _stream())
    return out


def cfg_ddim_step(eps_uncond, eps_cond, guidance_scale, latents, alpha_t, alpha_prev, return_eps=False):
    _check_cuda(eps_uncond, latents)
    assert eps_uncond.dtype == torch.float32 and latents.dtype == torch.float32
    eps_uncond = eps_uncond.contiguous()
    latents = latents.contiguous()
    if eps_cond is not None:
        eps_cond = eps_cond.contiguous()
    out = torch.empty_like(latents)
    eps_out = torch.empty_like(latents) if return_eps else None
    _cabi.call("fmc_cfg_ddim_step_f32", eps_uncond.data_ptr(), _ptr(eps_cond), float(guidance_scale), latents.data_ptr(),
               out.data_ptr(), _ptr(eps_out), float(alpha_t), float(alpha_prev), latents.numel(), _stream())
    return (out, eps_out) if return_eps else out


def window_combine_ddim(eps_windows, cfg, guidance_scale, latents, L, stride, alpha_t, alpha_prev):
    """eps_windows [n_win, (2)b, C, L, h, w] fp32, latents [b, C, F_total, h, w] fp32 -> new latents (window-averaged
    guided prediction + DDIM update, fmc_window_combine_ddim_f32)."""
    _check_cuda(eps_windows, latents)
    assert eps_windows.dtype == torch.float32 and latents.dtype == torch.float32
    assert eps_windows.is_contiguous() and latents.is_contiguous()
    n_win = eps_windows.shape[0]
    b, C, F_total, h, w = latents.shape
    assert eps_windows.shape[1:] == ((2 if cfg else 1) * b, C, L, h, w), (eps_windows.shape, latents.shape)
    out = torch.empty_like(latents)
    _cabi.call("fmc_window_combine_ddim_f32", eps_windows.data_ptr(), n_win, 1 if cfg else 0, float(guidance_scale),
               latents.data_ptr(), out.data_ptr(), b, C, F_total, h * w, L, stride, float(alpha_t), float(alpha_prev),
               _stream())
    return out


def conv3x3_supported(H, W, cin, cout, stride):
    """Geometry handled by fmc_conv3x3_bf16 (everything the FMC U-Net / encoders use at the BASELINE shapes)."""
    if stride not in (1, 2) or H % stride or W % stride or cin % 64 or cout % 32:
        return False
    ow = W // stride
    return 4 <= ow <= 128 and 128 % ow == 0


def conv3x3(x, w2d, bias=None, residual=None, stride=1, tile_n=0):
    """x [N, H, W, Cin] bf16 contiguous, w2d [Cout, 9 * Cin] bf16 (ky, kx, cin order) -> [N, H/s, W/s, Cout]."""
    _check_cuda(x, w2d)
    assert x.dtype == BF16 and x.is_contiguous() and w2d.dtype == BF16 and w2d.is_contiguous()
    N, H, W, cin = x.shape
    cout = w2d.shape[0]
    assert w2d.shape[1] == 9 * cin
    out = torch.empty((N, H // stride, W // stride, cout), device=x.device, dtype=BF16)
    if residual is not None:
        assert residual.dtype == BF16 and residual.is_contiguous() and residual.shape == out.shape
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == cout and bias.is_contiguous()
    _cabi.call("fmc_conv3x3_bf16", x.data_ptr(), w2d.data_ptr(), out.data_ptr(), _ptr(bias), _ptr(residual), N, H, W, cin,
               cout, stride, tile_n, _stream())
    return out


def conv2d_cl(x, weight, bias, stride=1, padding=1):
    """3x3 / strided convolutions on channels-last bf16 [N, h, w, Cin] -> [N, oh, ow, Cout].

    Round 1: cuDNN through torch (library call, like cuBLAS); the implicit-GEMM tcgen05 conv is SURVEY 8(f) row 1.
    `weight` must already be bf16 in torch.channels_last memory format."""
    _check_cuda(x)
    if _cabi.trace is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), weight, bias, stride=stride, padding=padding)
    if _cabi.trace is not None:
        e1.record()
        _cabi.trace.append(("cudnn_conv2d", (x.shape[0] * y.shape[2] * y.shape[3], weight.shape[0],
                                             weight.shape[1] * weight.shape[2] * weight.shape[3]), e0, e1))
    y = y.permute(0, 2, 3, 1)
    return y if y.is_contiguous() else y.contiguous()


# --------------------------------------------------------------------------------------------------------------
# pipeline edges (csrc/edge.cu): VAE + CLIP text encoder pieces
# --------------------------------------------------------------------------------------------------------------
def softmax_rows(scores, scale, out_dtype=BF16):
    """softmax(scale * scores) per row; scores fp32 [rows, n] (n % 4 == 0, n <= 4096) -> bf16 / fp32 probabilities."""
    _check_cuda(scores)
    _rows2d(scores, F32)
    rows, n = scores.shape
    out = torch.empty((rows, n), device=scores.device, dtype=out_dtype)
    _cabi.call("fmc_softmax_rows", scores.data_ptr(), scores.stride(0), out.data_ptr(), out.stride(0),
               1 if out_dtype == F32 else 0, rows, n, float(scale), _stream())
    return out


def small_mha(qkv, q_col0, k_col0, v_col0, seqs, tokens_per_seq, heads, head_dim, scale, causal):
    """Multi-head self-attention on the fused projection rows [seqs * tokens_per_seq, >= 3 * heads * head_dim]."""
    _check_cuda(qkv)
    dt = _act(qkv)
    _rows2d(qkv, dt)
    assert qkv.shape[0] == seqs * tokens_per_seq
    out = torch.empty((qkv.shape[0], heads * head_dim), device=qkv.device, dtype=dt)
    _cabi.call("fmc_small_mha", qkv.data_ptr(), qkv.stride(0), q_col0, k_col0, v_col0, out.data_ptr(), out.stride(0),
               1 if dt == F32 else 0, seqs, tokens_per_seq, heads, head_dim, float(scale), 1 if causal else 0, _stream())
    return out


def quick_gelu(x, out=None):
    _check_cuda(x)
    dt = _act(x)
    _rows2d(x, dt)
    if out is None:
        out = torch.empty_like(x)
    _cabi.call("fmc_quick_gelu", x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), 1 if dt == F32 else 0, x.shape[0],
               x.shape[1], _stream())
    return out


def embed_tokens(ids, token_table, position_table, out_dtype=BF16):
    """ids int64 [B, T]; fp32 tables [vocab, C], [>= T, C] -> rows [B * T, C]."""
    _check_cuda(ids, token_table, position_table)
    assert ids.dtype == torch.int64 and ids.ndim == 2 and ids.is_contiguous()
    _rows2d(token_table, F32)
    _rows2d(position_table, F32)
    assert token_table.is_contiguous() and position_table.is_contiguous()
    B, T = ids.shape
    vocab, C = token_table.shape
    assert position_table.shape[0] >= T and position_table.shape[1] == C
    out = torch.empty((B * T, C), device=ids.device, dtype=out_dtype)
    _cabi.call("fmc_embed_tokens", ids.data_ptr(), token_table.data_ptr(), position_table.data_ptr(), out.data_ptr(),
               out.stride(0), 1 if out_dtype == F32 else 0, B * T, T, C, vocab, _stream())
    return out


def vae_sample(moments, N, z, HW, noise=None, out_scale=1.0):
    """moments rows [N * HW, 2 z] (channels-last, mean | logvar) -> [N, z, HW] fp32; noise [N, z, HW] fp32 or None (mode)."""
    _check_cuda(moments, noise)
    dt = _act(moments)
    _rows2d(moments, dt)
    assert moments.shape == (N * HW, 2 * z)
    if noise is not None:
        assert noise.dtype == F32 and noise.is_contiguous() and noise.numel() == N * z * HW
    out = torch.empty((N, z, HW), device=moments.device, dtype=F32)
    _cabi.call("fmc_vae_sample_f32", moments.data_ptr(), moments.stride(0), 1 if dt == F32 else 0, _ptr(noise), out.data_ptr(),
               N, z, HW, float(out_scale), _stream())
    return out


def cl_to_video(x, B, C, F, HW, mul=1.0, add=0.0, lo=-3.0e38, hi=3.0e38):
    """channels-last rows [(B F) HW, >= C] -> [B, C, F, HW] fp32 = clamp(x * mul + add, lo, hi)."""
    _check_cuda(x)
    dt = _act(x)
    _rows2d(x, dt)
    assert x.shape[0] == B * F * HW and x.shape[1] >= C
    out = torch.empty((B, C, F, HW), device=x.device, dtype=torch.float32)
    _cabi.call("fmc_cl_to_video_f32", x.data_ptr(), x.stride(0), 1 if dt == F32 else 0, out.data_ptr(), B, C, F, HW, float(mul),
               float(add), float(lo), float(hi), _stream())
    return out
