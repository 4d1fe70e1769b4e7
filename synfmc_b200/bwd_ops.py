"""Tensor-level wrappers of the backward entry points (include/fmc_b200.h, csrc/backward.cu).  Same conventions as
ops.py: bf16 channels-last rows, CUDA tensors on the current device, no fallback."""
import torch

from . import _cabi, ops
from .ops import BF16, _check_cuda, _ptr, _rows2d, _stream

F32 = torch.float32


def transpose(x, pad_rows_to=1):
    """bf16 [rows, cols] (row-strided) -> contiguous [cols, rows]; with pad_rows_to = 8 the result is [cols, ceil8(rows)]
    with zero columns at the end (the K extent of a following GEMM must be a multiple of 8)."""
    _check_cuda(x)
    _rows2d(x)
    rows, cols = x.shape
    padded = (rows + pad_rows_to - 1) // pad_rows_to * pad_rows_to
    out = (torch.empty if padded == rows else torch.zeros)((cols, padded), device=x.device, dtype=BF16)
    _cabi.call("fmc_transpose_bf16", x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, cols, _stream())
    return out


def colsum(x, out=None, accumulate=False):
    """fp32 column sums of a bf16 / fp32 [rows, cols] tensor (bias gradients)."""
    _check_cuda(x)
    assert x.ndim == 2 and x.stride(1) == 1 and x.dtype in (BF16, F32)
    rows, cols = x.shape
    if out is None:
        out = torch.zeros(cols, device=x.device, dtype=F32)
        accumulate = False
    assert out.dtype == F32 and out.numel() == cols and out.is_contiguous()
    ws = torch.empty(max(1, _cabi.lib().fmc_colsum_workspace_floats(rows, cols)), device=x.device, dtype=F32)
    _cabi.call("fmc_colsum_f32", x.data_ptr(), x.stride(0), 1 if x.dtype == BF16 else 0, out.data_ptr(), ws.data_ptr(), rows,
               cols, 1 if accumulate else 0, _stream())
    return out


def layernorm_bwd(x, dy, gamma, eps=1e-5, want_params=False):
    """-> dx (bf16) [, d gamma, d beta (fp32)]."""
    _check_cuda(x, dy)
    _rows2d(x)
    _rows2d(dy)
    rows, C = x.shape
    dx = torch.empty((rows, C), device=x.device, dtype=BF16)
    part = None
    if want_params:
        nb = _cabi.lib().fmc_layernorm_bwd_blocks(rows)
        part = torch.empty((nb, 2 * C), device=x.device, dtype=F32)
    _cabi.call("fmc_layernorm_bwd_bf16", x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), gamma.data_ptr(), float(eps),
               dx.data_ptr(), dx.stride(0), _ptr(part), rows, C, _stream())
    if not want_params:
        return dx
    both = colsum(part)
    return dx, both[:C], both[C:]


def _gn_bwd_ws_floats(images, HW, groups):
    if ops.DRY_RUN:
        return 4 * images * ((HW + 31) // 32) * groups
    return int(_cabi.lib().fmc_groupnorm_bwd_workspace_floats(images, HW, groups))


def groupnorm_bwd(x, dy, gamma, beta, eps, images, HW, groups=32, silu=False, rowbias=None, rowbias_div=1):
    _check_cuda(x, dy)
    _rows2d(x)
    _rows2d(dy)
    rows, C = x.shape
    assert rows == images * HW and dy.shape == x.shape
    dx = torch.empty((rows, C), device=x.device, dtype=BF16)
    ws = torch.empty(_gn_bwd_ws_floats(images, HW, groups), device=x.device, dtype=F32)
    _cabi.call("fmc_groupnorm_bwd_bf16", x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), gamma.data_ptr(),
               beta.data_ptr(), float(eps), dx.data_ptr(), dx.stride(0), ws.data_ptr(), images, HW, C, groups,
               1 if silu else 0, _ptr(rowbias), rowbias.stride(0) if rowbias is not None else 0, rowbias_div, _stream())
    return dx


def geglu_fwd(proj):
    """proj bf16 [rows, 2H] in the interleaved (16 value | 16 gate) layout -> y [rows, H]."""
    _check_cuda(proj)
    _rows2d(proj)
    rows, H2 = proj.shape
    y = torch.empty((rows, H2 // 2), device=proj.device, dtype=BF16)
    _cabi.call("fmc_geglu_fwd_bf16", proj.data_ptr(), proj.stride(0), y.data_ptr(), y.stride(0), rows, H2 // 2, _stream())
    return y


def geglu_bwd(proj, dy):
    _check_cuda(proj, dy)
    _rows2d(proj)
    _rows2d(dy)
    rows, H2 = proj.shape
    assert dy.shape == (rows, H2 // 2)
    dproj = torch.empty((rows, H2), device=proj.device, dtype=BF16)
    _cabi.call("fmc_geglu_bwd_bf16", proj.data_ptr(), proj.stride(0), dy.data_ptr(), dy.stride(0), dproj.data_ptr(),
               dproj.stride(0), rows, H2 // 2, _stream())
    return dproj


def relu_bwd(y, dy):
    _check_cuda(y, dy)
    assert y.dtype == BF16 and dy.dtype == BF16 and y.is_contiguous() and dy.is_contiguous() and y.shape == dy.shape
    dx = torch.empty_like(dy)
    _cabi.call("fmc_relu_bwd_bf16", y.data_ptr(), dy.data_ptr(), dx.data_ptr(), y.numel(), _stream())
    return dx


def resize_nearest_bwd(dy, h, w):
    """dy [N, oh, ow, C] -> dx [N, h, w, C] (integer factors)."""
    _check_cuda(dy)
    assert dy.dtype == BF16 and dy.is_contiguous()
    N, oh, ow, C = dy.shape
    dx = torch.empty((N, h, w, C), device=dy.device, dtype=BF16)
    _cabi.call("fmc_resize_nearest_bwd_bf16", dy.data_ptr(), dx.data_ptr(), N, h, w, oh, ow, C, _stream())
    return dx


def avgpool2_bwd(dy, h, w):
    _check_cuda(dy)
    assert dy.dtype == BF16 and dy.is_contiguous()
    N, oh, ow, C = dy.shape
    assert oh == h // 2 and ow == w // 2
    dx = torch.empty((N, h, w, C), device=dy.device, dtype=BF16)
    _cabi.call("fmc_avgpool2_bwd_bf16", dy.data_ptr(), dx.data_ptr(), N, h, w, C, _stream())
    return dx


def attention_bwd(q, q_col0, k, k_col0, v, v_col0, head_stride, o, do, dq, dq_col0, dk, dk_col0, dv, dv_col0, images, heads,
                  head_dim, nq, nk, kv_div, kv_stride, inner, scale, lse=None):
    """Gradients written into dq (and dk, dv unless None) in the layouts of q / k / v.  `lse`: the forward's row
    log-sum-exp (ops.spatial_attn(..., lse=...)), fp32 [q rows, heads]; saves the tcgen05 path its first sweep."""
    _check_cuda(q, k, v, o, do, dq)
    for t in (q, k, v, o, do, dq) + ((dk, dv) if dk is not None else ()):
        _rows2d(t)
    lse_given = lse is not None
    if lse_given:
        assert lse.dtype == F32 and lse.shape == (q.shape[0], heads) and lse.is_contiguous()
    else:
        lse = torch.empty((q.shape[0], heads), device=q.device, dtype=F32)
    dsum = torch.empty((q.shape[0], heads), device=q.device, dtype=F32)
    _cabi.call("fmc_attention_bwd_bf16", q.data_ptr(), q.stride(0), q_col0, k.data_ptr(), k.stride(0), k_col0, v.data_ptr(),
               v.stride(0), v_col0, head_stride, o.data_ptr(), o.stride(0), do.data_ptr(), do.stride(0), dq.data_ptr(),
               dq.stride(0), dq_col0, _ptr(dk), dk.stride(0) if dk is not None else 0, dk_col0, _ptr(dv),
               dv.stride(0) if dv is not None else 0, dv_col0, lse.data_ptr(), dsum.data_ptr(), 1 if lse_given else 0, images, heads,
               head_dim, nq, nk,
               kv_div, kv_stride, inner, float(scale), _stream())
    return dq


def linear_dgrad(dy, w_t, residual=None):
    """dX = dY W for a forward y = x W^T: `w_t` is the transposed weight copy [K_in, N_out] bf16 (a GEMM weight with the
    roles of N and K swapped); `residual` adds another gradient flowing into x (fused in the GEMM epilogue)."""
    return ops.gemm(dy, w_t, residual=residual)


def linear_wgrad(dy, x, out=None, accumulate=False):
    """dW [N_out, K_in] (fp32) (+)= dY^T X on the tensor cores, straight from the row-major dY [T, N_out] and X [T, K_in]
    (fmc_wgrad_bf16: MN-major operands, token axis split across CTAs, deterministic fold)."""
    _check_cuda(dy, x)
    _rows2d(dy)
    _rows2d(x)
    T, M = dy.shape
    N = x.shape[1]
    assert x.shape[0] == T
    if M % 8 or N % 8 or dy.stride(0) % 8 or x.stride(0) % 8:
        # odd widths (not on the training path): K-major GEMM on transposed copies
        dw = ops.gemm(transpose(dy, 8), transpose(x, 8), out_f32=True)
        if out is None:
            return dw
        return out.add_(dw) if accumulate else out.copy_(dw)
    if out is None:
        assert not accumulate
        out = torch.empty((M, N), device=dy.device, dtype=F32)
    assert out.dtype == F32 and out.shape == (M, N) and out.stride(1) == 1
    nws = 4 * M * N if ops.DRY_RUN else int(_cabi.lib().fmc_wgrad_workspace_floats(T, M, N))
    ws = torch.empty(max(nws, 1), device=dy.device, dtype=F32)
    _cabi.call("fmc_wgrad_bf16", dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0),
               ws.data_ptr(), T, M, N, 1 if accumulate else 0, _stream())
    return out
