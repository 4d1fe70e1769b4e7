"""Training forward + backward of the hot path on the libfmc_b200 kernels (train_cam_ctrl.py:586-665,
train_cam_obj_ctrl.py:843-862): `noise_pred = pose_adaptor(noisy_latents, timesteps, text, pose_embedding[, traj])` with
autograd support for the trainable subsets of the two stages --

    CMC  CameraEncoder (all parameters) + the 20 `qkv_merge` layers of the CameraAdapter      (train_cam_ctrl.py:259-284)
    OMC  the ObjectEncoder `Adapter`                                                           (train_cam_obj_ctrl.py:386-391)

-- while the U-Net itself stays frozen: its ops only propagate ACTIVATION gradients (no weight gradients).

Design: a small reverse-mode tape over channels-last bf16 activations.  Every op of the training forward runs the same
forward kernel as inference (un-fused where the backward needs an intermediate: the GEGLU projection and the q|k|v
tensor are kept, LayerNorms are not folded into their GEMMs) and records a closure that turns the gradient of its output
into gradients of its inputs with the kernels of csrc/backward.cu:
    linear      dX = dY W         fmc_gemm_bf16 against a transposed weight copy (tensor cores)
                dW = dY^T X       fmc_gemm_bf16, fp32 output, on fmc_transpose_bf16'd operands; db = fmc_colsum_f32
    norms       fmc_layernorm_bwd_bf16 (+ d gamma / d beta), fmc_groupnorm_bwd_bf16 (bias + SiLU fused as in the forward)
    attention   fmc_attention_bwd_bf16 (probabilities recomputed; spatial self, text cross: dQ only, temporal)
    GEGLU / ReLU / nearest upsample / AvgPool / mask modulate / concat / adds: their elementwise backward kernels
    3x3 convs   cuDNN through torch.nn.grad (the forward's wide 3x3 convolutions are a cuDNN call as well)
The result is handed to torch.autograd through one Function whose inputs are the trainable parameters, so
`loss.backward()`, GradScaler, DDP-style gradient hooks (synfmc_b200.train.GradAllReduce) and optimizers see ordinary
`.grad`s."""
import torch
from torch import nn

from . import bwd_ops, engine, ops
from .engine import BF16, TEXT_PAD

F32 = torch.float32


# --------------------------------------------------------------------------------------------------------------
# tape
# --------------------------------------------------------------------------------------------------------------
class Var:
    """A tensor of the training graph (any layout) with an accumulated gradient."""
    __slots__ = ("t", "grad", "needs_grad", "_owned")

    def __init__(self, t, needs_grad=True):
        self.t, self.grad, self.needs_grad, self._owned = t, None, needs_grad, False

    def accumulate(self, g):
        """The first gradient is kept by reference (it may be shared with other Vars: an add hands the same tensor to both
        inputs); a second one is summed into a buffer this Var owns -- never in place into a shared tensor."""
        if not self.needs_grad or g is None:
            return
        if self.grad is None:
            self.grad, self._owned = g, False
            return
        a, b = self.grad.view(-1, self.grad.shape[-1]), g.reshape(-1, g.shape[-1])
        if self._owned:
            ops.add(a, b, out=a)
        else:
            self.grad, self._owned = ops.add(a, b).view(self.grad.shape), True


# Early gradient delivery (train.GradAllReduce.install_hooks sets it): `sink(param, grad) -> bool`, called from INSIDE the
# tape's backward as soon as the last contribution to a parameter's gradient has been produced; True = the sink has added
# the gradient to `param.grad` itself (and may have launched the bucket's all-reduce), so the autograd bridge returns
# None for that parameter.  Without it every parameter gradient reaches autograd at the END of the tape's backward, and a
# hook-launched all-reduce cannot overlap the backward pass.
EARLY_GRAD_SINK = None


class Tape:
    def __init__(self):
        self.nodes = []        # (outputs, backward closure)
        self.param_grads = {}  # nn.Parameter -> fp32 gradient tensor (None once delivered early)
        self.parts = {}        # nn.Parameter -> contributions seen in this backward

    def record(self, outputs, backward):
        self.nodes.append((outputs, backward))

    def add_param_grad(self, param, g):
        g = g.reshape(param.shape)
        if self.param_grads.get(param) is not None:
            self.param_grads[param].add_(g)
        else:
            self.param_grads[param] = g.contiguous()
        n = self.parts.get(param, 0) + 1
        self.parts[param] = n
        # the number of contributions per backward is learned in the first step (`_fmc_grad_parts`); from the second step
        # on a gradient is handed over the moment it is complete
        if EARLY_GRAD_SINK is not None and (engine.STRUCTURE_EPOCH, n) == getattr(param, "_fmc_grad_parts", None):
            if EARLY_GRAD_SINK(param, self.param_grads[param]):
                self.param_grads[param] = None

    def finish_params(self):
        """After the backward: remember how many contributions each parameter received (early delivery next time)."""
        for param, n in self.parts.items():
            param._fmc_grad_parts = (engine.STRUCTURE_EPOCH, n)  # re-learned after a processor / structure change
        self.parts = {}

    def backward(self):
        for outputs, fn in reversed(self.nodes):
            grads = [o.grad for o in outputs]
            if all(g is None for g in grads):
                continue
            fn(*grads)
            for o in outputs:
                o.grad, o._owned = None, False  # free as we go
        self.nodes = []


def const(t):
    return Var(t, needs_grad=False)


def _trainable(*params):
    return any(p is not None and p.requires_grad for p in params)


# --------------------------------------------------------------------------------------------------------------
# ops on rows [T, C] bf16
# --------------------------------------------------------------------------------------------------------------
class TrainLinear:
    """y = x W^T + b with bf16 weight copies in both orientations.  `w_eff` / `b_eff`: the effective fp32 weight the kernels
    use (folded LoRA, folded scale, padded heads, interleaved GEGLU rows ...).  `sink(tape, dw_eff, db_eff)`: called in
    backward with the fp32 gradients of the EFFECTIVE weight / bias (db_eff None without a bias) and responsible for mapping
    them onto the nn.Parameters; None = frozen layer (activation gradient only)."""

    def __init__(self, w_eff, b_eff, device, sink=None, live=None):
        """`live`: callable returning the current (w_eff, b_eff) -- given for TRAINABLE layers, whose device copies are
        rebuilt at every forward (the optimizer changes the parameters between steps); frozen layers are converted once."""
        self.device, self.sink, self.live = device, sink, live
        self._set(w_eff, b_eff)

    def _set(self, w_eff, b_eff):
        self.w = engine._dev_bf16(w_eff, self.device)                  # [N, K]
        # [K, N]: the "weight" of the dgrad GEMM (trainable layers redo this every step: the own transpose kernel, not a
        # strided torch copy)
        self.w_t = bwd_ops.transpose(self.w) if self.w.is_cuda else self.w.t().contiguous()
        self.b = engine._dev_f32(b_eff, self.device)
        self.N, self.K = self.w.shape

    def __call__(self, tape, x, residual=None):
        if self.live is not None:
            self._set(*self.live())
        y = ops.gemm(x.t, self.w, bias=self.b, residual=residual.t if residual is not None else None)
        out = Var(y)

        def bwd(dy):
            if x.needs_grad:
                x.accumulate(bwd_ops.linear_dgrad(dy, self.w_t))
            if residual is not None:
                residual.accumulate(dy)
            if self.sink is not None:
                self.sink(tape, bwd_ops.linear_wgrad(dy, x.t), bwd_ops.colsum(dy) if self.b is not None else None)
        tape.record([out], bwd)
        return out


def param_sink(weight, bias, scale=1.0, row_order=None):
    """Gradient sink of a plain layer: d(weight) = scale * d(w_eff) (rows un-permuted when `row_order` reordered them)."""
    if not _trainable(weight, bias):
        return None

    def sink(tape, dw, db):
        if row_order is not None:
            inv = torch.empty_like(row_order)
            inv[row_order] = torch.arange(row_order.numel(), device=inv.device)
            dw, db = dw[inv], (db[inv] if db is not None else None)
        if weight is not None and weight.requires_grad:
            tape.add_param_grad(weight, dw * scale if scale != 1.0 else dw)
        if bias is not None and bias.requires_grad and db is not None:
            tape.add_param_grad(bias, db * scale if scale != 1.0 else db)
    return sink


def linear_of(lin, device, geglu=False):
    """nn.Linear / 1x1 nn.Conv2d -> TrainLinear (GEGLU: rows interleaved as the inference plan does)."""
    order = None
    if geglu:
        half = lin.weight.shape[0] // 2
        idx = torch.arange(half).view(-1, 16)
        order = torch.cat([idx, idx + half], dim=1).reshape(-1).to(lin.weight.device)

    def current():
        w = lin.weight.detach().float().reshape(lin.weight.shape[0], -1)
        b = lin.bias.detach().float() if lin.bias is not None else None
        if order is not None:
            w, b = w[order], (b[order] if b is not None else None)
        return w, b
    sink = param_sink(lin.weight, lin.bias, row_order=order.to(device) if order is not None else None)
    return TrainLinear(*current(), device, sink=sink, live=current if sink is not None else None)


def layernorm(tape, x, norm, pe=None, F=0, HW=0, add=None):
    """LN(x) (+ pe[frame]); with `add` (the pose feature) also returns LN(x) + pe + add."""
    g, b = norm.weight.detach().float(), norm.bias.detach().float()
    res = ops.layernorm(x.t, g, b, norm.eps, pe=pe, F=F, HW=HW, add=add.t if add is not None else None)
    out, out2 = (Var(res[0]), Var(res[1])) if add is not None else (Var(res), None)
    wants = _trainable(norm.weight, norm.bias)

    def bwd(d1, d2=None):
        d = d1
        if d2 is not None:
            d = d2 if d1 is None else ops.add(d1, d2)
            add.accumulate(d2)
        if wants:
            dx, dg, db = bwd_ops.layernorm_bwd(x.t, d, g, norm.eps, want_params=True)
            tape.add_param_grad(norm.weight, dg)
            tape.add_param_grad(norm.bias, db)
        else:
            dx = bwd_ops.layernorm_bwd(x.t, d, g, norm.eps)
        x.accumulate(dx)
    tape.record([out] + ([out2] if out2 is not None else []), bwd)
    return (out, out2) if add is not None else out


def groupnorm(tape, x, norm, images, HW, silu=False, rowbias=None, rowbias_div=1):
    g, b = norm.weight.detach().float(), norm.bias.detach().float()
    assert not _trainable(norm.weight, norm.bias), "GroupNorm parameters belong to the frozen U-Net"
    y = ops.groupnorm(x.t, g, b, norm.eps, images, HW, groups=norm.num_groups, silu=silu, rowbias=rowbias,
                      rowbias_div=rowbias_div)
    out = Var(y)

    def bwd(dy):
        x.accumulate(bwd_ops.groupnorm_bwd(x.t, dy, g, b, norm.eps, images, HW, groups=norm.num_groups, silu=silu,
                                           rowbias=rowbias, rowbias_div=rowbias_div))
    tape.record([out], bwd)
    return out


def geglu(tape, proj):
    out = Var(bwd_ops.geglu_fwd(proj.t))
    tape.record([out], lambda dy: proj.accumulate(bwd_ops.geglu_bwd(proj.t, dy)))
    return out


def add(tape, a, b):
    out = Var(ops.add(a.t, b.t))

    def bwd(dy):
        a.accumulate(dy)
        b.accumulate(dy)
    tape.record([out], bwd)
    return out


def add_rowbias(tape, a, b, rowbias, rows_per_group):
    out = Var(ops.add(a.t, b.t, rowbias=rowbias, rows_per_group=rows_per_group))

    def bwd(dy):
        a.accumulate(dy)
        b.accumulate(dy)
    tape.record([out], bwd)
    return out


def relu(tape, x):
    y = ops.add(x.t, relu=True)
    out = Var(y)
    tape.record([out], lambda dy: x.accumulate(bwd_ops.relu_bwd(y, dy.contiguous())))
    return out


def concat(tape, a, b):
    """channel concat of rows [T, Ca], [T, Cb]"""
    Ca, Cb = a.t.shape[1], b.t.shape[1]
    y = torch.empty((a.t.shape[0], Ca + Cb), device=a.t.device, dtype=BF16)
    ops.copy2d(a.t, y[:, :Ca])
    ops.copy2d(b.t, y[:, Ca:])
    out = Var(y)

    def bwd(dy):
        if a.needs_grad:
            a.accumulate(ops.copy2d(dy[:, :Ca], torch.empty((dy.shape[0], Ca), device=dy.device, dtype=BF16)))
        if b.needs_grad:
            b.accumulate(ops.copy2d(dy[:, Ca:], torch.empty((dy.shape[0], Cb), device=dy.device, dtype=BF16)))
    tape.record([out], bwd)
    return out


class TrainConv:
    """3x3 / strided convolution on channels-last images through cuDNN (forward and both gradients), 1x1 as TrainLinear."""

    def __init__(self, conv, device):
        self.conv = conv
        self.k = conv.kernel_size[0]
        self.linear = linear_of(conv, device) if self.k == 1 else None
        self.device = device
        self.trainable = _trainable(conv.weight, conv.bias)
        if self.linear is None:
            self._set()

    def _set(self):
        conv = self.conv
        self.w = conv.weight.detach().to(device=self.device, dtype=BF16).contiguous(memory_format=torch.channels_last)
        self.b = conv.bias.detach().to(device=self.device, dtype=BF16) if conv.bias is not None else None

    def __call__(self, tape, x, shape):
        """x: Var of rows [(N h w), Cin]; shape = (N, h, w) -> (Var rows [(N oh ow), Cout], (N, oh, ow))."""
        N, h, w = shape
        if self.linear is not None:
            return self.linear(tape, x), shape
        if self.trainable:
            self._set()  # the optimizer moved the weights since the last step
        cin = x.t.shape[1]
        xi = x.t.view(N, h, w, cin).permute(0, 3, 1, 2)
        y = torch.nn.functional.conv2d(xi, self.w, self.b, stride=self.conv.stride, padding=self.conv.padding)
        oh, ow = y.shape[2], y.shape[3]
        yr = y.permute(0, 2, 3, 1)
        yr = yr if yr.is_contiguous() else yr.contiguous()
        out = Var(yr.view(-1, y.shape[1]))
        conv = self.conv
        wants = _trainable(conv.weight, conv.bias)

        def bwd(dy):
            dyi = dy.view(N, oh, ow, -1).permute(0, 3, 1, 2)
            if x.needs_grad:
                dx = torch.nn.grad.conv2d_input(xi.shape, self.w, dyi, stride=conv.stride, padding=conv.padding)
                dxr = dx.permute(0, 2, 3, 1)
                x.accumulate((dxr if dxr.is_contiguous() else dxr.contiguous()).view(-1, cin))
            if wants:
                if conv.weight.requires_grad:
                    dw = torch.nn.grad.conv2d_weight(xi, self.w.shape, dyi, stride=conv.stride, padding=conv.padding)
                    tape.add_param_grad(conv.weight, dw.float())
                if conv.bias is not None and conv.bias.requires_grad:
                    tape.add_param_grad(conv.bias, bwd_ops.colsum(dy))
        tape.record([out], bwd)
        return out, (N, oh, ow)


def upsample_nearest(tape, x, shape, oh, ow):
    N, h, w = shape
    C = x.t.shape[1]
    y = ops.resize_nearest(x.t.view(N, h, w, C), oh, ow)
    out = Var(y.view(-1, C))
    tape.record([out], lambda dy: x.accumulate(bwd_ops.resize_nearest_bwd(dy.view(N, oh, ow, C), h, w).view(-1, C)))
    return out, (N, oh, ow)


def avgpool2(tape, x, shape):
    N, h, w = shape
    C = x.t.shape[1]
    y = ops.avgpool2(x.t.view(N, h, w, C))
    out = Var(y.view(-1, C))
    tape.record([out], lambda dy: x.accumulate(bwd_ops.avgpool2_bwd(dy.view(N, h // 2, w // 2, C), h, w).view(-1, C)))
    return out, (N, h // 2, w // 2)


def mask_modulate(tape, x, shape, mask, ry, rx):
    N, h, w = shape
    C = x.t.shape[1]
    out = Var(ops.mask_modulate(x.t.view(N, h, w, C), mask, ry, rx).view(-1, C))
    tape.record([out], lambda dy: x.accumulate(ops.mask_modulate(dy.view(N, h, w, C), mask, ry, rx).view(-1, C)))
    return out


# --------------------------------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------------------------------
class TrainAttention:
    """One attention of the training graph: un-folded projections, the inference attention kernels for the forward, the
    [token, q|k|v] tensor kept for fmc_attention_bwd_bf16."""

    def __init__(self, attn, device, temporal=False):
        from .fmc.models.attention_processor import (LoRAAttnProcessor, LORAPoseAdaptorAttnProcessor,
                                                     PoseAdaptorAttnProcessor)
        proc = attn.processor
        self.heads = attn.heads
        C = attn.to_q.weight.shape[0]
        self.C, self.d = C, C // attn.heads
        self.hs = (self.d + 15) // 16 * 16
        self.scale = attn.scale
        self.is_cross = bool(getattr(attn, "is_cross_attention", False))
        has_lora = isinstance(proc, (LoRAAttnProcessor, LORAPoseAdaptorAttnProcessor))
        ls = proc.lora_scale if has_lora else 0.0
        lins = {"to_q": attn.to_q, "to_k": attn.to_k, "to_v": attn.to_v, "to_out": attn.to_out[0]}
        train_proj = any(l.weight.requires_grad or (l.bias is not None and l.bias.requires_grad) for l in lins.values())
        if has_lora and (train_proj or any(p.requires_grad for n, p in proc.named_parameters() if "lora" in n)):
            raise NotImplementedError("training Domain-LoRA or LoRA-carrying projections (stage 1 / train_image_lora) needs "
                                      "the un-folded path; the CMC / OMC stages keep them frozen")
        if any(lins[n].bias is not None for n in ("to_q", "to_k", "to_v")):
            raise NotImplementedError("attention_bias=True is not used by FMC")

        def folded(name):
            return engine._fold_lora(lins[name].weight, getattr(proc, f"{name}_lora") if has_lora else None, ls)
        heads, d, hs = self.heads, self.d, self.hs
        rescale = float(getattr(attn, "rescale_output_factor", 1.0))

        def qkv_current():
            return torch.cat([engine._pad_heads(folded("to_q"), heads, d, hs), engine._pad_heads(folded("to_k"), heads, d, hs),
                              folded("to_v")], dim=0), None

        def out_current():
            ol = lins["to_out"]
            return folded("to_out") / rescale, (ol.bias.detach().float() / rescale if ol.bias is not None else None)
        wq = engine._pad_heads(folded("to_q"), heads, d, hs)
        wk = engine._pad_heads(folded("to_k"), heads, d, hs)
        wv = folded("to_v")

        def unpad(dw):  # [heads * hs, K] -> [heads * d, K]
            return dw.view(heads, hs, -1)[:, :d].reshape(heads * d, -1) if hs != d else dw

        def qkv_sink(tape, dw, db):  # trainable projections (CameraEncoder temporal blocks): split the fused gradient
            k0, v0 = heads * hs, 2 * heads * hs
            for lin, g in ((attn.to_q, unpad(dw[:k0])), (attn.to_k, unpad(dw[k0:v0])), (attn.to_v, dw[v0:])):
                if lin.weight.requires_grad:
                    tape.add_param_grad(lin.weight, g)
        if self.is_cross:
            assert not train_proj, "trainable cross-attention projections are outside the FMC trainable sets"
            self.q = TrainLinear(wq, None, device)
            self.kv = TrainLinear(torch.cat([wk, wv], dim=0), None, device)
        else:
            self.qkv = TrainLinear(torch.cat([wq, wk, wv], dim=0), None, device, sink=qkv_sink if train_proj else None,
                                   live=qkv_current if train_proj else None)
        out_lin = lins["to_out"]
        out_sink = param_sink(out_lin.weight, out_lin.bias, scale=1.0 / rescale)
        self.out = TrainLinear(*out_current(), device, sink=out_sink, live=out_current if out_sink is not None else None)
        self.k0, self.v0 = self.heads * self.hs, 2 * self.heads * self.hs
        self.merge = None
        if isinstance(proc, (PoseAdaptorAttnProcessor, LORAPoseAdaptorAttnProcessor)):
            s = float(proc.scale)

            def merge_current():
                return proc.qkv_merge.weight.detach().float() * s, proc.qkv_merge.bias.detach().float() * s
            m_sink = param_sink(proc.qkv_merge.weight, proc.qkv_merge.bias, scale=s)
            self.merge = TrainLinear(*merge_current(), device, sink=m_sink, live=merge_current if m_sink is not None else None)

    def self_attention(self, tape, x, images, n, inner=1):
        """x rows -> ctx rows.  inner = 1: `images` sequences of n contiguous tokens; inner = HW: temporal (images = B * HW,
        n = frames)."""
        qkv = self.qkv(tape, x)
        ctx = torch.empty((x.t.shape[0], self.C), device=x.t.device, dtype=BF16)
        lse = None
        if inner == 1:
            # the forward hands its row log-sum-exp to the backward (which then skips its first sweep over the keys)
            lse = torch.empty((x.t.shape[0], self.heads), device=x.t.device, dtype=torch.float32)
            ops.spatial_attn(qkv.t, 0, qkv.t, self.k0, qkv.t, self.v0, self.hs, ctx, images, self.heads, self.d, n, n, 1, n,
                             self.scale, lse=lse)
        else:
            ops.temporal_attn(qkv.t, 0, self.k0, self.v0, self.hs, ctx, images // inner, n, inner, self.heads, self.d, self.scale)
        out = Var(ctx)

        def bwd(dctx):
            dqkv = torch.zeros_like(qkv.t)
            bwd_ops.attention_bwd(qkv.t, 0, qkv.t, self.k0, qkv.t, self.v0, self.hs, ctx, dctx.contiguous(), dqkv, 0, dqkv,
                                  self.k0, dqkv, self.v0, images, self.heads, self.d, n, n, 1, n, inner, self.scale, lse=lse)
            qkv.accumulate(dqkv)
        tape.record([out], bwd)
        return out

    def cross_attention(self, tape, x, kv_rows, images, n, text_len, frames):
        q = self.q(tape, x)
        ctx = torch.empty((x.t.shape[0], self.C), device=x.t.device, dtype=BF16)
        ops.spatial_attn(q.t, 0, kv_rows, 0, kv_rows, self.k0, self.hs, ctx, images, self.heads, self.d, n, text_len, frames,
                         TEXT_PAD, self.scale)
        out = Var(ctx)

        def bwd(dctx):
            dq = torch.zeros_like(q.t)
            bwd_ops.attention_bwd(q.t, 0, kv_rows, 0, kv_rows, self.k0, self.hs, ctx, dctx.contiguous(), dq, 0, None, 0, None,
                                  0, images, self.heads, self.d, n, text_len, frames, TEXT_PAD, 1, self.scale)
            q.accumulate(dq)
        tape.record([out], bwd)
        return out


# --------------------------------------------------------------------------------------------------------------
# module-level training plans (cached on the modules as `_tplan`; dropped with the inference plans)
# --------------------------------------------------------------------------------------------------------------
def _tplan(mod, device, build):
    key = (torch.device(device), "train")
    p = getattr(mod, "_tplan", None)
    if p is None or p["key"] != key:
        p = build()
        p["key"] = key
        mod._tplan = p
    return p


def invalidate(root):
    for m in root.modules():
        if hasattr(m, "_tplan"):
            m._tplan = None


def ff_block(tape, ff, norm, h, device):
    """h + FeedForward(LN(h)) with the GEGLU projection kept for backward."""
    p = _tplan(ff, device, lambda: {"ff1": linear_of(ff.net[0].proj, device, geglu=True), "ff2": linear_of(ff.net[2], device)})
    n = layernorm(tape, h, norm)
    proj = p["ff1"](tape, n)
    return p["ff2"](tape, geglu(tape, proj), residual=h)


def temporal_block(tape, block, h, B, F, HW, pose, device):
    """TemporalTransformerBlock.run (fmc/models/motion_module.py:287-300) on rows; `pose`: Var of pose-feature rows or None."""
    p = _tplan(block, device, lambda: {"attn": [TrainAttention(a, device, temporal=True) for a in block.attention_blocks]})
    for ta, attn, norm in zip(p["attn"], block.attention_blocks, block.norms):
        pe = attn.pos_encoder.pe[0].detach().to(device=device, dtype=F32).contiguous() if attn.pos_encoder is not None else None
        if pe is not None and F > pe.shape[0]:
            raise ValueError(f"{F} frames exceed the positional-encoding length {pe.shape[0]}")
        if ta.merge is not None:
            if pose is None:
                raise ValueError("PoseAdaptorAttnProcessor needs a pose_feature (attention_processor.py:210)")
            x, xp = layernorm(tape, h, norm, pe=pe, F=F, HW=HW, add=pose)
            src = ta.merge(tape, xp, residual=x)  # m = qkv_merge(x + pose) * s + x   (attention_processor.py:257)
        else:
            src = layernorm(tape, h, norm, pe=pe, F=F, HW=HW)
        ctx = ta.self_attention(tape, src, B * HW, F, inner=HW)
        h = ta.out(tape, ctx, residual=h)
    return ff_block(tape, block.ff, block.ff_norm, h, device)


def motion_module(tape, mm, x, dims, pose, device):
    """VanillaTemporalModule (motion_module.py:210-234): GN -> proj_in -> blocks -> proj_out -> + input."""
    tt = mm.temporal_transformer
    B, F, H, W, C = dims
    p = _tplan(tt, device, lambda: {"proj_in": linear_of(tt.proj_in, device), "proj_out": linear_of(tt.proj_out, device)})
    n = groupnorm(tape, x, tt.norm, B * F, H * W)
    h = p["proj_in"](tape, n)
    for block in tt.transformer_blocks:
        h = temporal_block(tape, block, h, B, F, H * W, pose, device)
    return p["proj_out"](tape, h, residual=x)


def transformer2d(tape, mod, x, dims, text_rows, text_len, device):
    """diffusers Transformer2DModel (unet_blocks.py:407): GN -> 1x1 -> [self, text cross, FF] -> 1x1 -> + input."""
    B, F, H, W, C = dims
    images, N = B * F, H * W

    def build():
        blocks = [{"attn1": TrainAttention(b.attn1, device), "attn2": TrainAttention(b.attn2, device)}
                  for b in mod.transformer_blocks]
        return {"blocks": blocks, "proj_in": linear_of(mod.proj_in, device), "proj_out": linear_of(mod.proj_out, device)}
    p = _tplan(mod, device, build)
    n = groupnorm(tape, x, mod.norm, images, N)
    h = p["proj_in"](tape, n)
    for bp, blk in zip(p["blocks"], mod.transformer_blocks):
        n1 = layernorm(tape, h, blk.norm1)
        h = bp["attn1"].out(tape, bp["attn1"].self_attention(tape, n1, images, N), residual=h)
        n2 = layernorm(tape, h, blk.norm2)
        kv = ops.gemm(text_rows, bp["attn2"].kv.w)  # text is frozen: no tape entry
        h = bp["attn2"].out(tape, bp["attn2"].cross_attention(tape, n2, kv, images, N, text_len, F), residual=h)
        h = ff_block(tape, blk.ff, blk.norm3, h, device)
    return p["proj_out"](tape, h, residual=x)


def resnet(tape, mod, x, dims, temb_act, device):
    """diffusers ResnetBlock2D per frame (unet_blocks.py:402-404); the time embedding is a constant of the step."""
    B, F, H, W, C = dims
    images, HW = B * F, H * W

    def build():
        return {"conv1": TrainConv(mod.conv1, device), "conv2": TrainConv(mod.conv2, device),
                "temb": engine.LinearPlan(mod.time_emb_proj.weight.detach().float(), mod.time_emb_proj.bias.detach().float(),
                                          device),
                "shortcut": linear_of(mod.conv_shortcut, device) if mod.conv_shortcut is not None else None}
    p = _tplan(mod, device, build)
    assert float(mod.output_scale_factor) == 1.0
    n1 = groupnorm(tape, x, mod.norm1, images, HW, silu=True)
    h, _ = p["conv1"](tape, n1, (images, H, W))
    tproj = p["temb"].f32out(temb_act)  # [B, Cout] fp32, broadcast over frames inside the second GroupNorm
    n2 = groupnorm(tape, h, mod.norm2, images, HW, silu=True, rowbias=tproj, rowbias_div=F)
    h2, _ = p["conv2"](tape, n2, (images, H, W))
    sc = p["shortcut"](tape, x) if p["shortcut"] is not None else x
    return add(tape, sc, h2), (B, F, H, W, h2.t.shape[1])


# --------------------------------------------------------------------------------------------------------------
# U-Net
# --------------------------------------------------------------------------------------------------------------
def unet_forward(tape, unet, sample, timestep, text, pose_feats, traj_feats=None):
    """Training forward of UNet3DConditionModel{PoseCond,CamObjCond} (unet.py:1033-1300): sample [B, 4, F, h, w] fp32,
    text [B, 77, 768], pose_feats: 4 Vars of rows [(B F h_l w_l), C_l], traj_feats: 4 Vars or None (injected after the
    motion modules of the cross-attention down blocks, modified_modules.py:115-117; feature 3 is unused).  Returns the Var of
    the conv_out rows [(B F h w), 32] (4 real channels)."""
    device = sample.device
    B, _, F, H, W = sample.shape
    p = unet.plan(device)
    adt = BF16
    timesteps = timestep
    if not torch.is_tensor(timesteps):
        timesteps = torch.tensor([timesteps], dtype=F32, device=device)
    elif timesteps.ndim == 0:
        timesteps = timesteps[None]
    timesteps = timesteps.to(device=device, dtype=F32).expand(B).contiguous()
    t_emb = ops.timestep_embedding(timesteps, unet.config.block_out_channels[0], dtype=adt)
    emb = p["t2"].f32out(ops.cast_act(p["t1"].f32out(t_emb), silu=True, dtype=adt))
    temb_act = ops.cast_act(emb, silu=True, dtype=adt)
    text_rows, text_len = engine.prepare_text(text, device)

    x_cl = ops.to_channels_last(sample, c_pad=64, dtype=adt)
    y = p["conv_in"](x_cl.view(B * F, H, W, 64))
    x = const(y.view(-1, y.shape[-1]))  # the latents need no gradient
    dims = (B, F, H, W, y.shape[-1])
    x.needs_grad = False

    def run_layers(block, x, dims, level, has_attn):
        outs = []
        for i, res in enumerate(block.resnets):
            x, dims = resnet(tape, res, x, dims, temb_act, device)
            if has_attn:
                x = transformer2d(tape, block.attentions[i], x, dims, text_rows, text_len, device)
            mm = block.motion_modules[i] if block.motion_modules is not None and len(block.motion_modules) > i else None
            if mm is not None:
                x = motion_module(tape, mm, x, dims, pose_feats[level], device)
            outs.append((x, dims))
        return x, dims, outs

    skips = [(x, dims)]
    for level, block in enumerate(unet.down_blocks):
        has_attn = getattr(block, "has_cross_attention", False)
        x, dims, outs = run_layers(block, x, dims, level, has_attn)
        idx = getattr(block, "traj_fea_idx", None)
        if traj_feats is not None and has_attn and idx is not None:
            # Adapted_CrossAttnDownBlock3D_forward (modified_modules.py:52-127): h = h + traj[idx]; the last skip is the sum
            x = add(tape, x, traj_feats[idx])
            outs[-1] = (x, dims)
        skips += outs
        if block.downsamplers is not None:
            for d in block.downsamplers:
                dp = _tplan(d, device, lambda d=d: {"conv": TrainConv(d.conv, device)})
                Bc, Fc, Hc, Wc, Cc = dims
                x, (_, oh, ow) = dp["conv"](tape, x, (Bc * Fc, Hc, Wc))
                dims = (Bc, Fc, oh, ow, x.t.shape[1])
            skips.append((x, dims))

    mid = unet.mid_block
    x, dims = resnet(tape, mid.resnets[0], x, dims, temb_act, device)
    for attn, res, mm in zip(mid.attentions, mid.resnets[1:], mid.motion_modules):
        x = transformer2d(tape, attn, x, dims, text_rows, text_len, device)
        if mm is not None:
            x = motion_module(tape, mm, x, dims, pose_feats[-1], device)
        x, dims = resnet(tape, res, x, dims, temb_act, device)

    n_levels = len(unet.up_blocks)
    for i, block in enumerate(unet.up_blocks):
        has_attn = getattr(block, "has_cross_attention", False)
        level = n_levels - 1 - i
        pose = pose_feats[level] if unet.decoder_add_posecond else None
        for j, res in enumerate(block.resnets):
            skip, sdims = skips.pop()
            x = concat(tape, x, skip)
            dims = dims[:4] + (x.t.shape[1],)
            x, dims = resnet(tape, res, x, dims, temb_act, device)
            if has_attn:
                x = transformer2d(tape, block.attentions[j], x, dims, text_rows, text_len, device)
            mm = block.motion_modules[j] if block.motion_modules is not None and len(block.motion_modules) > j else None
            if mm is not None:
                x = motion_module(tape, mm, x, dims, pose, device)
        if block.upsamplers is not None:
            for u in block.upsamplers:
                up = _tplan(u, device, lambda u=u: {"conv": TrainConv(u.conv, device)})
                Bc, Fc, Hc, Wc, Cc = dims
                x, shp = upsample_nearest(tape, x, (Bc * Fc, Hc, Wc), 2 * Hc, 2 * Wc)
                x, (_, oh, ow) = up["conv"](tape, x, shp)
                dims = (Bc, Fc, oh, ow, x.t.shape[1])

    Bc, Fc, Hc, Wc, C = dims
    n = groupnorm(tape, x, unet.conv_norm_out, Bc * Fc, Hc * Wc, silu=True)
    co = _tplan(unet.conv_out, device, lambda: {"conv": TrainConv(unet.conv_out, device)})
    out, _ = co["conv"](tape, n, (Bc * Fc, Hc, Wc))
    return out, dims[:4]


# --------------------------------------------------------------------------------------------------------------
# encoders
# --------------------------------------------------------------------------------------------------------------
def _encoder_resblock(tape, rb, x, shape, device):
    """ResnetBlock of the CameraEncoder / ObjectEncoder (pose_adaptor.py:102-135, adapter.py:64-98)."""
    p = _tplan(rb, device, lambda: {k: (TrainConv(m, device) if m is not None else None) for k, m in
                                    (("in_conv", rb.in_conv), ("block1", rb.block1), ("block2", rb.block2), ("skep", rb.skep),
                                     ("down", rb.down_opt.op if (rb.down and rb.down_opt.use_conv) else None))})
    if rb.down:
        x, shape = p["down"](tape, x, shape) if p["down"] is not None else avgpool2(tape, x, shape)
    if p["in_conv"] is not None:
        x, shape = p["in_conv"](tape, x, shape)
    h, _ = p["block1"](tape, x, shape)
    h = relu(tape, h)
    h, _ = p["block2"](tape, h, shape)
    skip = p["skep"](tape, x, shape)[0] if p["skep"] is not None else x
    return add(tape, h, skip), shape


def camera_encoder_forward(tape, enc, pose_embedding):
    """CameraPoseEncoder.forward (pose_adaptor.py:224-240) for training: pose_embedding [b, 6, f, H, W] -> 4 Vars of rows
    [(b f h_l w_l), C_l] in the U-Net's channels-last order."""
    from .fmc.models.pose_adaptor import unshuffle8_to_cl
    device = pose_embedding.device
    x_cl = unshuffle8_to_cl(pose_embedding.float())
    b, f, h, w, cin = x_cl.shape
    p = _tplan(enc, device, lambda: {"conv_in": TrainConv(enc.encoder_conv_in, device)})
    x = const(x_cl.view(-1, cin))
    x, shape = p["conv_in"](tape, x, (b * f, h, w))
    feats = []
    for res_block, attn_block in zip(enc.encoder_down_conv_blocks, enc.encoder_down_attention_blocks):
        for rb, tb in zip(res_block, attn_block):
            x, shape = _encoder_resblock(tape, rb, x, shape, device)
            x = temporal_block(tape, tb, x, b, f, shape[1] * shape[2], None, device)
        feats.append(x)
    return feats


def adapter_forward(tape, omcm, x_cl, mask):
    """ObjectEncoder `Adapter.encode_cl` (adapter.py:154-192) for training: x_cl [N, H/8, W/8, cin] bf16 (the scattered
    object features), mask [N, H, W] fp32 -> 4 Vars of rows [(N h_l w_l), C_l]."""
    device = x_cl.device
    N, h, w, cin = x_cl.shape
    p = _tplan(omcm, device, lambda: {
        "conv_in": TrainConv(omcm.conv_in, device),
        "zero_in": TrainConv(omcm.zero_conv_in, device) if isinstance(omcm.zero_conv_in, nn.Conv2d) else None,
        "zero_out": [TrainConv(m, device) if isinstance(m, nn.Conv2d) else None for m in omcm.zero_conv_out_list]})
    x, shape = const(x_cl.reshape(-1, cin)), (N, h, w)
    if p["zero_in"] is not None:
        x, shape = p["zero_in"](tape, x, shape)
    x, shape = p["conv_in"](tape, x, shape)
    feats = []
    sizes_h, sizes_w = [mask.shape[1]], [mask.shape[2]]
    for i in range(len(omcm.channels)):
        for j in range(omcm.nums_rb):
            x, shape = _encoder_resblock(tape, omcm.body[i * omcm.nums_rb + j], x, shape, device)
        if p["zero_out"][i] is not None:
            x, shape = p["zero_out"][i](tape, x, shape)
        sizes_h.append(shape[1])
        sizes_w.append(shape[2])
        ry = engine.nearest_index_on(sizes_h, device)
        rx = engine.nearest_index_on(sizes_w, device)
        x = mask_modulate(tape, x, shape, mask, ry, rx)
        feats.append(x)
    return feats


# --------------------------------------------------------------------------------------------------------------
# torch.autograd bridge
# --------------------------------------------------------------------------------------------------------------
class _TapeFunction(torch.autograd.Function):
    """One node of torch's graph for a whole tape: inputs = the trainable parameters + the input tensors that require grad
    (e.g. object features produced by another tape), outputs = the tape's results in the reference layout."""

    @staticmethod
    def forward(ctx, runner, *inputs):
        ctx.runner = runner
        ctx.set_materialize_grads(False)  # an unused output (ObjectEncoder feature 3, ...) stays None: its branch is skipped
        outs = tuple(runner.outputs)
        return outs if len(outs) > 1 else outs[0]

    @staticmethod
    def backward(ctx, *grad_outs):
        return (None,) + tuple(ctx.runner.run_backward(grad_outs))


class _Runner:
    def __init__(self, tape, out_vars, outputs, out_pad, params, in_vars, in_shapes):
        """out_vars[i]: Var of rows whose reference-layout tensor is outputs[i] ([B, C, F, h, w] fp32, C real channels of
        out_pad[i] stored ones); in_vars[j]: Vars of input rows that need a gradient, in_shapes[j] = (B, F, h, w, C)."""
        self.tape, self.out_vars, self.outputs, self.out_pad = tape, out_vars, outputs, out_pad
        self.params, self.in_vars, self.in_shapes = params, in_vars, in_shapes

    def run_backward(self, grad_outs):
        for var, g, pad in zip(self.out_vars, grad_outs, self.out_pad):
            if g is not None:
                cl = ops.to_channels_last(g.float().contiguous(), c_pad=pad, dtype=BF16)
                var.accumulate(cl.view(-1, cl.shape[-1]))
        with torch.no_grad():
            self.tape.backward()
        self.tape.finish_params()
        grads = [self.tape.param_grads.get(p) for p in self.params]
        for var, shp in zip(self.in_vars, self.in_shapes):
            grads.append(ops.from_channels_last(var.grad.view(shp)) if var.grad is not None else None)
            var.grad = None
        return grads

    def attach(self, in_tensors):
        inputs = list(self.params) + list(in_tensors)
        if not any(t.requires_grad for t in inputs):
            return tuple(self.outputs) if len(self.outputs) > 1 else self.outputs[0]
        return _TapeFunction.apply(self, *inputs)


def _trainable_params(*modules):
    return [p for m in modules if m is not None for p in m.parameters() if p.requires_grad]


def wants_training(*modules_and_tensors):
    """True when the call is part of a training step: autograd is on and a parameter / input tensor requires grad."""
    if not torch.is_grad_enabled():
        return False
    for x in modules_and_tensors:
        if x is None:
            continue
        if torch.is_tensor(x):
            if x.requires_grad:
                return True
        elif isinstance(x, (list, tuple)):
            if any(torch.is_tensor(t) and t.requires_grad for t in x):
                return True
        elif any(p.requires_grad for p in x.parameters()):
            return True
    return False


def pose_adaptor_train_forward(unet, pose_encoder, noisy_latents, timesteps, text, pose_embedding, traj_features=None):
    """`PoseAdaptor.forward` / `CamObjPoseAdaptor.forward` under autograd (train_cam_ctrl.py:586-600,
    train_cam_obj_ctrl.py:843-866): returns noise_pred [B, 4, F, h, w] fp32 attached to torch's graph through the trainable
    parameters and through `traj_features` when they require grad (object features produced by `adapter_train_forward`)."""
    ops.require_cuda(noisy_latents)
    tape = Tape()
    in_vars, in_shapes, in_tensors = [], [], []
    with torch.no_grad():
        pose_feats = camera_encoder_forward(tape, pose_encoder, pose_embedding)
        trajs = None
        if traj_features is not None:
            trajs = []
            for t in traj_features:
                cl = engine.CL.from_reference(t.detach() if torch.is_tensor(t) else t)
                v = Var(cl.rows(), needs_grad=torch.is_tensor(t) and t.requires_grad)
                if v.needs_grad:
                    in_vars.append(v)
                    in_shapes.append(cl.dims)
                    in_tensors.append(t)
                trajs.append(v)
        out, dims = unet_forward(tape, unet, noisy_latents, timesteps, text, pose_feats, trajs)
        B, F, H, W = dims
        n_out = unet.conv_out.out_channels
        pred = ops.from_channels_last(out.t.view(B, F, H, W, out.t.shape[1]), C=n_out)
    runner = _Runner(tape, [out], [pred], [out.t.shape[1]], _trainable_params(unet, pose_encoder), in_vars, in_shapes)
    return runner.attach(in_tensors)


def adapter_train_forward(omcm, x_cl, mask, b, f):
    """ObjectEncoder under autograd (get_traj_features_v2, fmc/util.py:147-213 -> Adapter.forward): x_cl [(b f), H/8, W/8, 832]
    bf16 scattered object features, mask [(b f), H, W]; returns the 4 features [b, C_l, f, h_l, w_l] fp32 attached to torch's
    graph through the ObjectEncoder's trainable parameters."""
    tape = Tape()
    with torch.no_grad():
        feats = adapter_forward(tape, omcm, x_cl, mask)
        outs, pads = [], []
        N = x_cl.shape[0]
        for v in feats:
            C = v.t.shape[1]
            hw = v.t.shape[0] // N
            h_l = None
            for cand in range(1, hw + 1):  # recover (h_l, w_l) from the aspect ratio of the input grid
                if hw % cand == 0 and cand * x_cl.shape[2] == (hw // cand) * x_cl.shape[1]:
                    h_l = cand
                    break
            assert h_l is not None
            outs.append(ops.from_channels_last(v.t.view(b, f, h_l, hw // h_l, C)))
            pads.append(C)
    runner = _Runner(tape, feats, outs, pads, _trainable_params(omcm), [], [])
    res = runner.attach([])
    return list(res) if isinstance(res, tuple) else [res]
