"""diffusers==0.24.0 `AutoencoderKL` (SD1.5 VAE) on the B200 kernels.

Reference call sites: `self.vae.decode(latents[frame_idx:frame_idx+1]).sample` frame by frame, then
`(video / 2 + 0.5).clamp(0, 1)` (fmc/pipelines/pipeline_animation.py:465-478); `vae.encode(pixel_values).latent_dist.sample()
* 0.18215` (train_cam_ctrl.py:544-545).  Parameter holders carry the diffusers state-dict keys (`encoder.down_blocks.i.
resnets.j.{norm1,conv1,norm2,conv2,conv_shortcut}`, `...downsamplers.0.conv`, `mid_block.{resnets.k,attentions.0.{group_norm,
to_q,to_k,to_v,to_out.0}}`, `decoder.up_blocks.i...upsamplers.0.conv`, `conv_norm_out`, `conv_out`, `quant_conv`,
`post_quant_conv`), so a real SD1.5 VAE checkpoint loads by key.

Execution (channels-last rows [(images h w), C], all frames of a clip at once instead of the reference's per-frame loop):
GroupNorm(+SiLU) = fmc_groupnorm_bf16; 3x3 convolutions through engine.ConvPlan (the convolution policy of the denoising
path: tcgen05 implicit GEMM up to 64 input channels, cuDNN for the wide ones); 1x1 shortcuts and the attention projections =
fmc_gemm_bf16 with the residual in the epilogue; the mid-block attention (1 head of 512 over h*w tokens, per image) = GEMM
(fp32 scores) -> fmc_softmax_rows -> GEMM; nearest 2x upsample = fmc_resize_nearest_bf16; the posterior sample and the
decoder's output conversion + clamp are fused single passes (fmc_vae_sample_f32, fmc_cl_to_video_f32)."""
import contextlib

import torch
from torch import nn

from .. import bwd_ops, engine, ops
from ..fmc._blocks import _Holder

BF16, F32 = torch.bfloat16, torch.float32


def _device_ctx(device):
    """make the tensor's GPU current for the calls below (the kernels launch on the current device's current stream)"""
    return torch.cuda.device(device) if torch.device(device).type == "cuda" else contextlib.nullcontext()


class _Resnet(_Holder):
    """diffusers ResnetBlock2D(temb_channels=None)"""

    def __init__(self, cin, cout, groups, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None


class _Attention(_Holder):
    """diffusers Attention(C, heads=1, dim_head=C, bias=True, norm_num_groups, residual_connection=True)"""

    def __init__(self, C, groups, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, C, eps=eps)
        self.to_q = nn.Linear(C, C)
        self.to_k = nn.Linear(C, C)
        self.to_v = nn.Linear(C, C)
        self.to_out = nn.ModuleList([nn.Linear(C, C), nn.Dropout(0.0)])


class _Mid(_Holder):
    def __init__(self, C, groups):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(C, C, groups), _Resnet(C, C, groups)])
        self.attentions = nn.ModuleList([_Attention(C, groups)])


class _Sampler(_Holder):
    def __init__(self, C, stride, padding):
        super().__init__()
        self.conv = nn.Conv2d(C, C, 3, stride=stride, padding=padding)


class _Block(_Holder):
    def __init__(self, cin, cout, layers, groups, down=False, up=False):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        if down:
            self.downsamplers = nn.ModuleList([_Sampler(cout, 2, 0)])  # Downsample2D(padding=0): pad right / bottom by one
        if up:
            self.upsamplers = nn.ModuleList([_Sampler(cout, 1, 1)])


class _Encoder(_Holder):
    def __init__(self, in_channels, latent_channels, ch, layers, groups):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList([_Block(ch[max(i - 1, 0)], ch[i], layers, groups, down=i != len(ch) - 1)
                                          for i in range(len(ch))])
        self.mid_block = _Mid(ch[-1], groups)
        self.conv_norm_out = nn.GroupNorm(groups, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 2 * latent_channels, 3, padding=1)


class _Decoder(_Holder):
    def __init__(self, latent_channels, out_channels, ch, layers, groups):
        super().__init__()
        ch = list(reversed(ch))
        self.conv_in = nn.Conv2d(latent_channels, ch[0], 3, padding=1)
        self.mid_block = _Mid(ch[0], groups)
        self.up_blocks = nn.ModuleList([_Block(ch[max(i - 1, 0)], ch[i], layers + 1, groups, up=i != len(ch) - 1)
                                        for i in range(len(ch))])
        self.conv_norm_out = nn.GroupNorm(groups, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], out_channels, 3, padding=1)


class _Config:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _Output:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class LatentDistribution:
    """diffusers DiagonalGaussianDistribution over channels-last moment rows kept on the device."""

    def __init__(self, moments_rows, N, z, h, w):
        self._m, self._N, self._z, self._h, self._w = moments_rows, N, z, h, w

    def sample(self, generator=None, noise=None, scale=1.0):
        """mean + std * noise, [N, z, h, w] fp32.  `noise`: an N(0, 1) draw in that shape (else drawn here from
        `generator` on the device); `scale` multiplies the result in the same pass (the trainers' `* 0.18215`)."""
        shape = (self._N, self._z, self._h, self._w)
        if noise is None:
            noise = torch.randn(shape, generator=generator, device=self._m.device, dtype=F32)
        noise = noise.to(device=self._m.device, dtype=F32).contiguous()
        assert tuple(noise.shape) == shape
        return ops.vae_sample(self._m, self._N, self._z, self._h * self._w, noise, scale).view(shape)

    def mode(self):
        return ops.vae_sample(self._m, self._N, self._z, self._h * self._w).view(self._N, self._z, self._h, self._w)


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, down_block_types=None, up_block_types=None,
                 block_out_channels=(128, 256, 512, 512), layers_per_block=2, act_fn="silu", latent_channels=4,
                 norm_num_groups=32, sample_size=512, scaling_factor=0.18215, **unused):
        super().__init__()
        assert act_fn == "silu", "the SD1.5 VAE uses SiLU"
        ch = tuple(block_out_channels)
        self.encoder = _Encoder(in_channels, latent_channels, ch, layers_per_block, norm_num_groups)
        self.decoder = _Decoder(latent_channels, out_channels, ch, layers_per_block, norm_num_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)
        self.config = _Config(in_channels=in_channels, out_channels=out_channels, block_out_channels=ch,
                              layers_per_block=layers_per_block, latent_channels=latent_channels,
                              norm_num_groups=norm_num_groups, sample_size=sample_size, scaling_factor=scaling_factor)
        self.requires_grad_(False)  # frozen in both trainers (train_cam_ctrl.py:246)
        self._plans = None

    # --- diffusers surface the pipelines touch -------------------------------------------------------------------
    def enable_slicing(self):
        """diffusers decodes one image at a time when slicing is on; the batched decode here needs no such switch."""

    def disable_slicing(self):
        pass

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def forward(self, *a, **k):
        raise RuntimeError("AutoencoderKL: call encode() / decode() (executed by the synfmc_b200 kernels; no eager fallback)")

    # --- plans -----------------------------------------------------------------------------------------------------
    def _plan(self, device):
        if engine.precise():
            raise NotImplementedError("the VAE runs in the bf16 mode: the reference-precision (tf32) kernels cover the "
                                      "denoising path, whose GEMM shapes they were built for (4 / 3 channel convolutions "
                                      "and the padding-0 stride-2 downsamplers are outside them)")
        key = engine.plan_key(device)
        if self._plans is None or self._plans["key"] != key or self._plans["fp"] != engine.fingerprint(self):
            plans = {"key": key, "fp": engine.fingerprint(self), "mods": {}}
            for mod in self.modules():
                if isinstance(mod, _Resnet):
                    sc = mod.conv_shortcut
                    plans["mods"][id(mod)] = {
                        "norm1": engine.NormPlan(mod.norm1, device), "conv1": engine.ConvPlan(mod.conv1, device),
                        "norm2": engine.NormPlan(mod.norm2, device), "conv2": engine.ConvPlan(mod.conv2, device),
                        "shortcut": engine.LinearPlan(sc.weight.detach().float().view(sc.out_channels, sc.in_channels),
                                                      sc.bias, device) if sc is not None else None}
                elif isinstance(mod, _Attention):
                    w = torch.cat([mod.to_q.weight, mod.to_k.weight, mod.to_v.weight], dim=0).detach().float()
                    b = torch.cat([mod.to_q.bias, mod.to_k.bias, mod.to_v.bias], dim=0).detach().float()
                    plans["mods"][id(mod)] = {
                        "norm": engine.NormPlan(mod.group_norm, device), "qkv": engine.LinearPlan(w, b, device),
                        "out": engine.LinearPlan(mod.to_out[0].weight.detach().float(), mod.to_out[0].bias, device)}
                elif isinstance(mod, nn.Conv2d) and mod in (self.quant_conv, self.post_quant_conv, self.encoder.conv_in,
                                                            self.encoder.conv_out, self.decoder.conv_in,
                                                            self.decoder.conv_out):
                    plans["mods"][id(mod)] = engine.ConvPlan(mod, device)
                elif isinstance(mod, _Sampler):
                    plans["mods"][id(mod)] = engine.ConvPlan(mod.conv, device)
            for side in (self.encoder, self.decoder):
                plans["mods"][id(side.conv_norm_out)] = engine.NormPlan(side.conv_norm_out, device)
            self._plans = plans
        return self._plans["mods"]

    # --- executors (x: [N, H, W, C] channels-last, bf16 or fp32 by engine precision) --------------------------------
    @staticmethod
    def _gn(p, x, silu):
        N, H, W, C = x.shape
        return ops.groupnorm(x.view(-1, C), p.g, p.b, p.eps, N, H * W, groups=p.groups, silu=silu).view(N, H, W, C)

    def _resnet(self, plans, mod, x):
        p = plans[id(mod)]
        N, H, W, C = x.shape
        h = p["conv1"](self._gn(p["norm1"], x, True))
        h = self._gn(p["norm2"], h, True)
        res = x if p["shortcut"] is None else p["shortcut"](x.view(-1, C)).view(N, H, W, -1)
        return p["conv2"](h, residual=res)

    def _attention(self, plans, mod, x):
        p = plans[id(mod)]
        N, H, W, C = x.shape
        HW = H * W
        if HW % 16 or HW > 4096:
            raise NotImplementedError(f"VAE attention over {H}x{W} = {HW} tokens: the token count must be a multiple of 16, "
                                      "at most 4096 (320x512 frames have 2560)")
        rows = x.view(-1, C)
        qkv = p["qkv"](self._gn(p["norm"], x, False).view(-1, C))
        ctx = torch.empty((N * HW, C), device=x.device, dtype=rows.dtype)
        scale = float(C) ** -0.5
        for i in range(N):  # one head of width C per image: scores = q k^T on the GEMM (fp32 out), softmax, P v
            blk = qkv[i * HW:(i + 1) * HW]
            q, k, v = blk[:, :C], blk[:, C:2 * C], blk[:, 2 * C:]
            if rows.dtype == F32:
                s = ops.gemm_f32(q.contiguous(), k.contiguous(), split=1)
                prob = ops.softmax_rows(s, scale, out_dtype=F32)
                ops.gemm_f32(prob, v.t().contiguous(), out=ctx[i * HW:(i + 1) * HW], split=1)
            else:
                s = ops.gemm(q, k, out_f32=True)
                prob = ops.softmax_rows(s, scale)
                ops.gemm(prob, bwd_ops.transpose(v), out=ctx[i * HW:(i + 1) * HW])
        return p["out"](ctx, residual=rows).view(N, H, W, C)

    def _mid(self, plans, mid, x):
        x = self._resnet(plans, mid.resnets[0], x)
        x = self._attention(plans, mid.attentions[0], x)
        return self._resnet(plans, mid.resnets[1], x)

    def _to_cl(self, x):
        """[N, C, H, W] (any float dtype) -> [N, H, W, C] in the activation dtype"""
        ops.require_cuda(x)
        N, C, H, W = x.shape
        return ops.to_channels_last(x.reshape(N, C, 1, H, W), dtype=engine.act_dtype()).view(N, H, W, C)

    def _decode_cl(self, z):
        with _device_ctx(z.device):
            plans = self._plan(z.device)
            dec = self.decoder
            x = plans[id(self.post_quant_conv)](self._to_cl(z))
            x = plans[id(dec.conv_in)](x)
            x = self._mid(plans, dec.mid_block, x)
            for blk in dec.up_blocks:
                for r in blk.resnets:
                    x = self._resnet(plans, r, x)
                if hasattr(blk, "upsamplers"):
                    N, H, W, C = x.shape
                    x = plans[id(blk.upsamplers[0])](ops.resize_nearest(x, 2 * H, 2 * W))
            x = self._gn(plans[id(dec.conv_norm_out)], x, True)
            return plans[id(dec.conv_out)](x)

    @torch.no_grad()
    def decode(self, z, return_dict=True):
        """z [N, latent, h, w] -> .sample [N, 3, 8h, 8w] fp32 (the reference calls this one frame at a time)."""
        y = self._decode_cl(z)
        N, H, W, C = y.shape
        with _device_ctx(z.device):
            sample = ops.cl_to_video(y.reshape(-1, C), N, C, 1, H * W).view(N, C, H, W)
        return _Output(sample=sample) if return_dict else (sample,)

    @torch.no_grad()
    def decode_video(self, latents, scaling_factor=None, chunk=None):
        """decode_latents of the reference pipelines in one call: latents [b, 4, f, h, w] -> video [b, 3, f, 8h, 8w] fp32 in
        [0, 1] on the device = clamp(decode(latents / scaling_factor) / 2 + 0.5, 0, 1); `chunk` frames are decoded per pass
        (default: all frames of a clip)."""
        b, c, f, h, w = latents.shape
        sf = self.config.scaling_factor if scaling_factor is None else scaling_factor
        frames = (latents.float() * (1.0 / sf)).permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        chunk = chunk or f
        out = torch.empty((b, self.config.out_channels, f, 8 * h, 8 * w), device=latents.device, dtype=F32)
        with _device_ctx(latents.device):
            for bi in range(b):
                for f0 in range(0, f, chunk):
                    n = min(chunk, f - f0)
                    y = self._decode_cl(frames[bi * f + f0:bi * f + f0 + n])
                    _, H, W, C = y.shape
                    out[bi, :, f0:f0 + n] = ops.cl_to_video(y.reshape(-1, C), 1, C, n, H * W, mul=0.5, add=0.5, lo=0.0,
                                                            hi=1.0).view(C, n, H, W)
        return out

    @torch.no_grad()
    def encode(self, x, return_dict=True):
        """x [N, 3, H, W] in [-1, 1] -> .latent_dist (sample() / mode() give [N, latent, H/8, W/8] fp32)."""
        with _device_ctx(x.device):
            plans = self._plan(x.device)
            enc = self.encoder
            h = plans[id(enc.conv_in)](self._to_cl(x))
            for blk in enc.down_blocks:
                for r in blk.resnets:
                    h = self._resnet(plans, r, h)
                if hasattr(blk, "downsamplers"):
                    # Downsample2D(padding=0): zero column on the right, zero row at the bottom, then the stride-2 conv
                    h = plans[id(blk.downsamplers[0])](torch.nn.functional.pad(h, (0, 0, 0, 1, 0, 1)))
            h = self._mid(plans, enc.mid_block, h)
            h = plans[id(enc.conv_out)](self._gn(plans[id(enc.conv_norm_out)], h, True))
            m = plans[id(self.quant_conv)](h)
            N, hh, ww, C2 = m.shape
            dist = LatentDistribution(m.reshape(-1, C2), N, C2 // 2, hh, ww)
        return _Output(latent_dist=dist) if return_dict else (dist,)
