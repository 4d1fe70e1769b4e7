"""TEST INFRASTRUCTURE ONLY (never imported by the product): fp32 CPU restatement of diffusers==0.24.0 `AutoencoderKL` as the
reference uses it at the pipeline edges -- `vae.decode(latents[i:i+1]).sample` frame by frame
(fmc/pipelines/pipeline_animation.py:465-478) and `vae.encode(pixel_values).latent_dist.sample() * 0.18215`
(train_cam_ctrl.py:544).  diffusers is a pinned third-party dependency of the reference (environment.yaml:13) that is neither
vendored nor installed here, so the module structure and state-dict keys below restate its published SD1.5 VAE
(`AutoencoderKL`, `Encoder`, `Decoder`, `DownEncoderBlock2D`, `UpDecoderBlock2D`, `UNetMidBlock2D`, `ResnetBlock2D` with
`temb_channels=None`, `Attention` with `group_norm`, `residual_connection=True`, heads = 1).

PINNING.  The same architecture exists, written independently, in the installed `torchtitan` package
(torchtitan/experiments/flux/model/autoencoder.py: the original CompVis/LDM `Encoder` / `Decoder` that diffusers ported).
`tests/test_oracle_edges.py` maps weights through the published LDM -> diffusers key correspondence (`LDM_TO_DIFFUSERS`) and
requires this restatement to reproduce that implementation's encoder and decoder outputs to 1e-5.  The two 1x1 convolutions
diffusers adds (`quant_conv`, `post_quant_conv`) and the `DiagonalGaussianDistribution` are not in that implementation and
remain restated-only."""
import re

import torch
import torch.nn.functional as F
from torch import nn


class ResnetBlock2D(nn.Module):
    """diffusers ResnetBlock2D(temb_channels=None, groups=32, eps=1e-6, output_scale_factor=1.0, non_linearity='silu')"""

    def __init__(self, cin, cout, groups=32, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class VaeAttention(nn.Module):
    """diffusers Attention(C, heads=C // attention_head_dim = 1, dim_head=C, bias=True, norm_num_groups=32, eps=1e-6,
    residual_connection=True, rescale_output_factor=1) on a [b, C, h, w] input (the deprecated AttentionBlock's successor)"""

    def __init__(self, C, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, C, eps=eps)
        self.to_q = nn.Linear(C, C)
        self.to_k = nn.Linear(C, C)
        self.to_v = nn.Linear(C, C)
        self.to_out = nn.ModuleList([nn.Linear(C, C), nn.Dropout(0.0)])
        self.scale = C ** -0.5

    def forward(self, x):
        b, c, h, w = x.shape
        hs = x.view(b, c, h * w).transpose(1, 2)
        hs = self.group_norm(hs.transpose(1, 2)).transpose(1, 2)
        q, k, v = self.to_q(hs), self.to_k(hs), self.to_v(hs)
        probs = torch.softmax(torch.baddbmm(torch.empty(b, h * w, h * w), q, k.transpose(1, 2), beta=0, alpha=self.scale), dim=-1)
        out = self.to_out[0](torch.bmm(probs, v))
        return out.transpose(1, 2).reshape(b, c, h, w) + x


class MidBlock(nn.Module):
    def __init__(self, C, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(C, C, groups), ResnetBlock2D(C, C, groups)])
        self.attentions = nn.ModuleList([VaeAttention(C, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _Conv(nn.Module):
    def __init__(self, C, stride, padding):
        super().__init__()
        self.conv = nn.Conv2d(C, C, 3, stride=stride, padding=padding)


class Downsample2D(_Conv):
    """use_conv=True, padding=0: asymmetric zero pad (right, bottom) then a stride-2 convolution"""

    def __init__(self, C):
        super().__init__(C, 2, 0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class Upsample2D(_Conv):
    def __init__(self, C):
        super().__init__(C, 1, 1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UpDecoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_upsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Encoder(nn.Module):
    def __init__(self, in_channels, latent_channels, block_out_channels, layers_per_block, groups):
        super().__init__()
        ch = block_out_channels
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList([
            DownEncoderBlock2D(ch[max(i - 1, 0)], ch[i], layers_per_block, groups, add_downsample=i != len(ch) - 1)
            for i in range(len(ch))])
        self.mid_block = MidBlock(ch[-1], groups)
        self.conv_norm_out = nn.GroupNorm(groups, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 2 * latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for blk in self.down_blocks:
            x = blk(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, latent_channels, out_channels, block_out_channels, layers_per_block, groups):
        super().__init__()
        ch = list(reversed(block_out_channels))
        self.conv_in = nn.Conv2d(latent_channels, ch[0], 3, padding=1)
        self.mid_block = MidBlock(ch[0], groups)
        self.up_blocks = nn.ModuleList([
            UpDecoderBlock2D(ch[max(i - 1, 0)], ch[i], layers_per_block + 1, groups, add_upsample=i != len(ch) - 1)
            for i in range(len(ch))])
        self.conv_norm_out = nn.GroupNorm(groups, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for blk in self.up_blocks:
            x = blk(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    def __init__(self, moments):
        self.mean, logvar = torch.chunk(moments, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None, noise=None):
        """`noise`: the N(0, 1) draw to use (tests pass the same draw to both sides); else drawn from `generator`"""
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class _Out:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class AutoencoderKL(nn.Module):
    """SD1.5 defaults: block_out_channels (128, 256, 512, 512), layers_per_block 2, latent_channels 4, 32 groups,
    scaling_factor 0.18215"""

    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 latent_channels=4, norm_num_groups=32, scaling_factor=0.18215):
        super().__init__()
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)
        self.config = _Out(scaling_factor=scaling_factor, latent_channels=latent_channels,
                           block_out_channels=tuple(block_out_channels), in_channels=in_channels, out_channels=out_channels,
                           layers_per_block=layers_per_block, norm_num_groups=norm_num_groups)

    def encode(self, x):
        return _Out(latent_dist=DiagonalGaussianDistribution(self.quant_conv(self.encoder(x))))

    def decode(self, z):
        return _Out(sample=self.decoder(self.post_quant_conv(z)))


# Published key correspondence of the CompVis/LDM autoencoder (as in torchtitan's flux AutoEncoder) to diffusers' AutoencoderKL
LDM_TO_DIFFUSERS = (
    (r"^(encoder|decoder)\.mid\.block_(\d)\.", lambda m: f"{m.group(1)}.mid_block.resnets.{int(m.group(2)) - 1}."),
    (r"^(encoder|decoder)\.mid\.attn_1\.norm\.", lambda m: f"{m.group(1)}.mid_block.attentions.0.group_norm."),
    (r"^(encoder|decoder)\.mid\.attn_1\.q\.", lambda m: f"{m.group(1)}.mid_block.attentions.0.to_q."),
    (r"^(encoder|decoder)\.mid\.attn_1\.k\.", lambda m: f"{m.group(1)}.mid_block.attentions.0.to_k."),
    (r"^(encoder|decoder)\.mid\.attn_1\.v\.", lambda m: f"{m.group(1)}.mid_block.attentions.0.to_v."),
    (r"^(encoder|decoder)\.mid\.attn_1\.proj_out\.", lambda m: f"{m.group(1)}.mid_block.attentions.0.to_out.0."),
    (r"^encoder\.down\.(\d)\.block\.(\d)\.", lambda m: f"encoder.down_blocks.{m.group(1)}.resnets.{m.group(2)}."),
    (r"^encoder\.down\.(\d)\.downsample\.", lambda m: f"encoder.down_blocks.{m.group(1)}.downsamplers.0."),
    (r"^decoder\.up\.(\d)\.block\.(\d)\.", lambda m: f"decoder.up_blocks.{3 - int(m.group(1))}.resnets.{m.group(2)}."),
    (r"^decoder\.up\.(\d)\.upsample\.", lambda m: f"decoder.up_blocks.{3 - int(m.group(1))}.upsamplers.0."),
    (r"^(encoder|decoder)\.norm_out\.", lambda m: f"{m.group(1)}.conv_norm_out."),
)


def ldm_state_to_diffusers(state):
    """LDM-keyed encoder / decoder state dict -> AutoencoderKL keys (1x1-conv attention weights become Linear weights)"""
    out = {}
    for k, v in state.items():
        nk = k.replace("nin_shortcut", "conv_shortcut")
        for pat, fn in LDM_TO_DIFFUSERS:
            nk, n = re.subn(pat, fn, nk)
            if n:
                break
        if ".attentions.0.to_" in nk and nk.endswith("weight") and v.dim() == 4:
            v = v[:, :, 0, 0]
        out[nk] = v
    return out
