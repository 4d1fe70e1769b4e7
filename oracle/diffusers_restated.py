"""ORACLE (test infrastructure, CPU fp32) -- restatement of the diffusers==0.24.0 building blocks that the
FMC reference imports but does not vendor (environment.yaml:13; call sites: fmc/models/motion_module.py:9-10,
fmc/models/attention_processor.py:5-7, fmc/models/unet_blocks.py:6-7, fmc/models/unet.py:13-20,
train_cam_ctrl.py:25).

PARITY UNPINNED: diffusers is not installed in this image and the reference holds no golden vectors, so this
file restates the published 0.24.0 semantics and is cross-checked in
tests/test_oracle_primitives.py against torch primitives that ARE installed (F.scaled_dot_product_attention,
F.gelu, F.group_norm, F.layer_norm).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
Parameter / module names are the diffusers ones so reference state-dicts load by key.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F
from torch import nn


class LoRALinearLayer(nn.Module):
    """diffusers.models.lora.LoRALinearLayer: up(down(x)) [* alpha/rank]; down ~ N(0, 1/rank), up = 0."""

    def __init__(self, in_features, out_features, rank=4, network_alpha=None):
        super().__init__()
        self.down = nn.Linear(in_features, rank, bias=False)
        self.up = nn.Linear(rank, out_features, bias=False)
        self.network_alpha = network_alpha
        self.rank = rank
        nn.init.normal_(self.down.weight, std=1 / rank)
        nn.init.zeros_(self.up.weight)

    def forward(self, hidden_states):
        orig_dtype = hidden_states.dtype
        dtype = self.down.weight.dtype
        out = self.up(self.down(hidden_states.to(dtype)))
        if self.network_alpha is not None:
            out = out * (self.network_alpha / self.rank)
        return out.to(orig_dtype)


class AttnProcessorSDPA:
    """diffusers AttnProcessor2_0 (the default processor when torch has SDPA): same math as the explicit
    baddbmm/softmax/bmm path; used by CameraPoseEncoder's temporal blocks, which the trainers never re-wire
    (fmc/models/unet.py:897 only touches the UNet)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0,
                 **unused):
        batch = hidden_states.shape[0]
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q = attn.to_q(hidden_states)
        k = attn.to_k(ctx)
        v = attn.to_v(ctx)
        hd = q.shape[-1] // attn.heads
        q = q.view(batch, -1, attn.heads, hd).transpose(1, 2)
        k = k.view(batch, -1, attn.heads, hd).transpose(1, 2)
        v = v.view(batch, -1, attn.heads, hd).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(batch, -1, attn.heads * hd).to(q.dtype)
        o = attn.to_out[0](o)
        o = attn.to_out[1](o)
        return o / attn.rescale_output_factor


class Attention(nn.Module):
    """diffusers.models.attention_processor.Attention with the options FMC uses: bias-free q/k/v,
    to_out = [Linear(bias), Dropout], scale = dim_head**-0.5, no group/spatial/cross norm."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, upcast_softmax=False, out_bias=True, processor=None, **unused):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.is_cross_attention = cross_attention_dim is not None
        self.upcast_attention = upcast_attention
        self.upcast_softmax = upcast_softmax
        self.rescale_output_factor = 1.0
        self.residual_connection = False
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.group_norm = None
        self.spatial_norm = None
        self.norm_cross = None
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])
        self.set_processor(processor if processor is not None else AttnProcessorSDPA())

    def set_processor(self, processor):
        # an nn.Module processor is (re)registered as the submodule `processor` so its weights appear in
        # state_dict() as `...attn1.processor.to_q_lora.down.weight` (SURVEY.md 8b key contract)
        if hasattr(self, "processor") and isinstance(self.processor, nn.Module) and not isinstance(processor, nn.Module):
            self._modules.pop("processor")
        self.processor = processor

    def head_to_batch_dim(self, t):
        b, s, d = t.shape
        h = self.heads
        return t.reshape(b, s, h, d // h).permute(0, 2, 1, 3).reshape(b * h, s, d // h)

    def batch_to_head_dim(self, t):
        bh, s, d = t.shape
        h = self.heads
        return t.reshape(bh // h, h, s, d).permute(0, 2, 1, 3).reshape(bh // h, s, d * h)

    def prepare_attention_mask(self, attention_mask, target_length, batch_size):
        if attention_mask is None:
            return None
        raise NotImplementedError("attention masks are never passed on the FMC hot path")

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        if self.upcast_attention:
            query, key = query.float(), key.float()
        if attention_mask is None:
            base = torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype, device=query.device)
            beta = 0
        else:
            base, beta = attention_mask, 1
        scores = torch.baddbmm(base, query, key.transpose(-1, -2), beta=beta, alpha=self.scale)
        if self.upcast_softmax:
            scores = scores.float()
        return scores.softmax(dim=-1).to(dtype)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        value, gate = self.proj(x).chunk(2, dim=-1)
        return value * F.gelu(gate)  # exact (erf) GELU


class FeedForward(nn.Module):
    """diffusers FeedForward(activation_fn='geglu'): net = [GEGLU(dim, 4dim), Dropout, Linear(4dim, dim)]."""

    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu"):
        super().__init__()
        assert activation_fn == "geglu"
        inner = dim * mult
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim_out or dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    """h += attn1(LN1 h); h += attn2(LN2 h, text); h += ff(LN3 h); cross_attention_kwargs splatted into both."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, cross_attention_dim=None, upcast_attention=False):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads=num_attention_heads, dim_head=attention_head_dim,
                               upcast_attention=upcast_attention)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_attention_dim=cross_attention_dim, heads=num_attention_heads,
                               dim_head=attention_head_dim, upcast_attention=upcast_attention)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, hidden_states, encoder_hidden_states=None, cross_attention_kwargs=None):
        kw = cross_attention_kwargs if cross_attention_kwargs is not None else {}
        hidden_states = self.attn1(self.norm1(hidden_states), encoder_hidden_states=None, **kw) + hidden_states
        hidden_states = self.attn2(self.norm2(hidden_states), encoder_hidden_states=encoder_hidden_states,
                                   **kw) + hidden_states
        hidden_states = self.ff(self.norm3(hidden_states)) + hidden_states
        return hidden_states


class Transformer2DModel(nn.Module):
    """GN32(eps 1e-6) -> conv1x1 -> [b,c,h,w]->[b,hw,c] -> blocks -> back -> conv1x1 -> + input."""

    def __init__(self, num_attention_heads, attention_head_dim, in_channels, num_layers=1, cross_attention_dim=None,
                 norm_num_groups=32, use_linear_projection=False, only_cross_attention=False, upcast_attention=False):
        super().__init__()
        assert not use_linear_projection and not only_cross_attention
        inner = num_attention_heads * attention_head_dim
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, kernel_size=1)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, num_attention_heads, attention_head_dim,
                                  cross_attention_dim=cross_attention_dim, upcast_attention=upcast_attention)
            for _ in range(num_layers)])
        self.proj_out = nn.Conv2d(inner, in_channels, kernel_size=1)

    def forward(self, hidden_states, encoder_hidden_states=None, cross_attention_kwargs=None, return_dict=True):
        b, c, h, w = hidden_states.shape
        residual = hidden_states
        x = self.proj_in(self.norm(hidden_states))
        inner = x.shape[1]
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, inner)
        for blk in self.transformer_blocks:
            x = blk(x, encoder_hidden_states=encoder_hidden_states, cross_attention_kwargs=cross_attention_kwargs)
        x = x.reshape(b, h, w, inner).permute(0, 3, 1, 2).contiguous()
        out = self.proj_out(x) + residual
        return SimpleNamespace(sample=out) if return_dict else (out,)


class ResnetBlock2D(nn.Module):
    """h = conv1(silu(GN x)); h += Linear(silu(temb))[:, :, None, None]; h = conv2(silu(GN h));
    x = conv_shortcut(x) if cin != cout; (x + h) / output_scale_factor."""

    def __init__(self, in_channels, out_channels=None, temb_channels=512, groups=32, eps=1e-6, dropout=0.0,
                 output_scale_factor=1.0, **unused):
        super().__init__()
        out_channels = out_channels or in_channels
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.output_scale_factor = output_scale_factor
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(self.dropout(F.silu(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return (x + h) / self.output_scale_factor


class Downsample2D(nn.Module):
    """use_conv=True, name='op' -> attribute `conv`: conv3x3 stride 2 pad `padding`."""

    def __init__(self, channels, use_conv=True, out_channels=None, padding=1, name="conv"):
        super().__init__()
        assert use_conv
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, stride=2, padding=padding)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels, use_conv=True, out_channels=None):
        super().__init__()
        assert use_conv
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, padding=1)

    def forward(self, x, output_size=None):
        if output_size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=output_size, mode="nearest")
        return self.conv(x)


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
        exponent = exponent / (half - self.downscale_freq_shift)
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        if self.flip_sin_to_cos:
            emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
        return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


class DDIMScheduler:
    """diffusers DDIMScheduler with the FMC kwargs (configs/cam.yaml:130-136): 1000 train steps, linear betas
    0.00085 -> 0.012, steps_offset 1, clip_sample False; defaults set_alpha_to_one=True, epsilon prediction,
    'leading' spacing."""

    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="linear",
                 steps_offset=1, clip_sample=False, set_alpha_to_one=True):
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise ValueError(beta_schedule)
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.clip_sample = clip_sample
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (torch.arange(num_inference_steps) * ratio).round().flip(0).to(torch.int64) + self.steps_offset
        self.timesteps = ts.to(device) if device is not None else ts

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, eta=0.0, **unused):
        assert eta == 0.0
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t].to(sample.device, sample.dtype)
        a_prev = (self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod).to(sample.device,
                                                                                               sample.dtype)
        pred_x0 = (sample - (1 - a_t) ** 0.5 * model_output) / a_t ** 0.5
        direction = (1 - a_prev) ** 0.5 * model_output
        return SimpleNamespace(prev_sample=a_prev ** 0.5 * pred_x0 + direction, pred_original_sample=pred_x0)

    def add_noise(self, original, noise, timesteps):
        a = self.alphas_cumprod.to(original.device, original.dtype)[timesteps]
        while a.ndim < original.ndim:
            a = a.unsqueeze(-1)
        return a ** 0.5 * original + (1 - a) ** 0.5 * noise
