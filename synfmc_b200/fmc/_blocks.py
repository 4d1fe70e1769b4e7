"""Parameter holders with diffusers==0.24.0 names (the reference imports these from diffusers; it does not vendor
them).  They own the fp32 parameters / state-dict keys; the arithmetic is executed by synfmc_b200.engine on the GPU.
Calling them like ordinary torch modules is not supported: there is no eager fallback."""
import torch
from torch import nn


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(f"{type(self).__name__} is a parameter holder; it is executed by the synfmc_b200 engine "
                           "(no eager PyTorch fallback)")


class LoRALinearLayer(_Holder):
    def __init__(self, in_features, out_features, rank=4, network_alpha=None):
        super().__init__()
        self.down = nn.Linear(in_features, rank, bias=False)
        self.up = nn.Linear(rank, out_features, bias=False)
        self.network_alpha = network_alpha
        self.rank = rank
        nn.init.normal_(self.down.weight, std=1 / rank)
        nn.init.zeros_(self.up.weight)


class Attention(_Holder):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, out_bias=True, processor=None, **unused):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.is_cross_attention = cross_attention_dim is not None
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.rescale_output_factor = 1.0
        self.residual_connection = False
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])
        from .models.attention_processor import AttnProcessor
        self.set_processor(processor if processor is not None else AttnProcessor())

    def set_processor(self, processor):
        if hasattr(self, "processor") and isinstance(self.processor, nn.Module) and not isinstance(processor, nn.Module):
            self._modules.pop("processor")
        self.processor = processor
        from .. import engine
        engine.structure_changed()  # cached processor lists of engine.fingerprint are rebuilt


class GEGLU(_Holder):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(_Holder):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu"):
        super().__init__()
        assert activation_fn == "geglu"
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(dropout), nn.Linear(dim * mult, dim_out or dim)])


class BasicTransformerBlock(_Holder):
    def __init__(self, dim, num_attention_heads, attention_head_dim, cross_attention_dim=None):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads=num_attention_heads, dim_head=attention_head_dim)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_attention_dim=cross_attention_dim, heads=num_attention_heads,
                               dim_head=attention_head_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)


class Transformer2DModel(_Holder):
    def __init__(self, num_attention_heads, attention_head_dim, in_channels, num_layers=1, cross_attention_dim=None,
                 norm_num_groups=32, use_linear_projection=False, only_cross_attention=False, upcast_attention=False):
        super().__init__()
        assert not use_linear_projection and not only_cross_attention
        inner = num_attention_heads * attention_head_dim
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, kernel_size=1)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, cross_attention_dim=cross_attention_dim)
            for _ in range(num_layers)])
        self.proj_out = nn.Conv2d(inner, in_channels, kernel_size=1)


class ResnetBlock2D(_Holder):
    def __init__(self, in_channels, out_channels=None, temb_channels=512, groups=32, eps=1e-6, dropout=0.0,
                 output_scale_factor=1.0, **unused):
        super().__init__()
        out_channels = out_channels or in_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.output_scale_factor = output_scale_factor
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None


class Downsample2D(_Holder):
    def __init__(self, channels, use_conv=True, out_channels=None, padding=1, name="conv"):
        super().__init__()
        assert use_conv
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, stride=2, padding=padding)


class Upsample2D(_Holder):
    def __init__(self, channels, use_conv=True, out_channels=None):
        super().__init__()
        assert use_conv
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, padding=1)


class Timesteps(_Holder):
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
        super().__init__()
        assert flip_sin_to_cos and downscale_freq_shift == 0, "only the SD1.5 time projection is implemented"
        self.num_channels = num_channels


class TimestepEmbedding(_Holder):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)


class DDIMScheduler:
    """Host-side schedule of diffusers DDIMScheduler with the FMC kwargs (configs/cam.yaml:130-136); the update itself
    runs in fmc_cfg_ddim_step_f32."""

    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="linear",
                 steps_offset=1, clip_sample=False, set_alpha_to_one=True, **unused):
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise ValueError(beta_schedule)
        assert not clip_sample
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = 1.0 if set_alpha_to_one else float(self.alphas_cumprod[0])
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        self.timesteps = (torch.arange(num_inference_steps) * ratio).round().flip(0).to(torch.int64) + self.steps_offset

    def scale_model_input(self, sample, timestep=None):
        return sample

    def alphas_for(self, timestep):
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_prev = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else self.final_alpha_cumprod
        return a_t, a_prev

    def step(self, model_output, timestep, sample, eta=0.0, **unused):
        from types import SimpleNamespace
        from .. import ops
        assert eta == 0.0
        a_t, a_prev = self.alphas_for(timestep)
        prev = ops.cfg_ddim_step(model_output, None, 1.0, sample, a_t, a_prev)
        return SimpleNamespace(prev_sample=prev)

    def add_noise(self, original, noise, timesteps):
        a = self.alphas_cumprod.to(original.device, original.dtype)[timesteps]
        while a.ndim < original.ndim:
            a = a.unsqueeze(-1)
        return a ** 0.5 * original + (1 - a) ** 0.5 * noise
