"""ORACLE (test infrastructure, CPU fp32; parity unpinned -- see oracle/diffusers_restated.py).

Restates fmc/models/motion_module.py: the AnimateDiff temporal module carrying the CameraAdapter.
  VanillaTemporalModule       :44-90
  TemporalTransformer3DModel  :93-234  (forward :210-234)
  TemporalTransformerBlock    :237-300 (forward :287-300)
  PositionalEncoding          :303-321
  TemporalSelfAttention       :324-389
and InflatedGroupNorm / InflatedConv3d from fmc/models/resnet.py:16-37.
The six causal-mask variants (:155-208) are unused by the shipped configs and are not restated.
"""
import math

import torch
from torch import nn

from .attention_processor import PoseAdaptorAttnProcessor
from .diffusers_restated import Attention, FeedForward


class InflatedConv3d(nn.Conv2d):
    """Per-frame 2-D conv on [b, c, f, h, w] (resnet.py:16-24)."""

    def forward(self, x):
        b, c, f, h, w = x.shape
        y = super().forward(x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w))
        return y.reshape(b, f, *y.shape[1:]).permute(0, 2, 1, 3, 4)


class InflatedGroupNorm(nn.GroupNorm):
    """Per-frame GroupNorm on [b, c, f, h, w] (resnet.py:27-37)."""

    def forward(self, x):
        b, c, f, h, w = x.shape
        y = super().forward(x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w))
        return y.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


class PositionalEncoding(nn.Module):
    """pe[p, 2i] = sin(p * exp(-2i ln(1e4)/d)), pe[p, 2i+1] = cos(same); x + pe[:, :len] (:303-321)."""

    def __init__(self, d_model, dropout=0.0, max_len=32):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(1, max_len, d_model)
        pe[0, :, 0::2] = torch.sin(position * div_term)
        pe[0, :, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)

    def forward(self, x):
        return self.dropout(x + self.pe[:, : x.size(1)])


class TemporalSelfAttention(Attention):
    def __init__(self, attention_mode=None, temporal_position_encoding=False, temporal_position_encoding_max_len=32,
                 rescale_output_factor=1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert attention_mode == "Temporal_Self"
        self.pos_encoder = PositionalEncoding(kwargs["query_dim"], max_len=temporal_position_encoding_max_len) \
            if temporal_position_encoding else None
        self.rescale_output_factor = rescale_output_factor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        """(:349-389) PE is added to the (already layer-normed) input; a 5-D pose feature is brought to the temporal
        token layout '(b h w) f c'; the caller's encoder_hidden_states is dropped (always self-attention)."""
        if self.pos_encoder is not None:
            hidden_states = self.pos_encoder(hidden_states)
        if "pose_feature" in cross_attention_kwargs:
            pose = cross_attention_kwargs["pose_feature"]
            if pose.ndim == 5:
                b, c, f, h, w = pose.shape
                pose = pose.permute(0, 3, 4, 2, 1).reshape(b * h * w, f, c)
            else:
                assert pose.ndim == 3
            cross_attention_kwargs["pose_feature"] = pose
        if isinstance(self.processor, PoseAdaptorAttnProcessor):
            return self.processor(self, hidden_states, cross_attention_kwargs.pop("pose_feature"),
                                  encoder_hidden_states=None, attention_mask=attention_mask, **cross_attention_kwargs)
        return self.processor(self, hidden_states, encoder_hidden_states=None, attention_mask=attention_mask,
                              **cross_attention_kwargs)


class TemporalTransformerBlock(nn.Module):
    def __init__(self, dim, num_attention_heads, attention_head_dim, attention_block_types=("Temporal_Self", "Temporal_Self"),
                 dropout=0.0, norm_num_groups=32, cross_attention_dim=768, activation_fn="geglu", attention_bias=False,
                 upcast_attention=False, temporal_position_encoding=False, temporal_position_encoding_max_len=32,
                 encoder_hidden_states_query=(False, False), attention_activation_scale=1.0,
                 attention_processor_kwargs=None, rescale_output_factor=1.0):
        super().__init__()
        self.attention_block_types = attention_block_types
        self.attention_blocks = nn.ModuleList([
            TemporalSelfAttention(attention_mode=name, cross_attention_dim=None, query_dim=dim, heads=num_attention_heads,
                                  dim_head=attention_head_dim, dropout=dropout, bias=attention_bias,
                                  upcast_attention=upcast_attention,
                                  temporal_position_encoding=temporal_position_encoding,
                                  temporal_position_encoding_max_len=temporal_position_encoding_max_len,
                                  rescale_output_factor=rescale_output_factor)
            for name in attention_block_types])
        self.norms = nn.ModuleList([nn.LayerNorm(dim) for _ in attention_block_types])
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn)
        self.ff_norm = nn.LayerNorm(dim)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, cross_attention_kwargs=None):
        kw = cross_attention_kwargs if cross_attention_kwargs is not None else {}
        for attention_block, norm in zip(self.attention_blocks, self.norms):
            normed = norm(hidden_states)
            # the kwargs dict is re-splatted for every block, so block 1 sees the original 5-D pose feature again
            hidden_states = attention_block(normed, encoder_hidden_states=normed, attention_mask=attention_mask,
                                            **kw) + hidden_states
        return self.ff(self.ff_norm(hidden_states)) + hidden_states


class TemporalTransformer3DModel(nn.Module):
    def __init__(self, in_channels, num_attention_heads, attention_head_dim, num_layers,
                 attention_block_types=("Temporal_Self", "Temporal_Self"), dropout=0.0, norm_num_groups=32,
                 cross_attention_dim=320, activation_fn="geglu", attention_bias=False, upcast_attention=False,
                 temporal_position_encoding=False, temporal_position_encoding_max_len=32,
                 encoder_hidden_states_query=(False, False), attention_activation_scale=1.0,
                 attention_processor_kwargs=None, causal_temporal_attention=None,
                 causal_temporal_attention_mask_type="", rescale_output_factor=1.0):
        super().__init__()
        assert causal_temporal_attention is not None and not causal_temporal_attention
        inner_dim = num_attention_heads * attention_head_dim
        self.norm = InflatedGroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner_dim)
        self.transformer_blocks = nn.ModuleList([
            TemporalTransformerBlock(dim=inner_dim, num_attention_heads=num_attention_heads,
                                     attention_head_dim=attention_head_dim, attention_block_types=attention_block_types,
                                     dropout=dropout, norm_num_groups=norm_num_groups,
                                     cross_attention_dim=cross_attention_dim, activation_fn=activation_fn,
                                     attention_bias=attention_bias, upcast_attention=upcast_attention,
                                     temporal_position_encoding=temporal_position_encoding,
                                     temporal_position_encoding_max_len=temporal_position_encoding_max_len,
                                     rescale_output_factor=rescale_output_factor)
            for _ in range(num_layers)])
        self.proj_out = nn.Linear(inner_dim, in_channels)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, cross_attention_kwargs=None):
        assert hidden_states.dim() == 5
        residual = hidden_states
        b, c, f, h, w = hidden_states.shape
        x = self.norm(hidden_states)
        x = x.permute(0, 3, 4, 2, 1).reshape(b * h * w, f, c)  # 'b c f h w -> (b h w) f c'
        x = self.proj_in(x)
        for block in self.transformer_blocks:
            x = block(x, encoder_hidden_states=encoder_hidden_states, attention_mask=attention_mask,
                      cross_attention_kwargs=cross_attention_kwargs)
        x = self.proj_out(x)
        x = x.reshape(b, h, w, f, c).permute(0, 4, 3, 1, 2)  # '(b h w) f c -> b c f h w'
        return x + residual


class VanillaTemporalModule(nn.Module):
    def __init__(self, in_channels, num_attention_heads=8, num_transformer_block=2,
                 attention_block_types=("Temporal_Self",), temporal_position_encoding=True,
                 temporal_position_encoding_max_len=32, temporal_attention_dim_div=1, cross_attention_dim=320,
                 zero_initialize=True, encoder_hidden_states_query=(False, False), attention_activation_scale=1.0,
                 attention_processor_kwargs=None, causal_temporal_attention=False,
                 causal_temporal_attention_mask_type="", rescale_output_factor=1.0):
        super().__init__()
        self.temporal_transformer = TemporalTransformer3DModel(
            in_channels=in_channels, num_attention_heads=num_attention_heads,
            attention_head_dim=in_channels // num_attention_heads // temporal_attention_dim_div,
            num_layers=num_transformer_block, attention_block_types=tuple(attention_block_types),
            cross_attention_dim=cross_attention_dim, temporal_position_encoding=temporal_position_encoding,
            temporal_position_encoding_max_len=temporal_position_encoding_max_len,
            causal_temporal_attention=causal_temporal_attention,
            causal_temporal_attention_mask_type=causal_temporal_attention_mask_type,
            rescale_output_factor=rescale_output_factor)
        if zero_initialize:
            self.temporal_transformer.proj_out = zero_module(self.temporal_transformer.proj_out)

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs=None):
        return self.temporal_transformer(hidden_states, encoder_hidden_states, attention_mask,
                                         cross_attention_kwargs=cross_attention_kwargs)


def get_motion_module(in_channels, motion_module_type, motion_module_kwargs):
    if motion_module_type == "Vanilla":
        return VanillaTemporalModule(in_channels=in_channels, **motion_module_kwargs)
    raise ValueError(motion_module_type)
