"""Mirror of fmc/models/motion_module.py: the AnimateDiff temporal module that carries the CameraAdapter.

Same classes / state-dict keys as the reference (VanillaTemporalModule :44, TemporalTransformer3DModel :93,
TemporalTransformerBlock :237, PositionalEncoding :303, TemporalSelfAttention :324).  forward() runs on channels-last
activations with the kernels of libfmc_b200:
    GN -> proj_in GEMM -> [LN+PE(+pose) -> (qkv_merge GEMM) -> fused qkv GEMM -> temporal attention -> out GEMM+res] x2
       -> LN -> GEGLU GEMM -> GEMM+res -> proj_out GEMM + input
"""
import math

import torch
from torch import nn

from ... import engine, ops
from ...engine import CL
from .._blocks import Attention, FeedForward
from .resnet import InflatedGroupNorm, zero_module


class PositionalEncoding(nn.Module):
    def __init__(self, d_model, dropout=0.0, max_len=32):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(1, max_len, d_model)
        pe[0, :, 0::2] = torch.sin(position * div_term)
        pe[0, :, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)


class TemporalSelfAttention(Attention):
    def __init__(self, attention_mode=None, temporal_position_encoding=False, temporal_position_encoding_max_len=32,
                 rescale_output_factor=1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert attention_mode == "Temporal_Self"
        self.pos_encoder = PositionalEncoding(kwargs["query_dim"], max_len=temporal_position_encoding_max_len) \
            if temporal_position_encoding else None
        self.rescale_output_factor = rescale_output_factor

    def set_use_memory_efficient_attention_xformers(self, *args, **kwargs):
        pass


class TemporalTransformerBlock(nn.Module):
    def __init__(self, dim, num_attention_heads, attention_head_dim, attention_block_types=("Temporal_Self", "Temporal_Self"),
                 dropout=0.0, norm_num_groups=32, cross_attention_dim=768, activation_fn="geglu", attention_bias=False,
                 upcast_attention=False, temporal_position_encoding=False, temporal_position_encoding_max_len=32,
                 encoder_hidden_states_query=(False, False), attention_activation_scale=1.0,
                 attention_processor_kwargs=None, rescale_output_factor=1.0):
        super().__init__()
        self.attention_block_types = tuple(attention_block_types)
        self.attention_blocks = nn.ModuleList([
            TemporalSelfAttention(attention_mode=name, cross_attention_dim=None, query_dim=dim, heads=num_attention_heads,
                                  dim_head=attention_head_dim, dropout=dropout, bias=attention_bias,
                                  temporal_position_encoding=temporal_position_encoding,
                                  temporal_position_encoding_max_len=temporal_position_encoding_max_len,
                                  rescale_output_factor=rescale_output_factor)
            for name in attention_block_types])
        self.norms = nn.ModuleList([nn.LayerNorm(dim) for _ in attention_block_types])
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn)
        self.ff_norm = nn.LayerNorm(dim)
        self._plan = None

    def plan(self, device):
        if self._plan is None or self._plan["device"] != engine.plan_key(device):
            p = {"device": engine.plan_key(device), "attn": [], "norms": [], "pe": []}
            for attn, norm in zip(self.attention_blocks, self.norms):
                p["attn"].append(engine.AttnPlan(attn, device, fused_temporal=True))
                p["norms"].append(engine.NormPlan(norm, device))
                p["pe"].append(attn.pos_encoder.pe[0].detach().to(device=device, dtype=torch.float32).contiguous()
                               if attn.pos_encoder is not None else None)
            p["ff_norm"] = engine.NormPlan(self.ff_norm, device)
            p["ff1"] = engine.LinearPlan(self.ff.net[0].proj.weight.detach().float(),
                                         self.ff.net[0].proj.bias.detach().float(), device, geglu=True,
                                         pre_norm=self.ff_norm if engine.ln_fold_wanted(*self.ff.net[0].proj.weight.shape) else None)
            p["ff2"] = engine.LinearPlan(self.ff.net[2].weight.detach().float(), self.ff.net[2].bias.detach().float(), device)
            self._plan = p
        return self._plan

    def run(self, h, B, F, HW, pose_rows=None):
        """h: rows [(B F HW), C] channels-last; pose_rows: same shape or None.  Returns new rows."""
        p = self.plan(h.device)
        for ap, npl, pe in zip(p["attn"], p["norms"], p["pe"]):
            if pe is not None and F > pe.shape[0]:
                raise ValueError(f"{F} frames exceed the positional-encoding length {pe.shape[0]} "
                                 "(motion_module.py:320; SURVEY H5)")
            if ap.merge is not None:
                if pose_rows is None:
                    raise ValueError("PoseAdaptorAttnProcessor needs a pose_feature (attention_processor.py:210)")
                x, xp = ops.layernorm(h, npl.g, npl.b, npl.eps, pe=pe, F=F, HW=HW, add=pose_rows)
            else:
                x, xp = ops.layernorm(h, npl.g, npl.b, npl.eps, pe=pe, F=F, HW=HW), None
            h = engine.run_temporal_attention(ap, x, xp, h, B, F, HW)
        if p["ff1"].colsum is not None:  # ff_norm folded into the GEGLU GEMM
            return p["ff2"](p["ff1"](h, ln_stats=ops.rowstats(h, p["ff_norm"].eps)), residual=h)
        n = ops.layernorm(h, p["ff_norm"].g, p["ff_norm"].b, p["ff_norm"].eps)
        return p["ff2"](p["ff1"](n), residual=h)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, cross_attention_kwargs=None):
        raise RuntimeError("TemporalTransformerBlock runs through .run() on channels-last rows (no eager fallback)")


class TemporalTransformer3DModel(nn.Module):
    def __init__(self, in_channels, num_attention_heads, attention_head_dim, num_layers,
                 attention_block_types=("Temporal_Self", "Temporal_Self"), dropout=0.0, norm_num_groups=32,
                 cross_attention_dim=320, activation_fn="geglu", attention_bias=False, upcast_attention=False,
                 temporal_position_encoding=False, temporal_position_encoding_max_len=32,
                 encoder_hidden_states_query=(False, False), attention_activation_scale=1.0,
                 attention_processor_kwargs=None, causal_temporal_attention=None,
                 causal_temporal_attention_mask_type="", rescale_output_factor=1.0):
        super().__init__()
        assert causal_temporal_attention is not None and not causal_temporal_attention, \
            "causal temporal masks are unused by the shipped configs and not implemented"
        inner_dim = num_attention_heads * attention_head_dim
        self.norm = InflatedGroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner_dim)
        self.transformer_blocks = nn.ModuleList([
            TemporalTransformerBlock(dim=inner_dim, num_attention_heads=num_attention_heads,
                                     attention_head_dim=attention_head_dim, attention_block_types=attention_block_types,
                                     dropout=dropout, norm_num_groups=norm_num_groups,
                                     cross_attention_dim=cross_attention_dim, activation_fn=activation_fn,
                                     attention_bias=attention_bias,
                                     temporal_position_encoding=temporal_position_encoding,
                                     temporal_position_encoding_max_len=temporal_position_encoding_max_len,
                                     rescale_output_factor=rescale_output_factor)
            for _ in range(num_layers)])
        self.proj_out = nn.Linear(inner_dim, in_channels)
        self._plan = None

    def plan(self, device):
        if self._plan is None or self._plan["device"] != engine.plan_key(device):
            self._plan = {
                "device": engine.plan_key(device),
                "norm": engine.NormPlan(self.norm, device),
                "proj_in": engine.LinearPlan(self.proj_in.weight.detach().float(), self.proj_in.bias.detach().float(), device),
                "proj_out": engine.LinearPlan(self.proj_out.weight.detach().float(), self.proj_out.bias.detach().float(), device),
            }
        return self._plan

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, cross_attention_kwargs=None):
        assert attention_mask is None
        x = CL.from_reference(hidden_states)
        B, F, H, W, C = x.dims
        HW = H * W
        rows = x.rows()
        p = self.plan(rows.device)
        pose_rows = None
        if cross_attention_kwargs and cross_attention_kwargs.get("pose_feature") is not None:
            pose = engine.as_cl_feature(cross_attention_kwargs["pose_feature"])
            assert pose.dims == x.dims, (pose.dims, x.dims)
            pose_rows = pose.rows()
        n = ops.groupnorm(rows, p["norm"].g, p["norm"].b, p["norm"].eps, B * F, HW, groups=p["norm"].groups)
        h = p["proj_in"](n)
        for block in self.transformer_blocks:
            h = block.run(h, B, F, HW, pose_rows)
        out = p["proj_out"](h, residual=rows)
        return CL(out.view(B, F, H, W, C))


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
