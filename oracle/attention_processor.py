"""ORACLE (test infrastructure, CPU fp32; parity unpinned -- see oracle/diffusers_restated.py).

Restates fmc/models/attention_processor.py: the four attention processors of the FMC U-Net.
  AttnProcessor                 :15-82   plain attention (temporal attention_blocks.1)
  LoRAAttnProcessor             :85-169  Domain-LoRA on q/k/v/out (every spatial attn1/attn2)
  PoseAdaptorAttnProcessor      :172-293 CameraAdapter: m = qkv_merge(x + pose) * scale + x (temporal attention_blocks.0)
  LORAPoseAdaptorAttnProcessor  :296-420 both
All four share one attention core: baddbmm(beta=0, alpha=scale) -> softmax -> bmm (:61-67).
"""
import torch
from torch import nn

from .diffusers_restated import LoRALinearLayer


def _to_tokens(t):
    """'b c h w -> b (h w) c' for 4-D inputs (:222-223); 3-D passes through.  (The 5-D branch at :220 tests a bound
    method against an int and never fires -- SURVEY Appendix B.1.)"""
    if t.ndim == 4:
        b, c, h, w = t.shape
        return t.reshape(b, c, h * w).transpose(1, 2)
    assert t.ndim == 3
    return t


def _attend(attn, q_in, kv_in, q_extra=None, k_extra=None, v_extra=None, out_extra=None, attention_mask=None):
    """q/k/v projections (+ optional additive LoRA branches) -> softmax(q k^T * scale) v -> to_out (+ LoRA)."""
    query = attn.to_q(q_in)
    key = attn.to_k(kv_in)
    value = attn.to_v(kv_in)
    if q_extra is not None:
        query, key, value = query + q_extra(q_in), key + k_extra(kv_in), value + v_extra(kv_in)
    query = attn.head_to_batch_dim(query)
    key = attn.head_to_batch_dim(key)
    value = attn.head_to_batch_dim(value)
    probs = attn.get_attention_scores(query, key, attention_mask)
    ctx = attn.batch_to_head_dim(torch.bmm(probs, value))
    out = attn.to_out[0](ctx)
    if out_extra is not None:
        out = out + out_extra(ctx)
    return attn.to_out[1](out)


class AttnProcessor:
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0,
                 pose_feature=None):
        residual = hidden_states
        x = hidden_states
        shape4 = x.shape if x.ndim == 4 else None
        x = _to_tokens(x)
        ctx = x if encoder_hidden_states is None else encoder_hidden_states
        out = _attend(attn, x, ctx, attention_mask=attention_mask)
        if shape4 is not None:
            out = out.transpose(-1, -2).reshape(shape4)
        if attn.residual_connection:
            out = out + residual
        return out / attn.rescale_output_factor


class LoRAAttnProcessor(nn.Module):
    def __init__(self, hidden_size=None, cross_attention_dim=None, rank=4, network_alpha=None, lora_scale=1.0):
        super().__init__()
        self.rank = rank
        self.lora_scale = lora_scale
        kv_dim = cross_attention_dim or hidden_size
        self.to_q_lora = LoRALinearLayer(hidden_size, hidden_size, rank, network_alpha)
        self.to_k_lora = LoRALinearLayer(kv_dim, hidden_size, rank, network_alpha)
        self.to_v_lora = LoRALinearLayer(kv_dim, hidden_size, rank, network_alpha)
        self.to_out_lora = LoRALinearLayer(hidden_size, hidden_size, rank, network_alpha)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 pose_feature=None, scale=None):
        s = self.lora_scale if scale is None else scale
        residual = hidden_states
        x = hidden_states
        shape4 = x.shape if x.ndim == 4 else None
        x = _to_tokens(x)
        ctx = x if encoder_hidden_states is None else encoder_hidden_states
        out = _attend(attn, x, ctx,
                      q_extra=lambda t: s * self.to_q_lora(t), k_extra=lambda t: s * self.to_k_lora(t),
                      v_extra=lambda t: s * self.to_v_lora(t), out_extra=lambda t: s * self.to_out_lora(t),
                      attention_mask=attention_mask)
        if shape4 is not None:
            out = out.transpose(-1, -2).reshape(shape4)
        if attn.residual_connection:
            out = out + residual
        return out / attn.rescale_output_factor


class _PoseMergeMixin:
    def _build_merge(self, hidden_size, pose_feature_dim, query_condition, key_value_condition):
        assert hidden_size == pose_feature_dim
        self.query_condition = query_condition
        self.key_value_condition = key_value_condition
        name = "qkv_merge" if (query_condition and key_value_condition) else ("q_merge" if query_condition else "kv_merge")
        layer = nn.Linear(hidden_size, hidden_size)
        nn.init.zeros_(layer.weight)
        nn.init.zeros_(layer.bias)
        setattr(self, name, layer)

    def _merge(self, x, ctx, pose, s):
        """(:255-264) returns (query source, key/value source)."""
        if self.query_condition and self.key_value_condition:
            m = self.qkv_merge(x + pose) * s + x
            return m, m
        if self.query_condition:
            return self.q_merge(x + pose) * s + x, ctx
        return x, self.kv_merge(ctx + pose) * s + ctx


class PoseAdaptorAttnProcessor(nn.Module, _PoseMergeMixin):
    def __init__(self, hidden_size, pose_feature_dim=None, cross_attention_dim=None, query_condition=False,
                 key_value_condition=False, scale=1.0):
        super().__init__()
        self.hidden_size = hidden_size
        self.pose_feature_dim = pose_feature_dim
        self.cross_attention_dim = cross_attention_dim
        self.scale = scale
        self._build_merge(hidden_size, pose_feature_dim, query_condition, key_value_condition)

    def forward(self, attn, hidden_states, pose_feature, encoder_hidden_states=None, attention_mask=None, temb=None,
                scale=None):
        assert pose_feature is not None
        s = scale or self.scale
        residual = hidden_states
        x = _to_tokens(hidden_states)
        if self.query_condition and self.key_value_condition:
            assert encoder_hidden_states is None
        ctx = _to_tokens(x if encoder_hidden_states is None else encoder_hidden_states)
        pose = _to_tokens(pose_feature)
        q_src, kv_src = self._merge(x, ctx, pose, s)
        out = _attend(attn, q_src, kv_src, attention_mask=attention_mask)
        if attn.residual_connection:
            out = out + residual
        return out / attn.rescale_output_factor


class LORAPoseAdaptorAttnProcessor(nn.Module, _PoseMergeMixin):
    def __init__(self, hidden_size, pose_feature_dim=None, cross_attention_dim=None, query_condition=False,
                 key_value_condition=False, scale=1.0, rank=4, network_alpha=None, lora_scale=1.0):
        super().__init__()
        self.hidden_size = hidden_size
        self.pose_feature_dim = pose_feature_dim
        self.cross_attention_dim = cross_attention_dim
        self.scale = scale
        self._build_merge(hidden_size, pose_feature_dim, query_condition, key_value_condition)
        self.rank = rank
        self.lora_scale = lora_scale
        kv_dim = cross_attention_dim or hidden_size
        self.to_q_lora = LoRALinearLayer(hidden_size, hidden_size, rank, network_alpha)
        self.to_k_lora = LoRALinearLayer(kv_dim, hidden_size, rank, network_alpha)
        self.to_v_lora = LoRALinearLayer(kv_dim, hidden_size, rank, network_alpha)
        self.to_out_lora = LoRALinearLayer(hidden_size, hidden_size, rank, network_alpha)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0,
                 pose_feature=None):
        assert pose_feature is not None
        # here the call-time `scale` is the LoRA scale and the pose scale is always self.scale (:347,:388)
        ls = self.lora_scale if scale is None else scale
        residual = hidden_states
        x = _to_tokens(hidden_states)
        if self.query_condition and self.key_value_condition:
            assert encoder_hidden_states is None
        ctx = _to_tokens(x if encoder_hidden_states is None else encoder_hidden_states)
        pose = _to_tokens(pose_feature)
        q_src, kv_src = self._merge(x, ctx, pose, self.scale)
        out = _attend(attn, q_src, kv_src,
                      q_extra=lambda t: ls * self.to_q_lora(t), k_extra=lambda t: ls * self.to_k_lora(t),
                      v_extra=lambda t: ls * self.to_v_lora(t), out_extra=lambda t: ls * self.to_out_lora(t),
                      attention_mask=attention_mask)
        if attn.residual_connection:
            out = out + residual
        return out / attn.rescale_output_factor
