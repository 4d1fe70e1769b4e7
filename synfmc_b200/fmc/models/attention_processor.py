"""Mirror of fmc/models/attention_processor.py: the four attention processors, as parameter / configuration holders.
The engine (synfmc_b200.engine.AttnPlan) reads them: Domain-LoRA (`to_{q,k,v,out}_lora`, :103-106) is folded into the
projection weights, CameraAdapter `qkv_merge` (:189-192) becomes one GEMM with the scale folded in, and the attention
core runs in fmc_spatial_attn_bf16 / fmc_temporal_attn_bf16."""
from torch import nn

from .._blocks import LoRALinearLayer


def _not_callable(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__} is executed by the synfmc_b200 engine, not called (no eager fallback)")


class AttnProcessor:
    """Plain attention (:15-82).  The trainers only use it in an isinstance() filter (train_cam_ctrl.py:263-266)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0,
                 pose_feature=None):
        return _not_callable(self, attn, hidden_states)


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

    forward = _not_callable


def _build_merge(self, hidden_size, pose_feature_dim, query_condition, key_value_condition):
    assert hidden_size == pose_feature_dim
    self.query_condition = query_condition
    self.key_value_condition = key_value_condition
    name = "qkv_merge" if (query_condition and key_value_condition) else ("q_merge" if query_condition else "kv_merge")
    layer = nn.Linear(hidden_size, hidden_size)
    nn.init.zeros_(layer.weight)
    nn.init.zeros_(layer.bias)
    setattr(self, name, layer)


class PoseAdaptorAttnProcessor(nn.Module):
    def __init__(self, hidden_size, pose_feature_dim=None, cross_attention_dim=None, query_condition=False,
                 key_value_condition=False, scale=1.0):
        super().__init__()
        self.hidden_size = hidden_size
        self.pose_feature_dim = pose_feature_dim
        self.cross_attention_dim = cross_attention_dim
        self.scale = scale
        _build_merge(self, hidden_size, pose_feature_dim, query_condition, key_value_condition)

    forward = _not_callable


class LORAPoseAdaptorAttnProcessor(nn.Module):
    def __init__(self, hidden_size, pose_feature_dim=None, cross_attention_dim=None, query_condition=False,
                 key_value_condition=False, scale=1.0, rank=4, network_alpha=None, lora_scale=1.0):
        super().__init__()
        self.hidden_size = hidden_size
        self.pose_feature_dim = pose_feature_dim
        self.cross_attention_dim = cross_attention_dim
        self.scale = scale
        _build_merge(self, hidden_size, pose_feature_dim, query_condition, key_value_condition)
        self.rank = rank
        self.lora_scale = lora_scale
        kv_dim = cross_attention_dim or hidden_size
        self.to_q_lora = LoRALinearLayer(hidden_size, hidden_size, rank, network_alpha)
        self.to_k_lora = LoRALinearLayer(kv_dim, hidden_size, rank, network_alpha)
        self.to_v_lora = LoRALinearLayer(kv_dim, hidden_size, rank, network_alpha)
        self.to_out_lora = LoRALinearLayer(hidden_size, hidden_size, rank, network_alpha)

    forward = _not_callable
