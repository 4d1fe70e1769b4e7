"""ORACLE (test infrastructure, CPU fp32; parity unpinned -- see oracle/diffusers_restated.py).

Restates fmc/models/unet_blocks.py: the 3-D U-Net blocks = per-frame ResnetBlock2D + Transformer2DModel
(spatial self + text cross attention) + temporal motion module.
  UNetMidBlock3DCrossAttn :144-265   CrossAttnDownBlock3D :268-426   DownBlock3D :429-540
  CrossAttnUpBlock3D      :543-706   UpBlock3D            :709-817
The reference wraps every per-frame op in 'b c f h w <-> (b f) c h w' rearranges; `_per_frame` does the same.
"""
import torch
from torch import nn

from .diffusers_restated import Downsample2D, ResnetBlock2D, Transformer2DModel, Upsample2D
from .motion_module import get_motion_module


def _per_frame(fn, x, *args, **kwargs):
    b, c, f, h, w = x.shape
    y = fn(x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w), *args, **kwargs)
    if hasattr(y, "sample"):
        y = y.sample
    return y.reshape(b, f, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def _repeat_temb(temb, f):
    return temb.repeat_interleave(f, dim=0)  # 'b c -> (b f) c'


def _resnet(cin, cout, temb_channels, eps, groups, scale=1.0):
    return ResnetBlock2D(in_channels=cin, out_channels=cout, temb_channels=temb_channels, eps=eps, groups=groups,
                         output_scale_factor=scale)


def _transformer(heads, channels, cross_attention_dim, groups, upcast_attention=False):
    # SD1.5's `attention_head_dim: 8` is used as the head COUNT (unet_blocks.py:323-326)
    return Transformer2DModel(heads, channels // heads, in_channels=channels, num_layers=1,
                              cross_attention_dim=cross_attention_dim, norm_num_groups=groups,
                              upcast_attention=upcast_attention)


def _motion(channels, use, mtype, mkwargs):
    return get_motion_module(in_channels=channels, motion_module_type=mtype, motion_module_kwargs=mkwargs) if use else None


class UNetMidBlock3DCrossAttn(nn.Module):
    def __init__(self, in_channels, temb_channels, num_layers=1, resnet_eps=1e-6, resnet_groups=32,
                 attn_num_head_channels=1, output_scale_factor=1.0, cross_attention_dim=1280, upcast_attention=False,
                 use_motion_module=None, motion_module_type=None, motion_module_kwargs=None, **unused):
        super().__init__()
        self.has_cross_attention = True
        self.attn_num_head_channels = attn_num_head_channels
        resnets = [_resnet(in_channels, in_channels, temb_channels, resnet_eps, resnet_groups, output_scale_factor)]
        attentions, motion_modules = [], []
        for _ in range(num_layers):
            attentions.append(_transformer(attn_num_head_channels, in_channels, cross_attention_dim, resnet_groups,
                                           upcast_attention))
            motion_modules.append(_motion(in_channels, use_motion_module, motion_module_type, motion_module_kwargs))
            resnets.append(_resnet(in_channels, in_channels, temb_channels, resnet_eps, resnet_groups,
                                   output_scale_factor))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules) if use_motion_module else motion_modules

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                motion_module_alpha=1.0, cross_attention_kwargs=None, motion_cross_attention_kwargs=None):
        temb_r = _repeat_temb(temb, hidden_states.shape[2])
        hidden_states = _per_frame(self.resnets[0], hidden_states, temb_r)
        for attn, resnet, mm in zip(self.attentions, self.resnets[1:], self.motion_modules):
            hidden_states = _per_frame(attn, hidden_states, encoder_hidden_states=encoder_hidden_states,
                                       cross_attention_kwargs=cross_attention_kwargs)
            if mm is not None:
                hidden_states = mm(hidden_states, temb=temb, encoder_hidden_states=encoder_hidden_states,
                                   cross_attention_kwargs=motion_cross_attention_kwargs)
            hidden_states = _per_frame(resnet, hidden_states, temb_r)
        return hidden_states


class CrossAttnDownBlock3D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers=1, resnet_eps=1e-6, resnet_groups=32,
                 attn_num_head_channels=1, cross_attention_dim=1280, output_scale_factor=1.0, downsample_padding=1,
                 add_downsample=True, upcast_attention=False, use_motion_module=None, motion_module_type=None,
                 motion_module_kwargs=None, **unused):
        super().__init__()
        self.has_cross_attention = True
        self.attn_num_head_channels = attn_num_head_channels
        resnets, attentions, motion_modules = [], [], []
        for i in range(num_layers):
            resnets.append(_resnet(in_channels if i == 0 else out_channels, out_channels, temb_channels, resnet_eps,
                                   resnet_groups, output_scale_factor))
            attentions.append(_transformer(attn_num_head_channels, out_channels, cross_attention_dim, resnet_groups,
                                           upcast_attention))
            motion_modules.append(_motion(out_channels, use_motion_module, motion_module_type, motion_module_kwargs))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules) if use_motion_module else motion_modules
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                        padding=downsample_padding, name="op")]) if add_downsample else None

    def run_layers(self, hidden_states, temb, encoder_hidden_states, cross_attention_kwargs,
                   motion_cross_attention_kwargs):
        temb_r = _repeat_temb(temb, hidden_states.shape[2])
        output_states = ()
        for resnet, attn, mm in zip(self.resnets, self.attentions, self.motion_modules):
            hidden_states = _per_frame(resnet, hidden_states, temb_r)
            hidden_states = _per_frame(attn, hidden_states, encoder_hidden_states=encoder_hidden_states,
                                       cross_attention_kwargs=cross_attention_kwargs)
            if mm is not None:
                hidden_states = mm(hidden_states, temb=temb, encoder_hidden_states=encoder_hidden_states,
                                   cross_attention_kwargs=motion_cross_attention_kwargs)
            output_states += (hidden_states,)
        return hidden_states, output_states

    def run_downsample(self, hidden_states, output_states):
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = _per_frame(d, hidden_states)
            output_states += (hidden_states,)
        return hidden_states, output_states

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                motion_module_alpha=1.0, cross_attention_kwargs=None, motion_cross_attention_kwargs=None):
        hidden_states, output_states = self.run_layers(hidden_states, temb, encoder_hidden_states,
                                                       cross_attention_kwargs, motion_cross_attention_kwargs)
        return self.run_downsample(hidden_states, output_states)


class DownBlock3D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers=1, resnet_eps=1e-6, resnet_groups=32,
                 output_scale_factor=1.0, add_downsample=True, downsample_padding=1, use_motion_module=None,
                 motion_module_type=None, motion_module_kwargs=None, **unused):
        super().__init__()
        resnets, motion_modules = [], []
        for i in range(num_layers):
            resnets.append(_resnet(in_channels if i == 0 else out_channels, out_channels, temb_channels, resnet_eps,
                                   resnet_groups, output_scale_factor))
            motion_modules.append(_motion(out_channels, use_motion_module, motion_module_type, motion_module_kwargs))
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules) if use_motion_module else motion_modules
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                        padding=downsample_padding, name="op")]) if add_downsample else None

    def run_layers(self, hidden_states, temb, encoder_hidden_states, motion_cross_attention_kwargs):
        temb_r = _repeat_temb(temb, hidden_states.shape[2])
        output_states = ()
        for resnet, mm in zip(self.resnets, self.motion_modules):
            hidden_states = _per_frame(resnet, hidden_states, temb_r)
            if mm is not None:
                hidden_states = mm(hidden_states, temb=temb, encoder_hidden_states=encoder_hidden_states,
                                   cross_attention_kwargs=motion_cross_attention_kwargs)
            output_states += (hidden_states,)
        return hidden_states, output_states

    run_downsample = CrossAttnDownBlock3D.run_downsample

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, motion_module_alpha=1.0,
                motion_cross_attention_kwargs=None, **kwargs):
        hidden_states, output_states = self.run_layers(hidden_states, temb, encoder_hidden_states,
                                                       motion_cross_attention_kwargs)
        return self.run_downsample(hidden_states, output_states)


class CrossAttnUpBlock3D(nn.Module):
    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, num_layers=1, resnet_eps=1e-6,
                 resnet_groups=32, attn_num_head_channels=1, cross_attention_dim=1280, output_scale_factor=1.0,
                 add_upsample=True, upcast_attention=False, use_motion_module=None, motion_module_type=None,
                 motion_module_kwargs=None, **unused):
        super().__init__()
        self.has_cross_attention = True
        self.attn_num_head_channels = attn_num_head_channels
        resnets, attentions, motion_modules = [], [], []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            cin = prev_output_channel if i == 0 else out_channels
            resnets.append(_resnet(cin + skip, out_channels, temb_channels, resnet_eps, resnet_groups,
                                   output_scale_factor))
            attentions.append(_transformer(attn_num_head_channels, out_channels, cross_attention_dim, resnet_groups,
                                           upcast_attention))
            motion_modules.append(_motion(out_channels, use_motion_module, motion_module_type, motion_module_kwargs))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules) if use_motion_module else motion_modules
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)]) \
            if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None,
                upsample_size=None, attention_mask=None, motion_module_alpha=1.0, cross_attention_kwargs=None,
                motion_cross_attention_kwargs=None):
        temb_r = _repeat_temb(temb, hidden_states.shape[2])
        for resnet, attn, mm in zip(self.resnets, self.attentions, self.motion_modules):
            hidden_states = torch.cat([hidden_states, res_hidden_states_tuple[-1]], dim=1)
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = _per_frame(resnet, hidden_states, temb_r)
            hidden_states = _per_frame(attn, hidden_states, encoder_hidden_states=encoder_hidden_states,
                                       cross_attention_kwargs=cross_attention_kwargs)
            if mm is not None:
                hidden_states = mm(hidden_states, temb=temb, encoder_hidden_states=encoder_hidden_states,
                                   cross_attention_kwargs=motion_cross_attention_kwargs)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = _per_frame(u, hidden_states, upsample_size)
        return hidden_states


class UpBlock3D(nn.Module):
    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers=1, resnet_eps=1e-6,
                 resnet_groups=32, output_scale_factor=1.0, add_upsample=True, use_motion_module=None,
                 motion_module_type=None, motion_module_kwargs=None, **unused):
        super().__init__()
        resnets, motion_modules = [], []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            cin = prev_output_channel if i == 0 else out_channels
            resnets.append(_resnet(cin + skip, out_channels, temb_channels, resnet_eps, resnet_groups,
                                   output_scale_factor))
            motion_modules.append(_motion(out_channels, use_motion_module, motion_module_type, motion_module_kwargs))
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules) if use_motion_module else motion_modules
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)]) \
            if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None,
                encoder_hidden_states=None, motion_module_alpha=1.0, motion_cross_attention_kwargs=None, **kwargs):
        temb_r = _repeat_temb(temb, hidden_states.shape[2])
        for resnet, mm in zip(self.resnets, self.motion_modules):
            hidden_states = torch.cat([hidden_states, res_hidden_states_tuple[-1]], dim=1)
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = _per_frame(resnet, hidden_states, temb_r)
            if mm is not None:
                hidden_states = mm(hidden_states, temb=temb, encoder_hidden_states=encoder_hidden_states,
                                   cross_attention_kwargs=motion_cross_attention_kwargs)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = _per_frame(u, hidden_states, upsample_size)
        return hidden_states


def get_down_block(down_block_type, **kw):
    cls = {"DownBlock3D": DownBlock3D, "CrossAttnDownBlock3D": CrossAttnDownBlock3D}[down_block_type]
    return cls(**kw)


def get_up_block(up_block_type, **kw):
    cls = {"UpBlock3D": UpBlock3D, "CrossAttnUpBlock3D": CrossAttnUpBlock3D}[up_block_type]
    return cls(**kw)
