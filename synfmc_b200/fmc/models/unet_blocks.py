"""Mirror of fmc/models/unet_blocks.py: the five 3-D U-Net blocks with the reference's class names, constructor
arguments, attribute names (resnets / attentions / motion_modules / downsamplers / upsamplers) and forward signatures
(so the OMC trainer's rebinding of `.forward`, train_cam_obj_ctrl.py:317-329, keeps working).  Activations are
`engine.CL` (channels-last bf16); every per-frame op works on the [(b f), h, w, C] view without a copy."""
from torch import nn

from ... import engine
from ...engine import CL
from .._blocks import Downsample2D, ResnetBlock2D, Transformer2DModel, Upsample2D
from .motion_module import get_motion_module


def _resnet(cin, cout, temb_channels, eps, groups, scale=1.0):
    return ResnetBlock2D(in_channels=cin, out_channels=cout, temb_channels=temb_channels, eps=eps, groups=groups,
                         output_scale_factor=scale)


def _transformer(heads, channels, cross_attention_dim, groups):
    return Transformer2DModel(heads, channels // heads, in_channels=channels, num_layers=1,
                              cross_attention_dim=cross_attention_dim, norm_num_groups=groups)


def _motion(channels, use, mtype, mkwargs):
    return get_motion_module(in_channels=channels, motion_module_type=mtype, motion_module_kwargs=mkwargs) if use else None


def _spatial_kwargs_ok(kw):
    # spatial processors receive {"pose_feature": ...} and ignore it (attention_processor.py:115); anything else is an error
    extra = set(kw or {}) - {"pose_feature", "scale"}
    if extra:
        raise TypeError(f"unexpected cross_attention_kwargs for the spatial attention: {sorted(extra)}")
    if (kw or {}).get("scale") is not None:
        # attention_processor.py:118: a call-time `scale` overrides the processor's lora_scale.  Domain-LoRA is folded into
        # the projection weights once per plan with processor.lora_scale, so a per-call value cannot be honoured: refuse
        # (set `processor.lora_scale` instead -- the plans follow it)
        raise NotImplementedError("cross_attention_kwargs['scale'] (call-time LoRA scale) is not supported: set "
                                  "processor.lora_scale, the folded weights are rebuilt automatically")


class UNetMidBlock3DCrossAttn(nn.Module):
    def __init__(self, in_channels, temb_channels, num_layers=1, resnet_eps=1e-6, resnet_groups=32,
                 attn_num_head_channels=1, output_scale_factor=1.0, cross_attention_dim=1280,
                 use_motion_module=None, motion_module_type=None, motion_module_kwargs=None, **unused):
        super().__init__()
        self.has_cross_attention = True
        self.attn_num_head_channels = attn_num_head_channels
        resnets = [_resnet(in_channels, in_channels, temb_channels, resnet_eps, resnet_groups, output_scale_factor)]
        attentions, motion_modules = [], []
        for _ in range(num_layers):
            attentions.append(_transformer(attn_num_head_channels, in_channels, cross_attention_dim, resnet_groups))
            motion_modules.append(_motion(in_channels, use_motion_module, motion_module_type, motion_module_kwargs))
            resnets.append(_resnet(in_channels, in_channels, temb_channels, resnet_eps, resnet_groups, output_scale_factor))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules) if use_motion_module else motion_modules

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                motion_module_alpha=1.0, cross_attention_kwargs=None, motion_cross_attention_kwargs=None):
        _spatial_kwargs_ok(cross_attention_kwargs)
        x = CL.from_reference(hidden_states)
        x = engine.run_resnet(self.resnets[0], x, temb)
        for attn, resnet, mm in zip(self.attentions, self.resnets[1:], self.motion_modules):
            x = engine.run_transformer2d(attn, x, encoder_hidden_states)
            if mm is not None:
                x = mm(x, temb=temb, encoder_hidden_states=encoder_hidden_states,
                       cross_attention_kwargs=motion_cross_attention_kwargs)
            x = engine.run_resnet(resnet, x, temb)
        return x


class CrossAttnDownBlock3D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers=1, resnet_eps=1e-6, resnet_groups=32,
                 attn_num_head_channels=1, cross_attention_dim=1280, output_scale_factor=1.0, downsample_padding=1,
                 add_downsample=True, use_motion_module=None, motion_module_type=None, motion_module_kwargs=None,
                 **unused):
        super().__init__()
        self.has_cross_attention = True
        self.attn_num_head_channels = attn_num_head_channels
        resnets, attentions, motion_modules = [], [], []
        for i in range(num_layers):
            resnets.append(_resnet(in_channels if i == 0 else out_channels, out_channels, temb_channels, resnet_eps,
                                   resnet_groups, output_scale_factor))
            attentions.append(_transformer(attn_num_head_channels, out_channels, cross_attention_dim, resnet_groups))
            motion_modules.append(_motion(out_channels, use_motion_module, motion_module_type, motion_module_kwargs))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules) if use_motion_module else motion_modules
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                        padding=downsample_padding, name="op")]) if add_downsample else None
        self.gradient_checkpointing = False

    def run_layers(self, hidden_states, temb, encoder_hidden_states, cross_attention_kwargs,
                   motion_cross_attention_kwargs):
        _spatial_kwargs_ok(cross_attention_kwargs)
        x = CL.from_reference(hidden_states)
        output_states = ()
        for resnet, attn, mm in zip(self.resnets, self.attentions, self.motion_modules):
            x = engine.run_resnet(resnet, x, temb)
            x = engine.run_transformer2d(attn, x, encoder_hidden_states)
            if mm is not None:
                x = mm(x, temb=temb, encoder_hidden_states=encoder_hidden_states,
                       cross_attention_kwargs=motion_cross_attention_kwargs)
            output_states += (x,)
        return x, output_states

    def run_downsample(self, x, output_states):
        if self.downsamplers is not None:
            for d in self.downsamplers:
                x = engine.run_downsample(d, x)
            output_states += (x,)
        return x, output_states

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                motion_module_alpha=1.0, cross_attention_kwargs=None, motion_cross_attention_kwargs=None):
        x, output_states = self.run_layers(hidden_states, temb, encoder_hidden_states, cross_attention_kwargs,
                                           motion_cross_attention_kwargs)
        return self.run_downsample(x, output_states)


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
        self.gradient_checkpointing = False

    def run_layers(self, hidden_states, temb, encoder_hidden_states, motion_cross_attention_kwargs):
        x = CL.from_reference(hidden_states)
        output_states = ()
        for resnet, mm in zip(self.resnets, self.motion_modules):
            x = engine.run_resnet(resnet, x, temb)
            if mm is not None:
                x = mm(x, temb=temb, encoder_hidden_states=encoder_hidden_states,
                       cross_attention_kwargs=motion_cross_attention_kwargs)
            output_states += (x,)
        return x, output_states

    run_downsample = CrossAttnDownBlock3D.run_downsample

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, motion_module_alpha=1.0,
                motion_cross_attention_kwargs=None, **kwargs):
        x, output_states = self.run_layers(hidden_states, temb, encoder_hidden_states, motion_cross_attention_kwargs)
        return self.run_downsample(x, output_states)


class CrossAttnUpBlock3D(nn.Module):
    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, num_layers=1, resnet_eps=1e-6,
                 resnet_groups=32, attn_num_head_channels=1, cross_attention_dim=1280, output_scale_factor=1.0,
                 add_upsample=True, use_motion_module=None, motion_module_type=None, motion_module_kwargs=None, **unused):
        super().__init__()
        self.has_cross_attention = True
        self.attn_num_head_channels = attn_num_head_channels
        resnets, attentions, motion_modules = [], [], []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            cin = prev_output_channel if i == 0 else out_channels
            resnets.append(_resnet(cin + skip, out_channels, temb_channels, resnet_eps, resnet_groups, output_scale_factor))
            attentions.append(_transformer(attn_num_head_channels, out_channels, cross_attention_dim, resnet_groups))
            motion_modules.append(_motion(out_channels, use_motion_module, motion_module_type, motion_module_kwargs))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules) if use_motion_module else motion_modules
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)]) \
            if add_upsample else None
        self.gradient_checkpointing = False

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None,
                upsample_size=None, attention_mask=None, motion_module_alpha=1.0, cross_attention_kwargs=None,
                motion_cross_attention_kwargs=None):
        _spatial_kwargs_ok(cross_attention_kwargs)
        x = CL.from_reference(hidden_states)
        for resnet, attn, mm in zip(self.resnets, self.attentions, self.motion_modules):
            x = engine.concat_channels(x, CL.from_reference(res_hidden_states_tuple[-1]))
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            x = engine.run_resnet(resnet, x, temb)
            x = engine.run_transformer2d(attn, x, encoder_hidden_states)
            if mm is not None:
                x = mm(x, temb=temb, encoder_hidden_states=encoder_hidden_states,
                       cross_attention_kwargs=motion_cross_attention_kwargs)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                x = engine.run_upsample(u, x, upsample_size)
        return x


class UpBlock3D(nn.Module):
    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers=1, resnet_eps=1e-6,
                 resnet_groups=32, output_scale_factor=1.0, add_upsample=True, use_motion_module=None,
                 motion_module_type=None, motion_module_kwargs=None, **unused):
        super().__init__()
        resnets, motion_modules = [], []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            cin = prev_output_channel if i == 0 else out_channels
            resnets.append(_resnet(cin + skip, out_channels, temb_channels, resnet_eps, resnet_groups, output_scale_factor))
            motion_modules.append(_motion(out_channels, use_motion_module, motion_module_type, motion_module_kwargs))
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules) if use_motion_module else motion_modules
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)]) \
            if add_upsample else None
        self.gradient_checkpointing = False

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None,
                encoder_hidden_states=None, motion_module_alpha=1.0, motion_cross_attention_kwargs=None, **kwargs):
        x = CL.from_reference(hidden_states)
        for resnet, mm in zip(self.resnets, self.motion_modules):
            x = engine.concat_channels(x, CL.from_reference(res_hidden_states_tuple[-1]))
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            x = engine.run_resnet(resnet, x, temb)
            if mm is not None:
                x = mm(x, temb=temb, encoder_hidden_states=encoder_hidden_states,
                       cross_attention_kwargs=motion_cross_attention_kwargs)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                x = engine.run_upsample(u, x, upsample_size)
        return x


def get_down_block(down_block_type, **kw):
    down_block_type = down_block_type[7:] if down_block_type.startswith("UNetRes") else down_block_type
    cls = {"DownBlock3D": DownBlock3D, "CrossAttnDownBlock3D": CrossAttnDownBlock3D}[down_block_type]
    return cls(**kw)


def get_up_block(up_block_type, **kw):
    up_block_type = up_block_type[7:] if up_block_type.startswith("UNetRes") else up_block_type
    cls = {"UpBlock3D": UpBlock3D, "CrossAttnUpBlock3D": CrossAttnUpBlock3D}[up_block_type]
    return cls(**kw)
