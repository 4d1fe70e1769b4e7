"""ORACLE (test infrastructure, CPU fp32; parity unpinned -- see oracle/diffusers_restated.py).

Restates fmc/models/unet.py and fmc/models/unet_cam_obj.py:
  UNet3DConditionModel            unet.py:49-826  (ctor :53-286, processor maps :323-468)
  UNet3DConditionModelPoseCond    unet.py:829-1300 (set_all_attn_processor :897-1031, forward :1033-1300)
  UNet3DConditionModelCamObjCond  unet_cam_obj.py (same + traj_features kwarg, forward :1107-...)
The dead options (class embeddings, fuse_first_frame fusers, controlnet residuals, attention masks) are not
restated; the constructor raises if they are requested.
"""
from types import SimpleNamespace

import torch
from torch import nn

from .attention_processor import (AttnProcessor, LoRAAttnProcessor, LORAPoseAdaptorAttnProcessor,
                                  PoseAdaptorAttnProcessor)
from .diffusers_restated import TimestepEmbedding, Timesteps
from .motion_module import InflatedConv3d
from .unet_blocks import UNetMidBlock3DCrossAttn, get_down_block, get_up_block

SD15_UNET_CONFIG = dict(
    sample_size=64, in_channels=4, out_channels=4, center_input_sample=False, flip_sin_to_cos=True, freq_shift=0,
    down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
    mid_block_type="UNetMidBlock3DCrossAttn",
    up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
    block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, downsample_padding=1, mid_block_scale_factor=1,
    act_fn="silu", norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=768, attention_head_dim=8,
)

# configs/cam.yaml:86-100 / configs/obj.yaml (identical block)
FMC_UNET_ADDITIONAL_KWARGS = dict(
    use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=False,
    motion_module_type="Vanilla",
    motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1,
                              attention_block_types=("Temporal_Self", "Temporal_Self"),
                              temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                              temporal_attention_dim_div=1, zero_initialize=False),
)


class UNet3DConditionModel(nn.Module):
    def __init__(self, sample_size=None, in_channels=4, out_channels=4, center_input_sample=False,
                 flip_sin_to_cos=True, freq_shift=0,
                 down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
                 mid_block_type="UNetMidBlock3DCrossAttn",
                 up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
                 only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 downsample_padding=1, mid_block_scale_factor=1, act_fn="silu", norm_num_groups=32, norm_eps=1e-5,
                 cross_attention_dim=1280, attention_head_dim=8, dual_cross_attention=False,
                 use_linear_projection=False, class_embed_type=None, num_class_embeds=None, upcast_attention=False,
                 resnet_time_scale_shift="default", use_motion_module=False, motion_module_resolutions=(1, 2, 4, 8),
                 motion_module_mid_block=False, motion_module_type=None, motion_module_kwargs=None,
                 fuse_first_frame=False):
        super().__init__()
        assert not (only_cross_attention or dual_cross_attention or use_linear_projection or fuse_first_frame)
        assert class_embed_type is None and num_class_embeds is None and act_fn == "silu"
        assert mid_block_type == "UNetMidBlock3DCrossAttn"
        motion_module_kwargs = dict(motion_module_kwargs or {})
        self.config = SimpleNamespace(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
            center_input_sample=center_input_sample, block_out_channels=tuple(block_out_channels),
            cross_attention_dim=cross_attention_dim, attention_head_dim=attention_head_dim,
            layers_per_block=layers_per_block, down_block_types=tuple(down_block_types),
            up_block_types=tuple(up_block_types), norm_num_groups=norm_num_groups, norm_eps=norm_eps)
        self.sample_size = sample_size
        self.in_channels = in_channels
        ch0 = block_out_channels[0]
        time_embed_dim = ch0 * 4
        self.conv_in = InflatedConv3d(in_channels, ch0, kernel_size=3, padding=(1, 1))
        self.time_proj = Timesteps(ch0, flip_sin_to_cos, freq_shift)
        self.time_embedding = TimestepEmbedding(ch0, time_embed_dim)
        heads = (attention_head_dim,) * len(down_block_types) if isinstance(attention_head_dim, int) \
            else tuple(attention_head_dim)

        self.down_blocks = nn.ModuleList()
        output_channel = ch0
        for i, btype in enumerate(down_block_types):
            input_channel, output_channel = output_channel, block_out_channels[i]
            is_final = i == len(block_out_channels) - 1
            self.down_blocks.append(get_down_block(
                btype, num_layers=layers_per_block, in_channels=input_channel, out_channels=output_channel,
                temb_channels=time_embed_dim, add_downsample=not is_final, resnet_eps=norm_eps,
                resnet_groups=norm_num_groups, cross_attention_dim=cross_attention_dim,
                attn_num_head_channels=heads[i], downsample_padding=downsample_padding,
                upcast_attention=upcast_attention,
                use_motion_module=use_motion_module and ((2 ** i) in motion_module_resolutions),
                motion_module_type=motion_module_type, motion_module_kwargs=motion_module_kwargs))

        self.mid_block = UNetMidBlock3DCrossAttn(
            in_channels=block_out_channels[-1], temb_channels=time_embed_dim, resnet_eps=norm_eps,
            output_scale_factor=mid_block_scale_factor, cross_attention_dim=cross_attention_dim,
            attn_num_head_channels=heads[-1], resnet_groups=norm_num_groups, upcast_attention=upcast_attention,
            use_motion_module=use_motion_module and motion_module_mid_block, motion_module_type=motion_module_type,
            motion_module_kwargs=motion_module_kwargs)

        self.num_upsamplers = 0
        self.up_blocks = nn.ModuleList()
        rev_ch = list(reversed(block_out_channels))
        rev_heads = list(reversed(heads))
        output_channel = rev_ch[0]
        for i, btype in enumerate(up_block_types):
            is_final = i == len(block_out_channels) - 1
            prev_output_channel, output_channel = output_channel, rev_ch[i]
            input_channel = rev_ch[min(i + 1, len(block_out_channels) - 1)]
            if not is_final:
                self.num_upsamplers += 1
            self.up_blocks.append(get_up_block(
                btype, num_layers=layers_per_block + 1, in_channels=input_channel, out_channels=output_channel,
                prev_output_channel=prev_output_channel, temb_channels=time_embed_dim, add_upsample=not is_final,
                resnet_eps=norm_eps, resnet_groups=norm_num_groups, cross_attention_dim=cross_attention_dim,
                attn_num_head_channels=rev_heads[i], upcast_attention=upcast_attention,
                use_motion_module=use_motion_module and ((2 ** (3 - i)) in motion_module_resolutions),
                motion_module_type=motion_module_type, motion_module_kwargs=motion_module_kwargs))

        self.conv_norm_out = nn.GroupNorm(num_channels=ch0, num_groups=norm_num_groups, eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = InflatedConv3d(ch0, out_channels, kernel_size=3, padding=1)

    # ---- processor maps: spatial vs motion-module attentions split on "motion_modules." (unet.py:323-468) ----
    def _collect_processors(self, want_motion):
        out = {}
        for name, module in self.named_modules():
            if hasattr(module, "set_processor") and (("motion_modules." in name) == want_motion):
                out[f"{name}.processor"] = module.processor
        return out

    def _assign_processors(self, processors, want_motion):
        for name, module in self.named_modules():
            if hasattr(module, "set_processor") and (("motion_modules." in name) == want_motion):
                module.set_processor(processors[f"{name}.processor"] if isinstance(processors, dict) else processors)

    @property
    def attn_processors(self):
        return self._collect_processors(False)

    @property
    def mm_attn_processors(self):
        return self._collect_processors(True)

    def set_attn_processor(self, processor):
        self._assign_processors(processor, False)

    def set_mm_attn_processor(self, processor):
        self._assign_processors(processor, True)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device


class UNet3DConditionModelPoseCond(UNet3DConditionModel):
    _accepts_traj_features = False

    def __init__(self, decoder_add_posecond=True, **kwargs):
        super().__init__(**kwargs)
        self.decoder_add_posecond = decoder_add_posecond

    def _hidden_size_of(self, name):
        ch = self.config.block_out_channels
        if name.startswith("mid_block"):
            return ch[-1], -1
        if name.startswith("up_blocks"):
            i = int(name[len("up_blocks.")])
            return list(reversed(ch))[i], i
        i = int(name[len("down_blocks.")])
        return ch[i], i

    def set_all_attn_processor(self, add_spatial=False, spatial_attn_names="attn1", add_temporal=False,
                               add_spatial_lora=True, add_motion_lora=False, temporal_attn_names="0",
                               pose_feature_dimensions=(320, 640, 1280, 1280), lora_kwargs=None,
                               motion_lora_kwargs=None, **attention_processor_kwargs):
        """(unet.py:897-1031) choose one processor per attention.  rank rule: `rank if rank > 16 else hidden // rank`."""
        lora_kwargs = dict(lora_kwargs or {})
        motion_lora_kwargs = dict(motion_lora_kwargs or {})
        lora_rank = lora_kwargs.pop("lora_rank")
        motion_lora_rank = motion_lora_kwargs.pop("lora_rank")

        def rank_of(r, hidden):
            return r if r > 16 else hidden // r

        def pose_dim(name, idx, is_up):
            dims = list(reversed(pose_feature_dimensions)) if is_up else list(pose_feature_dimensions)
            return dims[idx]

        def build(names, add_pose, pose_names, add_lora, rank, extra_lora_kwargs, is_motion):
            procs = {}
            chosen = pose_names.split(",")
            for name in names:
                hidden, idx = self._hidden_size_of(name)
                attention_name = name.split(".")[-2]
                if is_motion:
                    cross_dim = None
                else:
                    cross_dim = None if attention_name == "attn1" else self.config.cross_attention_dim
                with_pose = add_pose and attention_name in chosen
                if with_pose and is_motion and name.startswith("up_blocks"):
                    with_pose = self.decoder_add_posecond
                pdim = pose_dim(name, idx, name.startswith("up_blocks")) if with_pose else None
                if with_pose and add_lora:
                    procs[name] = LORAPoseAdaptorAttnProcessor(hidden_size=hidden, pose_feature_dim=pdim,
                                                               cross_attention_dim=cross_dim,
                                                               rank=rank_of(rank, hidden),
                                                               **attention_processor_kwargs, **extra_lora_kwargs)
                elif with_pose:
                    procs[name] = PoseAdaptorAttnProcessor(hidden_size=hidden, pose_feature_dim=pdim,
                                                           cross_attention_dim=cross_dim,
                                                           **attention_processor_kwargs)
                elif add_lora:
                    # note: lora_scale from lora_kwargs is dropped here (unet.py:962-966); default 1.0 applies
                    procs[name] = LoRAAttnProcessor(hidden_size=hidden, cross_attention_dim=cross_dim,
                                                    rank=rank_of(rank, hidden))
                else:
                    procs[name] = AttnProcessor()
            return procs

        self.set_attn_processor(build(list(self.attn_processors.keys()), add_spatial, spatial_attn_names,
                                      add_spatial_lora, lora_rank, lora_kwargs, False))
        self.set_mm_attn_processor(build(list(self.mm_attn_processors.keys()), add_temporal, temporal_attn_names,
                                         add_motion_lora, motion_lora_rank, motion_lora_kwargs, True))

    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, attention_mask=None,
                cross_attention_kwargs=None, pose_embedding_features=None, traj_features=None, return_dict=True,
                **unused):
        assert attention_mask is None and class_labels is None and cross_attention_kwargs is None
        if traj_features is not None:
            assert self._accepts_traj_features
        up_factor = 2 ** self.num_upsamplers
        forward_upsample_size = any(s % up_factor != 0 for s in sample.shape[-2:])
        upsample_size = None
        if self.config.center_input_sample:
            sample = 2 * sample - 1.0

        timesteps = timestep
        if not torch.is_tensor(timesteps):
            dtype = torch.float64 if isinstance(timestep, float) else torch.int64
            timesteps = torch.tensor([timesteps], dtype=dtype, device=sample.device)
        elif timesteps.ndim == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        emb = self.time_embedding(self.time_proj(timesteps).to(dtype=self.dtype))

        f = sample.shape[2]
        encoder_hidden_states = encoder_hidden_states.repeat_interleave(f, dim=0)  # 'b n c -> (b f) n c'
        sample = self.conv_in(sample)

        def spatial_kwargs(feat):
            kw = {"pose_feature": feat}
            if self._accepts_traj_features:
                kw["traj_features"] = traj_features  # unet_cam_obj.py:1222-1223 (cross-attn down blocks only)
            return kw

        down_res = (sample,)
        for block, feat in zip(self.down_blocks, pose_embedding_features):
            if getattr(block, "has_cross_attention", False):
                sample, res = block(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states,
                                    attention_mask=None, cross_attention_kwargs=spatial_kwargs(feat),
                                    motion_cross_attention_kwargs={"pose_feature": feat})
            else:
                # DownBlock3D receives cross_attention_kwargs as an ordinary keyword (swallowed by **kwargs), so a
                # nested 'traj_features' never reaches Adapted_DownBlock3D_forward
                sample, res = block(hidden_states=sample, temb=emb, cross_attention_kwargs={"pose_feature": feat},
                                    motion_cross_attention_kwargs={"pose_feature": feat})
            down_res += res

        feat = pose_embedding_features[-1]
        sample = self.mid_block(sample, emb, encoder_hidden_states=encoder_hidden_states, attention_mask=None,
                                cross_attention_kwargs={"pose_feature": feat},
                                motion_cross_attention_kwargs={"pose_feature": feat})

        for i, block in enumerate(self.up_blocks):
            is_final = i == len(self.up_blocks) - 1
            n = len(block.resnets)
            res, down_res = down_res[-n:], down_res[:-n]
            if not is_final and forward_upsample_size:
                upsample_size = down_res[-1].shape[2:]
            if self.decoder_add_posecond:
                feat = pose_embedding_features[-(i + 1)]
                ckw, mkw = {"pose_feature": feat}, {"pose_feature": feat}
            else:
                ckw, mkw = None, None
            if getattr(block, "has_cross_attention", False):
                sample = block(hidden_states=sample, temb=emb, res_hidden_states_tuple=res,
                               encoder_hidden_states=encoder_hidden_states, upsample_size=upsample_size,
                               attention_mask=None, cross_attention_kwargs=ckw,
                               **({"motion_cross_attention_kwargs": mkw} if mkw is not None else {}))
            else:
                sample = block(hidden_states=sample, temb=emb, res_hidden_states_tuple=res,
                               upsample_size=upsample_size, cross_attention_kwargs=ckw,
                               **({"motion_cross_attention_kwargs": mkw} if mkw is not None else {}))

        b, c, f, h, w = sample.shape
        x = sample.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        x = self.conv_norm_out(x).reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)
        sample = self.conv_out(self.conv_act(x))
        return SimpleNamespace(sample=sample) if return_dict else (sample,)


class UNet3DConditionModelCamObjCond(UNet3DConditionModelPoseCond):
    _accepts_traj_features = True
