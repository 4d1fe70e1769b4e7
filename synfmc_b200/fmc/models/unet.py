"""Mirror of fmc/models/unet.py: UNet3DConditionModel (:49-826) and UNet3DConditionModelPoseCond (:829-1300) with the
reference's constructor arguments, processor maps, `set_all_attn_processor` and forward signature; state-dict keys
are the reference's.  forward() converts the fp32 `[B, 4, f, h, w]` sample to channels-last bf16 once,
runs every block on the libfmc_b200 kernels and converts the prediction back."""
import json
import os
from types import SimpleNamespace

import torch
from torch import nn

from ... import engine, ops
from ...engine import CL
from .._blocks import TimestepEmbedding, Timesteps
from .attention_processor import (AttnProcessor, LoRAAttnProcessor, LORAPoseAdaptorAttnProcessor,
                                  PoseAdaptorAttnProcessor)
from .resnet import InflatedConv3d
from .unet_blocks import UNetMidBlock3DCrossAttn, get_down_block, get_up_block

CustomizedAttnProcessor = AttnProcessor
CustomizedLoRAAttnProcessor = LoRAAttnProcessor


class UNet3DConditionOutput(SimpleNamespace):
    pass


class UNet3DConditionModel(nn.Module):
    def __init__(self, sample_size=None, in_channels=4, out_channels=4, center_input_sample=False,
                 flip_sin_to_cos=True, freq_shift=0,
                 down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
                 mid_block_type="UNetMidBlock3DCrossAttn",
                 up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
                 only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 downsample_padding=1, mid_block_scale_factor=1, act_fn="silu", norm_num_groups=32, norm_eps=1e-5,
                 cross_attention_dim=1280, attention_head_dim=8, dual_cross_attention=False,
                 use_linear_projection=False, class_embed_type=None, addition_embed_type=None, num_class_embeds=None,
                 upcast_attention=False, resnet_time_scale_shift="default", use_motion_module=False,
                 motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=False, motion_module_type=None,
                 motion_module_kwargs=None, fuse_first_frame=False):
        super().__init__()
        unsupported = dict(only_cross_attention=only_cross_attention, dual_cross_attention=dual_cross_attention,
                           use_linear_projection=use_linear_projection, fuse_first_frame=fuse_first_frame,
                           class_embed_type=class_embed_type, num_class_embeds=num_class_embeds,
                           center_input_sample=center_input_sample)
        bad = {k: v for k, v in unsupported.items() if v}
        if bad or act_fn != "silu" or mid_block_type != "UNetMidBlock3DCrossAttn":
            raise NotImplementedError(f"options outside the FMC hot path (SURVEY.md section 2): {bad}")
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
                use_motion_module=use_motion_module and ((2 ** i) in motion_module_resolutions),
                motion_module_type=motion_module_type, motion_module_kwargs=motion_module_kwargs))
        self.mid_block = UNetMidBlock3DCrossAttn(
            in_channels=block_out_channels[-1], temb_channels=time_embed_dim, resnet_eps=norm_eps,
            output_scale_factor=mid_block_scale_factor, cross_attention_dim=cross_attention_dim,
            attn_num_head_channels=heads[-1], resnet_groups=norm_num_groups,
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
                attn_num_head_channels=rev_heads[i],
                use_motion_module=use_motion_module and ((2 ** (3 - i)) in motion_module_resolutions),
                motion_module_type=motion_module_type, motion_module_kwargs=motion_module_kwargs))
        self.conv_norm_out = nn.GroupNorm(num_channels=ch0, num_groups=norm_num_groups, eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = InflatedConv3d(ch0, out_channels, kernel_size=3, padding=1)
        self._plan = None
        self._cat_cache = {}   # concatenated time-embedding / text K|V projection weights (dropped with the plans)
        self._resnets = None  # ResnetBlock2D modules, for the batched time-embedding projection
        self._transformers = None  # Transformer2DModel modules, for the batched text K | V projection

    # ---- processor maps, split on "motion_modules." (unet.py:323-468) ----
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
        self.invalidate_plans()

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

    def invalidate_plans(self):
        """Drop every cached device plan (folded / fused weights) and everything derived from them (the concatenated
        projection weights, captured CUDA graphs -- their keys carry engine's plan generation).  Called automatically when
        the parameter fingerprint changes (optimizer steps, in-place copies, load_state_dict on a submodule, processor
        edits); calling it by hand is never needed but harmless."""
        engine.invalidate_plans(self)
        self._cat_cache = {}

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self.invalidate_plans()
        return out

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    @classmethod
    def from_pretrained_2d(cls, pretrained_model_path, subfolder=None, unet_additional_kwargs=None, logger=None):
        """unet.py:762-826: build from an SD1.5 `unet/config.json`, rename the 2-D block types to their 3-D
        counterparts and load the 2-D weights non-strictly (motion-module keys stay at their initial values).
        Like diffusers' `from_config` behind the reference (with the `extract_init_dict` override at unet.py:832-880),
        entries of the config and of `unet_additional_kwargs` that no constructor takes are dropped and reported, not an
        error: the shipped yamls carry `unet_use_cross_frame_attention` / `unet_use_temporal_attention`
        (configs/cam.yaml:88-89), which nothing reads; `unet_additional_kwargs` wins over the config file."""
        if logger is not None:
            logger.info(f"Loading unet's pretrained weights from {pretrained_model_path} ...")
        path = os.path.join(pretrained_model_path, subfolder) if subfolder is not None else pretrained_model_path
        config_file = os.path.join(path, "config.json")
        if not os.path.isfile(config_file):
            raise RuntimeError(f"{config_file} does not exist")
        with open(config_file) as f:
            config = json.load(f)
        config = {k: v for k, v in config.items() if not k.startswith("_")}
        config["down_block_types"] = ["CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"]
        config["up_block_types"] = ["UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"]
        if "mid_block_type" in config:
            config["mid_block_type"] = "UNetMidBlock3DCrossAttn"
        import inspect
        accepted = (set(inspect.signature(UNet3DConditionModel.__init__).parameters) |
                    set(inspect.signature(cls.__init__).parameters)) - {"self", "kwargs"}
        merged = dict(config)
        merged.update(unet_additional_kwargs or {})
        unused = {k: v for k, v in merged.items() if k not in accepted}
        if logger is not None:
            logger.info("please check unused kwargs in 'unet_additional_kwargs' config:")
            for k, v in unused.items():
                logger.info(f"{k:50s}: {repr(v)}")
        model = cls(**{k: v for k, v in merged.items() if k in accepted})
        weights = None
        for name in ("diffusion_pytorch_model.bin", "diffusion_pytorch_model.safetensors"):
            file = os.path.join(path, name)
            if os.path.isfile(file):
                if name.endswith(".safetensors"):
                    from safetensors.torch import load_file
                    weights = load_file(file)
                else:
                    weights = torch.load(file, map_location="cpu")
                break
        if weights is None:
            raise RuntimeError(f"{os.path.join(path, 'diffusion_pytorch_model.bin')} does not exist")
        missing, unexpected = model.load_state_dict(weights, strict=False)
        print(f"### missing keys: {len(missing)}; \n### unexpected keys: {len(unexpected)};")
        params = [p.numel() if "motion_modules." in n else 0 for n, p in model.named_parameters()]
        print(f"### Motion Module Parameters: {sum(params) / 1e6} M")
        return model


def _reject_unsupported(down_block_additional_residuals, mid_block_additional_residual, motion_module_alphas, debug):
    """trailing parameters of the reference's forward (unet.py:1033-1047) that no trainer or pipeline of FMC sets"""
    if down_block_additional_residuals is not None or mid_block_additional_residual is not None:
        raise NotImplementedError("ControlNet-style additional residuals are not part of the FMC hot path")
    if not (isinstance(motion_module_alphas, (int, float)) and float(motion_module_alphas) == 1.0):
        raise NotImplementedError("motion_module_alphas other than 1.0 is not part of the FMC hot path")
    if debug:
        raise NotImplementedError("debug=True (intermediate feature dump) is not supported")


class UNet3DConditionModelPoseCond(UNet3DConditionModel):
    _accepts_traj_features = False

    def __init__(self, decoder_add_posecond=True, **kwargs):
        super().__init__(**kwargs)
        self.decoder_add_posecond = decoder_add_posecond

    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, attention_mask=None,
                cross_attention_kwargs=None, pose_embedding_features=None, return_dict=True,
                down_block_additional_residuals=None, mid_block_additional_residual=None, motion_module_alphas=1.0,
                debug=False):
        """Parameter order and defaults of the reference's forward (unet.py:1033-1047)."""
        _reject_unsupported(down_block_additional_residuals, mid_block_additional_residual, motion_module_alphas, debug)
        return self._forward_impl(sample, timestep, encoder_hidden_states, class_labels, attention_mask,
                                  cross_attention_kwargs, pose_embedding_features, None, return_dict)

    def _hidden_size_of(self, name):
        ch = self.config.block_out_channels
        if name.startswith("mid_block"):
            return ch[-1], -1
        if name.startswith("up_blocks"):
            i = int(name[len("up_blocks."):].split(".")[0])
            return list(reversed(ch))[i], i
        i = int(name[len("down_blocks."):].split(".")[0])
        return ch[i], i

    def set_all_attn_processor(self, add_spatial=False, spatial_attn_names="attn1", add_temporal=False,
                               add_spatial_lora=True, add_motion_lora=False, temporal_attn_names="0",
                               pose_feature_dimensions=(320, 640, 1280, 1280), lora_kwargs=None,
                               motion_lora_kwargs=None, **attention_processor_kwargs):
        """unet.py:897-1031.  rank rule: `rank if rank > 16 else hidden // rank` (lora_rank 2 -> rank C/2)."""
        lora_kwargs = dict(lora_kwargs or {})
        motion_lora_kwargs = dict(motion_lora_kwargs or {})
        lora_rank = lora_kwargs.pop("lora_rank")
        motion_lora_rank = motion_lora_kwargs.pop("lora_rank")

        def rank_of(r, hidden):
            return r if r > 16 else hidden // r

        def build(names, add_pose, pose_names, add_lora, rank, extra, is_motion):
            procs = {}
            chosen = pose_names.split(",")
            for name in names:
                hidden, idx = self._hidden_size_of(name)
                attention_name = name.split(".")[-2]
                cross_dim = None if (is_motion or attention_name == "attn1") else self.config.cross_attention_dim
                with_pose = add_pose and attention_name in chosen
                if with_pose and is_motion and name.startswith("up_blocks"):
                    with_pose = self.decoder_add_posecond
                dims = list(reversed(pose_feature_dimensions)) if name.startswith("up_blocks") else list(pose_feature_dimensions)
                pdim = dims[idx] if with_pose else None
                if with_pose and add_lora:
                    procs[name] = LORAPoseAdaptorAttnProcessor(hidden_size=hidden, pose_feature_dim=pdim,
                                                               cross_attention_dim=cross_dim, rank=rank_of(rank, hidden),
                                                               **attention_processor_kwargs, **extra)
                elif with_pose:
                    procs[name] = PoseAdaptorAttnProcessor(hidden_size=hidden, pose_feature_dim=pdim,
                                                           cross_attention_dim=cross_dim, **attention_processor_kwargs)
                elif add_lora:
                    procs[name] = CustomizedLoRAAttnProcessor(hidden_size=hidden, cross_attention_dim=cross_dim,
                                                              rank=rank_of(rank, hidden))
                else:
                    procs[name] = CustomizedAttnProcessor()
            return procs

        self.set_attn_processor(build(list(self.attn_processors.keys()), add_spatial, spatial_attn_names,
                                      add_spatial_lora, lora_rank, lora_kwargs, False))
        self.set_mm_attn_processor(build(list(self.mm_attn_processors.keys()), add_temporal, temporal_attn_names,
                                         add_motion_lora, motion_lora_rank, motion_lora_kwargs, True))

    def text_context(self, encoder_hidden_states, device):
        """encoder_hidden_states [B, 77, 768] -> engine.TextCtx with the K | V projections of all 16 text cross-attentions
        done (one GEMM).  The text is constant over a denoising loop, so the pipeline builds this once and passes it as
        `encoder_hidden_states` to every step instead of re-projecting inside each forward."""
        engine.refresh_plans(self)
        text = engine.TextCtx(encoder_hidden_states, device)
        if self._transformers is None:
            self._transformers = [m for m in self.modules() if m.__class__.__name__ == "Transformer2DModel"]
        text.project_all(self._transformers, device, self._cat_cache)
        return text

    # ---- device plan of the UNet-level layers ----
    def plan(self, device):
        if self._plan is None or self._plan["device"] != engine.plan_key(device):
            def pad_conv(conv, cin_pad, cout_pad):
                w = torch.zeros(cout_pad, cin_pad, *conv.weight.shape[2:])
                w[:conv.out_channels, :conv.in_channels] = conv.weight.detach().float()
                b = torch.zeros(cout_pad)
                b[:conv.out_channels] = conv.bias.detach().float()
                padded = nn.Conv2d(cin_pad, cout_pad, conv.kernel_size, padding=conv.padding)
                padded.weight.data.copy_(w)
                padded.bias.data.copy_(b)
                return engine.ConvPlan(padded, device)

            te = self.time_embedding
            self._plan = {
                "device": engine.plan_key(device),
                # the 4 latent channels are zero-padded to 64 in / 32 out: one k-block / one output sub-tile of the tcgen05
                # implicit-GEMM convolution (fmc_conv3x3_bf16), which then carries conv_in / conv_out as well
                "conv_in": pad_conv(self.conv_in, 64, self.conv_in.out_channels),
                "conv_out": pad_conv(self.conv_out, self.conv_out.in_channels, 32),
                "t1": engine.LinearPlan(te.linear_1.weight.detach().float(), te.linear_1.bias.detach().float(), device),
                "t2": engine.LinearPlan(te.linear_2.weight.detach().float(), te.linear_2.bias.detach().float(), device),
                "norm_out": engine.NormPlan(self.conv_norm_out, device),
            }
        return self._plan

    def _forward_impl(self, sample, timestep, encoder_hidden_states, class_labels=None, attention_mask=None,
                      cross_attention_kwargs=None, pose_embedding_features=None, traj_features=None, return_dict=True):
        engine.require_no_grad(self, sample, encoder_hidden_states, *(pose_embedding_features or ()),
                               *(traj_features or ()))
        if attention_mask is not None or class_labels is not None or cross_attention_kwargs is not None:
            raise NotImplementedError("attention_mask / class_labels / cross_attention_kwargs are unused on the FMC hot path")
        if traj_features is not None and not self._accepts_traj_features:
            raise TypeError("traj_features is only accepted by UNet3DConditionModelCamObjCond")
        ops.require_cuda(sample)
        device = sample.device
        engine.refresh_plans(self)
        p = self.plan(device)
        B, _, F, H, W = sample.shape
        up_factor = 2 ** self.num_upsamplers
        forward_upsample_size = any(s % up_factor != 0 for s in (H, W))
        upsample_size = None

        # time embedding (unet.py:1075-1096): sinusoid -> Linear -> SiLU -> Linear
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.float32, device=device)
        elif timesteps.ndim == 0:
            timesteps = timesteps[None]
        timesteps = timesteps.to(device=device, dtype=torch.float32).expand(B).contiguous()
        adt = engine.act_dtype()
        t_emb = ops.timestep_embedding(timesteps, self.config.block_out_channels[0], dtype=adt)
        h1 = p["t1"].f32out(t_emb)
        emb = p["t2"].f32out(ops.cast_act(h1, silu=True, dtype=adt))
        temb = engine.Temb(emb)
        if self._resnets is None:
            self._resnets = [m for m in self.modules() if m.__class__.__name__ == "ResnetBlock2D"]
        temb.project_all(self._resnets, device, self._cat_cache)

        text = engine.TextCtx.of(encoder_hidden_states, B, 1, device)
        if self._transformers is None:
            self._transformers = [m for m in self.modules() if m.__class__.__name__ == "Transformer2DModel"]
        if not text.kv:
            text.project_all(self._transformers, device, self._cat_cache)
        x = CL(ops.to_channels_last(sample, c_pad=64, dtype=adt))
        y = p["conv_in"](x.images())
        x = CL(y.view(B, F, H, W, y.shape[-1]))

        feats = [engine.as_cl_feature(f) for f in pose_embedding_features]

        def spatial_kwargs(feat):
            kw = {"pose_feature": feat}
            if self._accepts_traj_features:
                kw["traj_features"] = traj_features  # unet_cam_obj.py:1222-1223: cross-attention down blocks only
            return kw

        down_res = (x,)
        for block, feat in zip(self.down_blocks, feats):
            if getattr(block, "has_cross_attention", False):
                x, res = block(hidden_states=x, temb=temb, encoder_hidden_states=text, attention_mask=None,
                               cross_attention_kwargs=spatial_kwargs(feat),
                               motion_cross_attention_kwargs={"pose_feature": feat})
            else:
                x, res = block(hidden_states=x, temb=temb, cross_attention_kwargs={"pose_feature": feat},
                               motion_cross_attention_kwargs={"pose_feature": feat})
            down_res += res

        feat = feats[-1]
        x = self.mid_block(x, temb, encoder_hidden_states=text, attention_mask=None,
                           cross_attention_kwargs={"pose_feature": feat},
                           motion_cross_attention_kwargs={"pose_feature": feat})

        for i, block in enumerate(self.up_blocks):
            is_final = i == len(self.up_blocks) - 1
            n = len(block.resnets)
            res, down_res = down_res[-n:], down_res[:-n]
            if not is_final and forward_upsample_size:
                upsample_size = down_res[-1].shape[-2:]
            mkw = {"pose_feature": feats[-(i + 1)]} if self.decoder_add_posecond else None
            ckw = dict(mkw) if mkw is not None else None
            if getattr(block, "has_cross_attention", False):
                x = block(hidden_states=x, temb=temb, res_hidden_states_tuple=res, encoder_hidden_states=text,
                          upsample_size=upsample_size, attention_mask=None, cross_attention_kwargs=ckw,
                          motion_cross_attention_kwargs=mkw)
            else:
                x = block(hidden_states=x, temb=temb, res_hidden_states_tuple=res, upsample_size=upsample_size,
                          cross_attention_kwargs=ckw, motion_cross_attention_kwargs=mkw)

        Bc, Fc, Hc, Wc, C = x.dims
        no = p["norm_out"]
        n = ops.groupnorm(x.rows(), no.g, no.b, no.eps, Bc * Fc, Hc * Wc, groups=no.groups, silu=True)
        y = p["conv_out"](n.view(Bc * Fc, Hc, Wc, C))
        out = ops.from_channels_last(y.view(Bc, Fc, Hc, Wc, y.shape[-1]), C=self.conv_out.out_channels)
        return UNet3DConditionOutput(sample=out) if return_dict else (out,)
