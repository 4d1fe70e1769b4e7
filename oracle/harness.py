"""Shared builders for the parity tests: an oracle model and the product model with identical (bf16-exact) weights.
The product is filled through load_state_dict(strict=True) from the oracle's state dict, which is also the
state-dict key contract check."""
import torch

from oracle.adapter import Adapter as OAdapter
from oracle.modified_modules import bind_omcm_forwards as o_bind
from oracle.pose_adaptor import CameraPoseEncoder as OCameraPoseEncoder
from oracle.unet import (FMC_UNET_ADDITIONAL_KWARGS, SD15_UNET_CONFIG, UNet3DConditionModelCamObjCond as OUNetObj,
                         UNet3DConditionModelPoseCond as OUNetCam)
from synfmc_b200.synth import synth_init_

# configs/cam.yaml:121-129 + train_cam_ctrl.py:230-234
ATTN_PROC_KWARGS = dict(add_spatial=False, spatial_attn_names="attn1", add_temporal=True, temporal_attn_names="0",
                        query_condition=True, key_value_condition=True, scale=1.0)
POSE_ENCODER_KWARGS = dict(downscale_factor=8, nums_rb=2, cin=384, ksize=1, sk=True, use_conv=False,
                           compression_factor=1, temporal_attention_nhead=8, attention_block_types=["Temporal_Self"],
                           temporal_position_encoding=True, temporal_position_encoding_max_len=16)
OMCM_KWARGS = dict(nums_rb=2, cin=832, sk=True, use_conv=False, use_pre_zero_conv=True, use_post_zero_conv=True)

TINY = dict(block_out_channels=(320, 640), down_block_types=("CrossAttnDownBlock3D", "DownBlock3D"),
            up_block_types=("UpBlock3D", "CrossAttnUpBlock3D"), layers_per_block=1)


def unet_config(tiny=True):
    cfg = dict(SD15_UNET_CONFIG)
    cfg.update(FMC_UNET_ADDITIONAL_KWARGS)
    if tiny:
        cfg.update(TINY)
    return cfg


def set_processors(unet, channels):
    unet.set_all_attn_processor(add_spatial_lora=True, add_motion_lora=False,
                                lora_kwargs={"lora_rank": 2, "lora_scale": 1.0},
                                motion_lora_kwargs={"lora_rank": -1, "lora_scale": 1.0},
                                pose_feature_dimensions=list(channels), **ATTN_PROC_KWARGS)


def build_oracle_unet(tiny=True, obj=False, seed=0):
    cfg = unet_config(tiny)
    unet = (OUNetObj if obj else OUNetCam)(**cfg)
    set_processors(unet, cfg["block_out_channels"])
    synth_init_(unet, seed=seed)
    if obj:
        o_bind(unet)
    return unet.eval()


def build_product_unet(oracle_unet, tiny=True, obj=False, device="cuda"):
    from synfmc_b200.fmc.models.unet import UNet3DConditionModelPoseCond
    from synfmc_b200.fmc.models.unet_cam_obj import UNet3DConditionModelCamObjCond
    from synfmc_b200.fmc.modified_modules import bind_omcm_forwards
    cfg = unet_config(tiny)
    unet = (UNet3DConditionModelCamObjCond if obj else UNet3DConditionModelPoseCond)(**cfg)
    set_processors(unet, cfg["block_out_channels"])
    missing, unexpected = unet.load_state_dict(oracle_unet.state_dict(), strict=True)
    assert not missing and not unexpected
    if obj:
        bind_omcm_forwards(unet)
    return unet.to(device).eval().requires_grad_(False)  # inference: the mirror raises under autograd (forward only)


def build_oracle_pose_encoder(channels, seed=1):
    enc = OCameraPoseEncoder(channels=list(channels), **POSE_ENCODER_KWARGS)
    synth_init_(enc, seed=seed)
    return enc.eval()


def build_product_pose_encoder(oracle_enc, channels, device="cuda"):
    from synfmc_b200.fmc.models.pose_adaptor import CameraPoseEncoder
    enc = CameraPoseEncoder(channels=list(channels), **POSE_ENCODER_KWARGS)
    enc.load_state_dict(oracle_enc.state_dict(), strict=True)
    return enc.to(device).eval().requires_grad_(False)


def build_oracle_omcm(channels, seed=2):
    m = OAdapter(channels=list(channels), **OMCM_KWARGS)
    synth_init_(m, seed=seed)
    return m.eval()


def build_product_omcm(oracle_m, channels, device="cuda"):
    from synfmc_b200.fmc.adapter import Adapter
    m = Adapter(channels=list(channels), **OMCM_KWARGS)
    m.load_state_dict(oracle_m.state_dict(), strict=True)
    return m.to(device).eval().requires_grad_(False)


def rel_l2(got, want):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    return float((got - want).norm() / want.norm().clamp_min(1e-30))
