"""The model configuration of the shipped FMC configs, as plain dicts (the reference reads them from OmegaConf yaml
and forwards them verbatim to the constructors): SD1.5 U-Net config.json + `unet_additional_kwargs`
(configs/cam.yaml:86-100), `attention_processor_kwargs` (:121-129), `pose_encoder_kwargs` (:106-120), `lora_rank` /
`lora_scale` (:103-104) and `omcm_config.params` (configs/obj.yaml:175-190)."""

SD15_UNET_CONFIG = dict(
    sample_size=64, in_channels=4, out_channels=4, center_input_sample=False, flip_sin_to_cos=True, freq_shift=0,
    down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
    mid_block_type="UNetMidBlock3DCrossAttn",
    up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
    block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, downsample_padding=1, mid_block_scale_factor=1,
    act_fn="silu", norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=768, attention_head_dim=8,
)

UNET_ADDITIONAL_KWARGS = dict(
    use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=False,
    motion_module_type="Vanilla",
    motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1,
                              attention_block_types=("Temporal_Self", "Temporal_Self"),
                              temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                              temporal_attention_dim_div=1, zero_initialize=False),
)

ATTENTION_PROCESSOR_KWARGS = dict(add_spatial=False, spatial_attn_names="attn1", add_temporal=True,
                                  temporal_attn_names="0", query_condition=True, key_value_condition=True, scale=1.0)

POSE_ENCODER_KWARGS = dict(downscale_factor=8, nums_rb=2, cin=384, ksize=1, sk=True, use_conv=False,
                           compression_factor=1, temporal_attention_nhead=8, attention_block_types=["Temporal_Self"],
                           temporal_position_encoding=True, temporal_position_encoding_max_len=16)

OMCM_KWARGS = dict(nums_rb=2, cin=832, sk=True, use_conv=False, use_pre_zero_conv=True, use_post_zero_conv=True)

LORA_KWARGS = {"lora_rank": 2, "lora_scale": 1.0}


def unet_config():
    cfg = dict(SD15_UNET_CONFIG)
    cfg.update(UNET_ADDITIONAL_KWARGS)
    return cfg


def set_processors(unet, channels):
    """train_cam_ctrl.py:230-234 with the yaml values above."""
    unet.set_all_attn_processor(add_spatial_lora=True, add_motion_lora=False, lora_kwargs=dict(LORA_KWARGS),
                                motion_lora_kwargs={"lora_rank": -1, "lora_scale": 1.0},
                                pose_feature_dimensions=list(channels), **ATTENTION_PROCESSOR_KWARGS)
