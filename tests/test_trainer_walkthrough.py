"""The model-setup section of train_cam_obj_ctrl.py (:236-400), statement by statement, against the mirror classes on
CPU (construction, processors, checkpoint loading, forward rebinding, trainable-set selection -- no kernel runs).
A 4-level U-Net with narrow channels keeps it light; the trainers' logic does not depend on the widths."""
import json
import os

import pytest
import torch

from oracle import harness as helpers
from synfmc_b200 import workload_config as wc

CHANNELS = [32, 64, 64, 64]


def _sd15_like_dir(tmp_path):
    """a `unet/` folder shaped like the SD1.5 one the trainers point `from_pretrained_2d` at (2-D block names,
    diffusers private keys), with weights for the 2-D part only"""
    from synfmc_b200.fmc.models.unet_cam_obj import UNet3DConditionModelCamObjCond
    cfg = dict(wc.SD15_UNET_CONFIG)
    cfg.update(block_out_channels=CHANNELS, cross_attention_dim=64)
    cfg["down_block_types"] = ["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"]
    cfg["up_block_types"] = ["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3
    on_disk = dict(cfg, _class_name="UNet2DConditionModel", _diffusers_version="0.6.0")
    path = tmp_path / "sd15" / "unet"
    os.makedirs(path)
    json.dump(on_disk, open(path / "config.json", "w"))
    torch.manual_seed(0)
    cfg3d = dict(cfg)
    cfg3d["down_block_types"] = ["CrossAttnDownBlock3D"] * 3 + ["DownBlock3D"]
    cfg3d["up_block_types"] = ["UpBlock3D"] + ["CrossAttnUpBlock3D"] * 3
    donor = UNet3DConditionModelCamObjCond(**cfg3d, **wc.UNET_ADDITIONAL_KWARGS)
    two_d = {k: v.clone() for k, v in donor.state_dict().items() if "motion_modules" not in k}
    torch.save(two_d, path / "diffusion_pytorch_model.bin")
    return str(tmp_path / "sd15"), two_d


def test_train_cam_obj_ctrl_model_setup(tmp_path):
    from synfmc_b200.fmc.adapter import Adapter
    from synfmc_b200.fmc.models.attention_processor import AttnProcessor as CustomizedAttnProcessor
    from synfmc_b200.fmc.models.pose_adaptor import CameraPoseEncoder
    from synfmc_b200.fmc.models.pose_obj_adaptor import CamObjPoseAdaptor
    from synfmc_b200.fmc.models.unet_cam_obj import UNet3DConditionModelCamObjCond
    from synfmc_b200.fmc.modified_modules import Adapted_CrossAttnDownBlock3D_forward, Adapted_DownBlock3D_forward

    pretrained_model_path, two_d = _sd15_like_dir(tmp_path)
    # :236-238
    unet = UNet3DConditionModelCamObjCond.from_pretrained_2d(pretrained_model_path, subfolder="unet",
                                                             unet_additional_kwargs=dict(wc.UNET_ADDITIONAL_KWARGS))
    sd = unet.state_dict()
    assert all(torch.equal(sd[k], v) for k, v in two_d.items())
    assert any("motion_modules" in k for k in sd)
    pose_encoder = CameraPoseEncoder(channels=CHANNELS, **wc.POSE_ENCODER_KWARGS)
    # :242-249
    unet.requires_grad_(False)
    lora_rank, lora_scale = 2, 1.0
    unet.set_all_attn_processor(add_spatial_lora=True, add_motion_lora=False,
                                lora_kwargs={"lora_rank": lora_rank, "lora_scale": lora_scale},
                                motion_lora_kwargs={"lora_rank": -1, "lora_scale": 1.0},
                                pose_feature_dimensions=CHANNELS, **wc.ATTENTION_PROCESSOR_KWARGS)
    # :252-264 image-LoRA checkpoint: a dict of `...processor.to_*_lora.{down,up}.weight` keys, loaded non-strictly
    lora_ckpt = {"lora_state_dict": {k: torch.randn_like(v) for k, v in unet.state_dict().items() if "lora" in k}}
    assert lora_ckpt["lora_state_dict"]
    _, lora_u = unet.load_state_dict(lora_ckpt["lora_state_dict"], strict=False)
    assert len(lora_u) == 0
    # :267-276 motion-module checkpoint saved from a DDP-wrapped model
    mm_ckpt = {"module." + k: torch.randn_like(v) for k, v in unet.state_dict().items()
               if "motion_modules" in k and "processor" not in k}
    mm_ckpt = {k.replace("module.", ""): v for k, v in mm_ckpt.items()}
    _, mm_u = unet.load_state_dict(mm_ckpt, strict=False)
    assert len(mm_u) == 0
    key = "down_blocks.0.motion_modules.0.temporal_transformer.proj_in.weight"
    assert torch.equal(unet.state_dict()[key], mm_ckpt[key])
    # :278-292 CMCM checkpoint: pose encoder + the `qkv_merge` processors
    pose_adaptor = CamObjPoseAdaptor(unet, pose_encoder)
    ckpt = {"pose_encoder_state_dict": {k: v.clone() for k, v in pose_encoder.state_dict().items()},
            "attention_processor_state_dict": {k: torch.randn_like(v) for k, v in unet.state_dict().items() if "qkv_merge" in k}}
    pose_enc_m, pose_enc_u = pose_adaptor.pose_encoder.load_state_dict(ckpt["pose_encoder_state_dict"], strict=False)
    assert len(pose_enc_m) == 0 and len(pose_enc_u) == 0
    _, attention_processor_u = pose_adaptor.unet.load_state_dict(ckpt["attention_processor_state_dict"], strict=False)
    assert len(attention_processor_u) == 0 and pose_adaptor.unet is unet and pose_adaptor.pose_encoder is pose_encoder
    # :295-312 ObjectEncoder, checkpoint keys carrying a DDP prefix
    omcm = Adapter(channels=CHANNELS, **wc.OMCM_KWARGS)
    omcm_state_dict = {("module." + k).replace("module.", ""): v for k, v in omcm.state_dict().items()}
    m, u = omcm.load_state_dict(omcm_state_dict, strict=True)
    assert len(m) == 0 and len(u) == 0
    # :317-329 forward rebinding in named_modules() order
    idx = 0
    bound = []
    for _name, _module in unet.down_blocks.named_modules():
        if _module.__class__.__name__ == "CrossAttnDownBlock3D":
            setattr(_module, "forward", Adapted_CrossAttnDownBlock3D_forward.__get__(_module, _module.__class__))
            setattr(_module, "traj_fea_idx", idx)
            bound.append((_name, idx))
            idx += 1
        elif _module.__class__.__name__ == "DownBlock3D":
            setattr(_module, "forward", Adapted_DownBlock3D_forward.__get__(_module, _module.__class__))
            setattr(_module, "traj_fea_idx", idx)
            bound.append((_name, idx))
            idx += 1
    assert bound == [("0", 0), ("1", 1), ("2", 2), ("3", 3)]
    assert unet.down_blocks[3].forward.__func__ is Adapted_DownBlock3D_forward
    # :331-362 trainable set of the camera stage
    pose_encoder.requires_grad_(False)
    spatial = torch.nn.ModuleList([v for v in unet.attn_processors.values() if not isinstance(v, CustomizedAttnProcessor)])
    temporal = torch.nn.ModuleList([v for v in unet.mm_attn_processors.values() if not isinstance(v, CustomizedAttnProcessor)])
    assert len(spatial) == 32 and len(temporal) == 20   # 16 Transformer2D x (attn1, attn2); 20 motion modules x block 0
    spatial.requires_grad_(True)
    temporal.requires_grad_(True)
    pose_encoder.requires_grad_(True)
    for n, p in spatial.named_parameters():
        if "lora" in n:
            p.requires_grad = False
    attention_trainable = [k for k, v in unet.named_parameters() if v.requires_grad and "merge" in k and "lora" not in k]
    assert len(attention_trainable) == 40 and all(k.endswith(("qkv_merge.weight", "qkv_merge.bias")) for k in attention_trainable)
    assert not any(v.requires_grad for k, v in unet.named_parameters() if "lora" in k)
    # :366-381 motion-module trainables selected by module class name
    mm_param_names = []
    for _name, _module in unet.named_modules():
        if _module.__class__.__name__ == "TemporalTransformer3DModel":
            mm_param_names += [f"{_name}.norm", f"{_name}.proj_in", f"{_name}.proj_out"]
    assert len(mm_param_names) == 60
    mm_params = [p for n, p in unet.named_parameters() if any(t in n for t in mm_param_names)]
    assert len(mm_params) == 120   # weight + bias of norm / proj_in / proj_out in 20 modules
    # :386-391
    assert len(list(omcm.parameters())) == len(omcm.state_dict()) > 0
    # what the pipeline reads (pipeline_animation_cm_om.py:630)
    assert unet.in_channels == 4 and tuple(unet.config.block_out_channels) == tuple(CHANNELS)
    assert unet.config.cross_attention_dim == 64 and unet.device.type == "cpu"


def test_train_cam_ctrl_model_setup_and_checkpoint_round_trip(tmp_path):
    """train_cam_ctrl.py:225-300 (CMC stage) plus the checkpoint it writes (:672-687: trainable keys only) loaded back
    the way train_cam_obj_ctrl.py:281-292 does."""
    from synfmc_b200.fmc.models.attention_processor import AttnProcessor as CustomizedAttnProcessor
    from synfmc_b200.fmc.models.pose_adaptor import CameraPoseEncoder, PoseAdaptor
    from synfmc_b200.fmc.models.unet import UNet3DConditionModelPoseCond

    pretrained_model_path, _ = _sd15_like_dir(tmp_path)
    unet = UNet3DConditionModelPoseCond.from_pretrained_2d(pretrained_model_path, subfolder="unet",
                                                           unet_additional_kwargs=dict(wc.UNET_ADDITIONAL_KWARGS))
    pose_encoder = CameraPoseEncoder(channels=CHANNELS, **wc.POSE_ENCODER_KWARGS)
    unet.set_all_attn_processor(add_spatial_lora=True, add_motion_lora=False,
                                lora_kwargs={"lora_rank": 2, "lora_scale": 1.0},
                                motion_lora_kwargs={"lora_rank": -1, "lora_scale": 1.0},
                                pose_feature_dimensions=CHANNELS, **wc.ATTENTION_PROCESSOR_KWARGS)
    unet.requires_grad_(False)
    spatial = torch.nn.ModuleList([v for v in unet.attn_processors.values() if not isinstance(v, CustomizedAttnProcessor)])
    temporal = torch.nn.ModuleList([v for v in unet.mm_attn_processors.values() if not isinstance(v, CustomizedAttnProcessor)])
    spatial.requires_grad_(True)
    temporal.requires_grad_(True)
    pose_encoder.requires_grad_(True)
    for n, p in spatial.named_parameters():
        if "lora" in n:
            p.requires_grad = False
    pose_adaptor = PoseAdaptor(unet, pose_encoder)
    encoder_names = [n for n, p in pose_encoder.named_parameters() if p.requires_grad]
    attention_names = [k for k, v in unet.named_parameters() if v.requires_grad and "merge" in k and "lora" not in k]
    assert len(attention_names) == 40 and len(encoder_names) == len(list(pose_encoder.parameters()))
    # every parameter that requires grad is in one of the two lists: nothing else of the U-Net is trainable
    assert {k for k, v in unet.named_parameters() if v.requires_grad} == set(attention_names)
    # the wrapper exposes what `DDP(pose_adaptor).module.<x>` is asked for (:480-488, :672-687)
    assert pose_adaptor.unet is unet and pose_adaptor.pose_encoder is pose_encoder
    with torch.no_grad():
        for k, v in unet.named_parameters():
            if k in attention_names:
                v.normal_()
    state = {"pose_encoder_state_dict": pose_adaptor.pose_encoder.state_dict(),
             "attention_processor_state_dict": {k: v for k, v in unet.state_dict().items() if k in attention_names}}
    torch.save(state, tmp_path / "cmcm.ckpt")
    ckpt = torch.load(tmp_path / "cmcm.ckpt", map_location="cpu")
    fresh = UNet3DConditionModelPoseCond.from_pretrained_2d(pretrained_model_path, subfolder="unet",
                                                            unet_additional_kwargs=dict(wc.UNET_ADDITIONAL_KWARGS))
    fresh.set_all_attn_processor(add_spatial_lora=True, add_motion_lora=False,
                                 lora_kwargs={"lora_rank": 2, "lora_scale": 1.0},
                                 motion_lora_kwargs={"lora_rank": -1, "lora_scale": 1.0},
                                 pose_feature_dimensions=CHANNELS, **wc.ATTENTION_PROCESSOR_KWARGS)
    _, unexpected = fresh.load_state_dict(ckpt["attention_processor_state_dict"], strict=False)
    assert len(unexpected) == 0
    for k in attention_names:
        assert torch.equal(fresh.state_dict()[k], unet.state_dict()[k])


def _shipped(name):
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fmc_reference_configs.json")
    return json.load(open(path))[name]


def test_shipped_yaml_sections_construct_the_mirror_verbatim(tmp_path):
    """configs/cam.yaml / obj.yaml sections (tests/golden/fmc_reference_configs.json, copied from the reference by
    make_golden_surface.py) go into the constructors exactly as the trainers pass them: `unet_additional_kwargs` with
    its two entries nothing reads (dropped and reported like diffusers' from_config does), `pose_encoder_kwargs`,
    `attention_processor_kwargs`, `omcm_config.params`, `noise_scheduler_kwargs`.  Only the channel widths are
    narrowed to keep the test light."""
    from synfmc_b200.fmc._blocks import DDIMScheduler
    from synfmc_b200.fmc.adapter import Adapter
    from synfmc_b200.fmc.models.pose_adaptor import CameraPoseEncoder
    from synfmc_b200.fmc.models.unet_cam_obj import UNet3DConditionModelCamObjCond
    y = _shipped("obj")
    assert y["unet_additional_kwargs"] == _shipped("cam")["unet_additional_kwargs"]
    pretrained_model_path, _ = _sd15_like_dir(tmp_path)
    seen = []
    logger = type("L", (), {"info": lambda self, m: seen.append(m)})()
    unet = UNet3DConditionModelCamObjCond.from_pretrained_2d(pretrained_model_path, subfolder="unet",
                                                             unet_additional_kwargs=y["unet_additional_kwargs"], logger=logger)
    assert any(m.startswith("unet_use_cross_frame_attention") for m in seen)
    assert any(m.startswith("unet_use_temporal_attention") for m in seen)
    apk = dict(y["attention_processor_kwargs"], pose_feature_dimensions=CHANNELS)
    unet.set_all_attn_processor(add_spatial_lora=True, add_motion_lora=False,
                                lora_kwargs={"lora_rank": y["lora_rank"], "lora_scale": y["lora_scale"]},
                                motion_lora_kwargs={"lora_rank": -1, "lora_scale": 1.0}, **apk)
    assert len(unet.mm_attn_processors) == 40
    CameraPoseEncoder(**dict(y["pose_encoder_kwargs"], channels=CHANNELS))
    Adapter(**dict(y["omcm_config"]["params"], channels=CHANNELS))
    sched = DDIMScheduler(**y["noise_scheduler_kwargs"])
    sched.set_timesteps(y["validation_data"]["num_inference_steps"])
    assert sched.timesteps.tolist()[:3] == [961, 921, 881] and y["validation_data"]["guidance_scale"] == 8.0
    with pytest.raises(RuntimeError):
        UNet3DConditionModelCamObjCond.from_pretrained_2d(str(tmp_path / "nowhere"), subfolder="unet")

    # the constants the bench / tests build the workload from are these yaml values
    uak = {k: v for k, v in y["unet_additional_kwargs"].items() if not k.startswith("unet_use_")}
    norm = lambda d: json.loads(json.dumps(d))   # tuples -> lists  # noqa: E731
    assert norm(wc.UNET_ADDITIONAL_KWARGS) == uak
    assert norm(wc.POSE_ENCODER_KWARGS) == {k: v for k, v in y["pose_encoder_kwargs"].items() if k != "channels"}
    assert norm(wc.ATTENTION_PROCESSOR_KWARGS) == {k: v for k, v in y["attention_processor_kwargs"].items()
                                                  if k != "pose_feature_dimensions"}
    assert norm(wc.OMCM_KWARGS) == {k: v for k, v in y["omcm_config"]["params"].items() if k != "channels"}
    assert wc.LORA_KWARGS == {"lora_rank": y["lora_rank"], "lora_scale": y["lora_scale"]}
