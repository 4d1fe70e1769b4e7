#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE'S OWN fmc/* SOURCES (from
/root/reference, read-only, never copied) in this container.

The reference imports diffusers==0.24.0 / decord / cv2 / imageio / nltk / omegaconf at module top; none is installable
offline.  So the third-party layer is shimmed: `diffusers.*` names resolve to the restated classes of
oracle/diffusers_restated.py (Appendix A of SURVEY.md), the data-loading libraries to empty stubs.  Everything under
fmc/ -- attention processors, motion module, 3D blocks, U-Net wiring and set_all_attn_processor, CameraPoseEncoder,
Adapter, get_traj_features_v2, the Adapted_* forwards, ray_condition -- is the reference's real code.  The vectors
therefore pin the oracle's restatement of fmc/* (tests/test_golden.py); the diffusers boundary itself stays restated
("parity unpinned at the diffusers boundary").

    python tests/golden/make_golden.py         # needs /root/reference; writes tests/golden/*.pt (a few hundred KB)

Inputs are not stored: they are regenerated from seeds by golden_inputs() on both sides; weights come from
synfmc_b200.synth.synth_init_, which seeds every parameter by its NAME, so reference, oracle and product get
identical weights without sharing code.
"""
import importlib
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TINY = dict(block_out_channels=(320, 640), down_block_types=("CrossAttnDownBlock3D", "DownBlock3D"),
            up_block_types=("UpBlock3D", "CrossAttnUpBlock3D"), layers_per_block=1)


# ---------------------------------------------------------------------------------------------------------------
# seeded inputs shared by the generator and tests/test_golden.py
# ---------------------------------------------------------------------------------------------------------------
def golden_inputs():
    from synfmc_b200 import synth
    g = torch.Generator().manual_seed(1234)
    b, f, h, w = 1, 4, 8, 8
    channels = (320, 640)
    out = {"b": b, "f": f, "h": h, "w": w, "channels": channels}
    out["sample"] = torch.randn(b, 4, f, h, w, generator=g)
    out["text"] = 0.5 * torch.randn(b, 77, 768, generator=g)
    out["pose_feats"] = [torch.randn(b, c, f, h >> l, w >> l, generator=g) for l, c in enumerate(channels)]
    out["traj_feats"] = [0.5 * torch.randn(b, c, f, h >> l, w >> l, generator=g) for l, c in enumerate(channels)]
    H, W = 64, 64
    K, c2w = synth.synth_camera(b, f, H, W, seed=11)
    out["K"], out["c2w"], out["H"], out["W"] = K, c2w, H, W
    infos, masks = synth.synth_objects(b, 2, H, W, 3, seed=12, gaussian=True)
    out["obj_infos"], out["obj_masks"] = infos, masks
    out["mm_x"] = torch.randn(2, 320, f, 3, 5, generator=g)
    out["mm_pose"] = torch.randn(2, 320, f, 3, 5, generator=g)
    return out


# ---------------------------------------------------------------------------------------------------------------
# third-party shims
# ---------------------------------------------------------------------------------------------------------------
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything:
    """Stand-in for a never-used third-party symbol."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("third-party stub called")

    def __getattr__(self, name):
        return _Anything()


def _stub_package(name, *subs):
    class _Stub(types.ModuleType):
        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            return _Anything
    m = _Stub(name)
    m.__path__ = []
    sys.modules[name] = m
    for s in subs:
        sm = _Stub(f"{name}.{s}")
        sm.__path__ = []
        sys.modules[f"{name}.{s}"] = sm
        setattr(m, s.split(".")[0], sm)
    return m


def install_shims():
    import dataclasses

    from torch import nn

    from oracle import diffusers_restated as R

    class BaseOutput(dict):
        """diffusers.utils.BaseOutput: a dataclass-style container with attribute access."""

        def __post_init__(self):
            for fld in dataclasses.fields(self):
                self[fld.name] = getattr(self, fld.name)

    class _Logger:
        def __getattr__(self, name):
            return lambda *a, **k: None

    logging = SimpleNamespace(get_logger=lambda *a, **k: _Logger())

    def register_to_config(init):
        import functools
        import inspect

        @functools.wraps(init)
        def wrapper(self, *args, **kwargs):
            sig = inspect.signature(init)
            bound = sig.bind(self, *args, **kwargs)
            bound.apply_defaults()
            cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
            init(self, *args, **kwargs)
            self.config = SimpleNamespace(**cfg)
        return wrapper

    class ConfigMixin:
        pass

    class ModelMixin(nn.Module):
        @property
        def dtype(self):
            return next(self.parameters()).dtype

        @property
        def device(self):
            return next(self.parameters()).device

        def __getattr__(self, name):
            """diffusers 0.24.0 ModelMixin.__getattr__: a missing attribute that is a config entry is served from the
            config (with a deprecation warning there) -- the reference pipelines read `unet.in_channels` this way
            (pipeline_animation_cm_om.py:630)."""
            try:
                return super().__getattr__(name)
            except AttributeError:
                cfg = self.__dict__.get("config")
                if cfg is not None and not name.startswith("_") and hasattr(cfg, name):
                    return getattr(cfg, name)
                raise

    class UNet2DConditionLoadersMixin:
        pass

    class LoRACompatibleLinear(nn.Linear):
        """diffusers 0.24.0 builds Attention.to_q/k/v/out[0] from this class when peft is absent: forward(x, scale)
        adds scale * lora_layer(x) only if a lora_layer was attached -- FMC never attaches one (its LoRA lives in
        the processors), so it is a plain Linear that tolerates the extra argument (attention_processor.py:32,51)."""

        def forward(self, hidden_states, scale=1.0):
            return super().forward(hidden_states)

    class Attention(R.Attention):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            for lin in (self.to_q, self.to_k, self.to_v, self.to_out[0]):
                lin.__class__ = LoRACompatibleLinear

    def get_activation(name):
        return {"silu": nn.SiLU(), "swish": nn.SiLU(), "mish": nn.Mish(), "gelu": nn.GELU(), "relu": nn.ReLU()}[name]

    _module("diffusers", __path__=[])
    _module("diffusers.utils", BaseOutput=BaseOutput, logging=logging, USE_PEFT_BACKEND=False,
            is_accelerate_available=lambda: False, deprecate=lambda *a, **k: None)
    _module("diffusers.models", __path__=[], AutoencoderKL=_Anything)
    _module("diffusers.models.lora", LoRALinearLayer=R.LoRALinearLayer)
    _module("diffusers.models.attention", Attention=Attention, FeedForward=R.FeedForward)
    _module("diffusers.models.attention_processor", Attention=Attention, AttentionProcessor=object,
            LoRAAttnProcessor=_Anything, SpatialNorm=_Anything)
    _module("diffusers.models.resnet", Downsample2D=R.Downsample2D, Upsample2D=R.Upsample2D,
            ResnetBlock2D=R.ResnetBlock2D)
    _module("diffusers.models.transformer_2d", Transformer2DModel=R.Transformer2DModel)
    _module("diffusers.models.embeddings", TimestepEmbedding=R.TimestepEmbedding, Timesteps=R.Timesteps)
    _module("diffusers.models.activations", get_activation=get_activation)
    _module("diffusers.models.normalization", AdaGroupNorm=_Anything)
    _module("diffusers.models.modeling_utils", ModelMixin=ModelMixin)
    _module("diffusers.configuration_utils", ConfigMixin=ConfigMixin, register_to_config=register_to_config,
            FrozenDict=dict)
    _module("diffusers.loaders", AttnProcsLayers=_Anything, UNet2DConditionLoadersMixin=UNet2DConditionLoadersMixin,
            LoraLoaderMixin=object)
    for name, subs in (("decord", ()), ("cv2", ()), ("imageio", ()), ("nltk", ("stem",)), ("omegaconf", ()),
                       ("torchvision", ("transforms", "transforms.functional"))):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub_package(name, *subs)
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)


def reference_modules():
    install_shims()
    mods = SimpleNamespace()
    mods.unet = importlib.import_module("fmc.models.unet")
    mods.unet_obj = importlib.import_module("fmc.models.unet_cam_obj")
    mods.motion = importlib.import_module("fmc.models.motion_module")
    mods.procs = importlib.import_module("fmc.models.attention_processor")
    mods.pose = importlib.import_module("fmc.models.pose_adaptor")
    mods.adapter = importlib.import_module("fmc.adapter")
    mods.util = importlib.import_module("fmc.util")
    mods.modified = importlib.import_module("fmc.modified_modules")
    try:
        mods.dataset = importlib.import_module("fmc.data.dataset")
    except Exception as e:  # the 5.6 k-line dataset module drags many data libraries; fall back to exec of the function
        mods.dataset = None
        mods.dataset_error = repr(e)
    return mods


def _ray_condition_from_source():
    """Execute just `custom_meshgrid` + `ray_condition` (fmc/data/dataset.py:922-972) out of the reference file."""
    src = open(os.path.join(REFERENCE, "fmc/data/dataset.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("def custom_meshgrid"))
    end = next(i for i, l in enumerate(src) if i > start and l.startswith("class ") or (i > start + 5 and l.startswith("def ") and "ray_condition" not in l and "custom_meshgrid" not in l))
    from packaging import version as pver
    ns = {"torch": torch, "pver": pver, "np": np}
    exec(compile("\n".join(src[start:end]), "fmc/data/dataset.py[ray_condition]", "exec"), ns)
    return ns["ray_condition"]


def main():
    from oracle import harness
    from synfmc_b200.synth import synth_init_
    torch.manual_seed(0)
    ref = reference_modules()
    inp = golden_inputs()
    out = {}

    cfg = harness.unet_config(True)

    # ---- U-Net (cam): UNet3DConditionModelPoseCond, unet.py:829-1300
    unet = ref.unet.UNet3DConditionModelPoseCond(**cfg)
    harness.set_processors(unet, cfg["block_out_channels"])
    synth_init_(unet, seed=0)
    unet.eval()
    with torch.no_grad():
        out["unet_cam"] = unet(inp["sample"], 961, inp["text"], pose_embedding_features=inp["pose_feats"]).sample.clone()
    out["unet_state_keys"] = sorted(unet.state_dict().keys())

    # ---- U-Net (cam + obj): UNet3DConditionModelCamObjCond with the trainer's forward rebinding
    # (train_cam_obj_ctrl.py:317-329 restated here: it is trainer code, not importable without its dependencies)
    unet_o = ref.unet_obj.UNet3DConditionModelCamObjCond(**cfg)
    harness.set_processors(unet_o, cfg["block_out_channels"])
    synth_init_(unet_o, seed=0)
    idx = 0
    for _n, m in unet_o.down_blocks.named_modules():
        if m.__class__.__name__ == "CrossAttnDownBlock3D":
            m.forward = ref.modified.Adapted_CrossAttnDownBlock3D_forward.__get__(m, m.__class__)
        elif m.__class__.__name__ == "DownBlock3D":
            m.forward = ref.modified.Adapted_DownBlock3D_forward.__get__(m, m.__class__)
        else:
            continue
        m.traj_fea_idx = idx
        idx += 1
    unet_o.eval()
    with torch.no_grad():
        out["unet_obj"] = unet_o(inp["sample"], 961, inp["text"], pose_embedding_features=inp["pose_feats"],
                                 traj_features=inp["traj_feats"]).sample.clone()

    # ---- motion module with CameraAdapter (motion_module.py:44-90, attention_processor.py:172-293)
    from oracle.unet import FMC_UNET_ADDITIONAL_KWARGS as KW
    mm = ref.motion.get_motion_module(320, "Vanilla", dict(KW["motion_module_kwargs"]))
    blocks = mm.temporal_transformer.transformer_blocks[0].attention_blocks
    blocks[0].set_processor(ref.procs.PoseAdaptorAttnProcessor(hidden_size=320, pose_feature_dim=320,
                                                              query_condition=True, key_value_condition=True, scale=1.0))
    blocks[1].set_processor(ref.procs.AttnProcessor())
    synth_init_(mm, seed=320)
    mm.eval()
    with torch.no_grad():
        out["motion_module"] = mm(inp["mm_x"], None, None, None,
                                  cross_attention_kwargs={"pose_feature": inp["mm_pose"]}).clone()

    # ---- rays (dataset.py:930-972) -> CameraPoseEncoder (pose_adaptor.py:159-240)
    ray_condition = ref.dataset.ray_condition if ref.dataset is not None else _ray_condition_from_source()
    b, f, H, W = inp["b"], inp["f"], inp["H"], inp["W"]
    bottom = torch.tensor([0, 0, 0, 1.0]).view(1, 1, 1, 4).expand(b, f, 1, 4)
    c2w44 = torch.cat([inp["c2w"], bottom], dim=2)
    flip = torch.zeros(f, dtype=torch.bool)
    rays = ray_condition(inp["K"], c2w44, H, W, device="cpu", flip_flag=flip)  # [b, f, H, W, 6]
    out["rays"] = rays.clone()
    enc = ref.pose.CameraPoseEncoder(channels=list(inp["channels"]), **harness.POSE_ENCODER_KWARGS)
    synth_init_(enc, seed=1)
    enc.eval()
    plucker = rays.permute(0, 4, 1, 2, 3).contiguous()  # b c f h w (train_cam_ctrl.py:585)
    with torch.no_grad():
        out["pose_encoder"] = [t.clone() for t in enc(plucker)]

    # ---- get_traj_features_v2 (util.py:147-213) -> Adapter (adapter.py:109-192)
    omcm = ref.adapter.Adapter(channels=list(inp["channels"]), **harness.OMCM_KWARGS)
    synth_init_(omcm, seed=2)
    omcm.eval()
    with torch.no_grad():
        feats = ref.util.get_traj_features_v2(inp["obj_infos"], inp["obj_masks"], omcm, False, 0.0, None, "cpu",
                                              torch.float32)
    out["traj_features"] = [t.clone() for t in feats]

    torch.save(out, os.path.join(HERE, "fmc_reference_vectors.pt"))
    size = os.path.getsize(os.path.join(HERE, "fmc_reference_vectors.pt"))
    print(f"wrote tests/golden/fmc_reference_vectors.pt ({size / 1024:.0f} KB); dataset import: "
          f"{'ok' if ref.dataset is not None else 'function exec (' + ref.dataset_error + ')'}")


if __name__ == "__main__":
    main()
