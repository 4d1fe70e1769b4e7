#!/usr/bin/env python
"""Record the call signatures of every reference symbol the trainers touch on the hot path by importing the
reference's own modules (shims of make_golden.py / make_golden_pipeline.py for the third-party layer).  Writes
tests/golden/fmc_reference_surface.json; tests/test_surface.py holds the mirror package to it.

    python tests/golden/make_golden_surface.py      # needs /root/reference
"""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

# (module under fmc, qualified name inside it)
SYMBOLS = [
    ("models.unet", "UNet3DConditionModelPoseCond.forward"),
    ("models.unet", "UNet3DConditionModelPoseCond.set_all_attn_processor"),
    ("models.unet", "UNet3DConditionModelPoseCond.from_pretrained_2d"),
    ("models.unet_cam_obj", "UNet3DConditionModelCamObjCond.forward"),
    ("models.unet_cam_obj", "UNet3DConditionModelCamObjCond.set_all_attn_processor"),
    ("models.unet_cam_obj", "UNet3DConditionModelCamObjCond.from_pretrained_2d"),
    ("models.pose_adaptor", "CameraPoseEncoder.__init__"),
    ("models.pose_adaptor", "CameraPoseEncoder.forward"),
    ("models.pose_adaptor", "PoseAdaptor.__init__"),
    ("models.pose_adaptor", "PoseAdaptor.forward"),
    ("models.pose_obj_adaptor", "CamObjPoseAdaptor.__init__"),
    ("models.pose_obj_adaptor", "CamObjPoseAdaptor.forward"),
    ("models.attention_processor", "AttnProcessor.__call__"),
    ("models.attention_processor", "LoRAAttnProcessor.__init__"),
    ("models.attention_processor", "PoseAdaptorAttnProcessor.__init__"),
    ("models.motion_module", "get_motion_module"),
    ("adapter", "Adapter.__init__"),
    ("adapter", "Adapter.forward"),
    ("util", "get_traj_features_v2"),
    ("modified_modules", "Adapted_CrossAttnDownBlock3D_forward"),
    ("modified_modules", "Adapted_DownBlock3D_forward"),
    ("data.dataset", "ray_condition"),
    ("data.utils", "create_absolute_matrix_from_ref_cam_list"),
    ("data.utils", "create_relative_matrix_of_cam_list"),
    ("data.utils", "create_relative_matrix_of_two_torch_matrix"),
    ("utils.util", "setup_logger"),
    ("utils.util", "format_time"),
    ("utils.util", "save_videos_grid"),
    ("utils.util", "instantiate_from_config"),
]


def describe(fn):
    fn = getattr(fn, "__func__", fn)
    fn = inspect.unwrap(fn)
    out = []
    for name, q in inspect.signature(fn).parameters.items():
        if name in ("self", "cls"):
            continue
        default = None if q.default is inspect.Parameter.empty else repr(q.default)
        out.append([name, default, q.kind.name])
    return out


def resolve(module, qualname):
    obj = module
    for part in qualname.split("."):
        obj = obj.__dict__[part] if inspect.isclass(obj) and part in obj.__dict__ else getattr(obj, part)
    return obj


def main():
    import importlib
    import types

    import make_golden as mg
    mg.reference_modules()
    sys.modules.setdefault("termcolor", types.SimpleNamespace(colored=lambda s, *a, **k: s))
    surface = {}
    for mod, qual in SYMBOLS:
        try:
            m = importlib.import_module("fmc." + mod)
            surface[f"{mod}:{qual}"] = describe(resolve(m, qual))
        except Exception as e:  # the 5.6 k-line dataset module needs data libraries: take the function from its source
            if (mod, qual) == ("data.dataset", "ray_condition"):
                surface[f"{mod}:{qual}"] = describe(mg._ray_condition_from_source())
            else:
                raise RuntimeError(f"{mod}:{qual}: {e!r}")
    with open(os.path.join(HERE, "fmc_reference_surface.json"), "w") as fh:
        json.dump(surface, fh, indent=1, sort_keys=True)
    print(f"recorded {len(surface)} signatures")
    # the constructor / processor / scheduler sections of the shipped configs, verbatim
    import yaml
    keep = ("unet_additional_kwargs", "lora_rank", "lora_scale", "pose_encoder_kwargs", "attention_processor_kwargs",
            "noise_scheduler_kwargs", "omcm_config", "validation_data")
    configs = {}
    for name in ("cam", "obj"):
        y = yaml.safe_load(open(os.path.join(mg.REFERENCE, "configs", f"{name}.yaml")))
        configs[name] = {k: y[k] for k in keep if k in y}
    with open(os.path.join(HERE, "fmc_reference_configs.json"), "w") as fh:
        json.dump(configs, fh, indent=1, sort_keys=True)
    print({k: sorted(v) for k, v in configs.items()})


if __name__ == "__main__":
    main()
