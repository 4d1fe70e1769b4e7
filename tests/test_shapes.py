"""Full-size state-dict layouts (every key and shape of the 4-level U-Net of both variants with the shipped processors,
CameraPoseEncoder, Adapter) of the mirror and of the oracle against the reference's own classes
(tests/golden/fmc_reference_shapes.json.gz, built on the meta device by make_golden_shapes.py)."""
import gzip
import json
import os

import pytest

from tests.golden.make_golden_shapes import build_all

GOLD = json.load(gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fmc_reference_shapes.json.gz"), "rt"))


def _diff(got, want):
    missing = sorted(set(want) - set(got))
    extra = sorted(set(got) - set(want))
    wrong = sorted(k for k in set(got) & set(want) if got[k] != want[k])
    return missing[:5], extra[:5], [(k, got[k], want[k]) for k in wrong[:5]]


@pytest.mark.parametrize("which", ["mirror", "oracle"])
def test_full_size_state_dict_layout_equals_the_reference(which):
    if which == "mirror":
        from synfmc_b200.fmc.adapter import Adapter
        from synfmc_b200.fmc.models.pose_adaptor import CameraPoseEncoder
        from synfmc_b200.fmc.models.unet import UNet3DConditionModelPoseCond
        from synfmc_b200.fmc.models.unet_cam_obj import UNet3DConditionModelCamObjCond
    else:
        from oracle.adapter import Adapter
        from oracle.pose_adaptor import CameraPoseEncoder
        from oracle.unet import UNet3DConditionModelCamObjCond, UNet3DConditionModelPoseCond
    got = build_all(UNet3DConditionModelPoseCond, UNet3DConditionModelCamObjCond, CameraPoseEncoder, Adapter)
    for name in ("unet_cam", "unet_obj", "pose_encoder", "adapter"):
        assert _diff(got[name], GOLD[name]) == ([], [], []), name
        assert len(got[name]) == len(GOLD[name])


def test_parameter_budget_of_the_trainable_sets():
    """Domain-LoRA 96.3 M ('rank' = C / 2), CameraAdapter qkv_merge 19.0 M over 20 modules, CameraPoseEncoder
    199.3 M, Adapter 152.5 M -- computed from the reference's layout."""
    def numel(shape):
        n = 1
        for s in shape:
            n *= s
        return n
    unet = GOLD["unet_obj"]
    lora = sum(numel(s) for k, s in unet.items() if "lora" in k)
    merge = sum(numel(s) for k, s in unet.items() if "merge" in k and "lora" not in k)
    assert round(lora / 1e6, 1) == 96.3 and round(merge / 1e6, 1) == 19.0
    assert len([k for k in unet if k.endswith("qkv_merge.weight")]) == 20
    assert round(sum(numel(s) for k, s in GOLD["pose_encoder"].items() if not k.endswith(".pe")) / 1e6, 1) == 199.3
    assert round(sum(numel(s) for s in GOLD["adapter"].values()) / 1e6, 1) == 152.5
