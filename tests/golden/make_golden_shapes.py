#!/usr/bin/env python
"""State-dict layout (every key with its shape) of the FULL-SIZE reference modules -- the 4-level SD1.5-shaped U-Net of
both variants with the shipped processors, CameraPoseEncoder, Adapter -- built from the reference's own classes on the
meta device (no memory).  Writes tests/golden/fmc_reference_shapes.json.gz; tests/test_shapes.py holds the mirror and
the oracle to it, so checkpoints interchange at full size and not only on the 2-level test U-Net.

    python tests/golden/make_golden_shapes.py      # needs /root/reference
"""
import gzip
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

CHANNELS = [320, 640, 1280, 1280]


def layout(module):
    return {k: list(v.shape) for k, v in module.state_dict().items()}


def build_all(unet_cam_cls, unet_obj_cls, pose_cls, adapter_cls):
    """the four modules at the shipped configuration, on the meta device"""
    from oracle import harness
    cfg = harness.unet_config(False)
    out = {}
    with torch.device("meta"):
        for name, cls in (("unet_cam", unet_cam_cls), ("unet_obj", unet_obj_cls)):
            u = cls(**cfg)
            harness.set_processors(u, cfg["block_out_channels"])
            out[name] = layout(u)
        out["pose_encoder"] = layout(pose_cls(channels=list(CHANNELS), **harness.POSE_ENCODER_KWARGS))
        out["adapter"] = layout(adapter_cls(channels=list(CHANNELS), **harness.OMCM_KWARGS))
    return out


def main():
    import make_golden as mg
    ref = mg.reference_modules()
    shapes = build_all(ref.unet.UNet3DConditionModelPoseCond, ref.unet_obj.UNet3DConditionModelCamObjCond,
                       ref.pose.CameraPoseEncoder, ref.adapter.Adapter)
    with gzip.open(os.path.join(HERE, "fmc_reference_shapes.json.gz"), "wt") as fh:
        json.dump(shapes, fh, sort_keys=True)
    for name, lay in shapes.items():
        n = sum(int(torch.Size(s).numel()) for s in lay.values())
        print(f"{name}: {len(lay)} tensors, {n / 1e6:.1f} M elements")


if __name__ == "__main__":
    main()
