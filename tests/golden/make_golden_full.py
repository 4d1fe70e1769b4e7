#!/usr/bin/env python
"""Golden output of the FULL-DEPTH reference U-Net (4 levels, 2 layers per block, mid block, all 16 Transformer2D and 20
motion modules, shipped processors, trainer-bound object-feature injection) at the latent size of BASELINE config 1 (32x32) with 4 frames, by
executing the reference's own classes -- the 2-level U-Net of make_golden.py does not exercise level 2 / 3, the mid block
or the three-resnet up blocks.  Writes tests/golden/fmc_reference_full_unet.pt (a few KB).

    python tests/golden/make_golden_full.py      # needs /root/reference, ~12 GB of RAM, a few minutes
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

CHANNELS = (320, 640, 1280, 1280)


def full_inputs():
    g = torch.Generator().manual_seed(777)
    b, f, h, w = 1, 4, 32, 32   # the level sizes of BASELINE config 1 (32, 16, 8, 4)
    inp = {"sample": torch.randn(b, 4, f, h, w, generator=g), "text": 0.5 * torch.randn(b, 77, 768, generator=g)}
    inp["pose_feats"] = [torch.randn(b, c, f, h >> l, w >> l, generator=g) for l, c in enumerate(CHANNELS)]
    inp["traj_feats"] = [0.5 * torch.randn(b, c, f, h >> l, w >> l, generator=g) for l, c in enumerate(CHANNELS)]
    return inp


def main():
    import make_golden as mg
    from oracle import harness
    from synfmc_b200.synth import synth_init_
    ref = mg.reference_modules()
    cfg = harness.unet_config(False)
    t0 = time.time()
    unet = ref.unet_obj.UNet3DConditionModelCamObjCond(**cfg)
    harness.set_processors(unet, cfg["block_out_channels"])
    synth_init_(unet, seed=0)
    idx = 0
    for _n, m in unet.down_blocks.named_modules():   # train_cam_obj_ctrl.py:317-329
        if m.__class__.__name__ == "CrossAttnDownBlock3D":
            m.forward = ref.modified.Adapted_CrossAttnDownBlock3D_forward.__get__(m, m.__class__)
        elif m.__class__.__name__ == "DownBlock3D":
            m.forward = ref.modified.Adapted_DownBlock3D_forward.__get__(m, m.__class__)
        else:
            continue
        m.traj_fea_idx = idx
        idx += 1
    unet.eval()
    t1 = time.time()
    inp = full_inputs()
    out = {}
    with torch.no_grad():
        out["unet_obj_full"] = unet(inp["sample"], 961, inp["text"], pose_embedding_features=inp["pose_feats"],
                                    traj_features=inp["traj_feats"]).sample.clone()
        out["unet_obj_full_no_traj"] = unet(inp["sample"], 41, inp["text"], pose_embedding_features=inp["pose_feats"],
                                            traj_features=None).sample.clone()
    # ---- the two encoders at their full four-level configuration (make_golden.py runs them with two levels)
    from make_golden import golden_inputs
    gi = golden_inputs()
    ray_condition = ref.dataset.ray_condition if ref.dataset is not None else mg._ray_condition_from_source()
    b, f, H, W = gi["b"], gi["f"], gi["H"], gi["W"]
    bottom = torch.tensor([0, 0, 0, 1.0]).view(1, 1, 1, 4).expand(b, f, 1, 4)
    rays = ray_condition(gi["K"], torch.cat([gi["c2w"], bottom], dim=2), H, W, device="cpu",
                         flip_flag=torch.zeros(f, dtype=torch.bool))
    enc = ref.pose.CameraPoseEncoder(channels=list(CHANNELS), **harness.POSE_ENCODER_KWARGS)
    synth_init_(enc, seed=1)
    enc.eval()
    omcm = ref.adapter.Adapter(channels=list(CHANNELS), **harness.OMCM_KWARGS)
    synth_init_(omcm, seed=2)
    omcm.eval()
    with torch.no_grad():
        out["pose_encoder_full_c16"] = [t[:, ::16].clone() for t in enc(rays.permute(0, 4, 1, 2, 3).contiguous())]
        feats = ref.util.get_traj_features_v2(gi["obj_infos"], gi["obj_masks"], omcm, False, 0.0, None, "cpu", torch.float32)
        out["traj_features_full_c16"] = [t[:, ::16].clone() for t in feats]
    torch.save(out, os.path.join(HERE, "fmc_reference_full_unet.pt"))
    print(f"build + init {t1 - t0:.0f} s, two forwards {time.time() - t1:.0f} s;", {k: (tuple(v.shape) if torch.is_tensor(v) else [tuple(t.shape) for t in v]) for k, v in out.items()},
          float(out["unet_obj_full"].std()))


if __name__ == "__main__":
    main()
