"""Mirror of fmc/util.py:147-213 `get_traj_features_v2`: builds the ObjectEncoder input (6-D object pose broadcast x
Gaussian-mask scatter) and runs the ObjectEncoder.  The reference's Python triple loop with boolean-mask indexing
(:161-182) is one kernel here (fmc_traj_scatter_unshuffle_bf16): last object with mask > 0 wins per pixel, channels
(info*m)*m and m*m, written directly in the PixelUnshuffle(8) channels-last layout the Adapter consumes."""
import random

import numpy as np
import torch

from .. import engine, ops
from ..engine import CL
from .models.pose_adaptor import unshuffle8_to_cl


def pack_objects(obj_info_list_list, obj_mask_list_list, device):
    """Nested lists (clip -> frame -> [n_obj, 12] numpy / [n_obj, 1, H, W] tensor) -> dense fp32 device tensors
    info [B, F, n_max, 12], masks [B, F, n_max, H, W]; missing objects are zero masks (never selected)."""
    assert len(obj_info_list_list) == len(obj_mask_list_list)
    B, Fn = len(obj_info_list_list), len(obj_info_list_list[0])
    H, W = obj_mask_list_list[0][0].shape[-2:]
    n_max = max(int(np.asarray(i).shape[0]) for clip in obj_info_list_list for i in clip)
    info = torch.zeros(B, Fn, n_max, 12, dtype=torch.float32)
    masks = torch.zeros(B, Fn, n_max, H, W, dtype=torch.float32)
    for b, (infos, ms) in enumerate(zip(obj_info_list_list, obj_mask_list_list)):
        for f, (obj_info, obj_mask) in enumerate(zip(infos, ms)):
            oi = torch.from_numpy(np.asarray(obj_info)).to(torch.float32)
            om = torch.as_tensor(obj_mask).to(torch.float32)
            info[b, f, :oi.shape[0]] = oi
            masks[b, f, :om.shape[0]] = om[:, 0]
    return info.to(device), masks.to(device)


def traj_features_cl(info, masks, omcm, null_clips=None):
    """info [B, F, n, 12], masks [B, F, n, H, W] (device fp32) -> 4 CL features [B, F, h_l, w_l, C_l]."""
    B, Fn, n, H, W = masks.shape
    if engine.precise():
        # reference-precision mode: the bit-exact fp32 scatter (reference layout), then PixelUnshuffle(8) as plumbing
        feat13, mask = ops.traj_scatter(info.view(B * Fn, n, 12), masks.view(B * Fn, n, H, W))
        feat = unshuffle8_to_cl(feat13.view(B * Fn, 13, 1, H, W)).view(B * Fn, H // 8, W // 8, 13 * 64)
    else:
        feat, mask = ops.traj_scatter_unshuffle(info.view(B * Fn, n, 12), masks.view(B * Fn, n, H, W))
    if null_clips:
        fv = feat.view(B, Fn, *feat.shape[1:])
        for i in null_clips:
            fv[i].zero_()
    outs = omcm.encode_cl(feat, mask)
    return [CL(o.view(B, Fn, *o.shape[1:])) for o in outs]


def _scatter_inputs(info, masks, null_clips=None):
    B, Fn, n, H, W = masks.shape
    feat, mask = ops.traj_scatter_unshuffle(info.view(B * Fn, n, 12), masks.view(B * Fn, n, H, W))
    if null_clips:
        fv = feat.view(B, Fn, *feat.shape[1:])
        for i in null_clips:
            fv[i].zero_()
    return feat, mask


def traj_features_from_circles(info, circles, omcm, H, W):
    """Device-side form of the `use_sphere_mask` preprocessing (fmc/data/dataset.py:5350-5403) + get_traj_features_v2:
    info [B, F, n, 12] and circles [B, F, n, 3] = (cx, cy, r) of every object's minimum enclosing circle (device fp32;
    r <= 0 = absent) -> 4 CL features.  The Gaussian masks are generated inside the scatter kernel: no [n, H, W] mask
    tensors on the host, no H2D copy of them, no mask reads."""
    B, Fn, n, _ = circles.shape
    if engine.precise():
        masks = ops.sphere_masks(circles.reshape(B * Fn, n, 3), H, W)
        return traj_features_cl(info, masks.view(B, Fn, n, H, W), omcm)
    feat, mask = ops.traj_scatter_circles_unshuffle(info.reshape(B * Fn, n, 12), circles.reshape(B * Fn, n, 3), H, W)
    outs = omcm.encode_cl(feat, mask)
    return [CL(o.view(B, Fn, *o.shape[1:])) for o in outs]


def get_traj_features_v2(obj_info_list_list, obj_mask_list_list, omcm, cfg_random_null_om, cfg_random_null_om_ratio,
                         is_cm_condition_null_list, local_rank, dtype):
    """Reference signature; returns 4 tensors [b, C_l, f, h_l, w_l] (fp32, reference layout)."""
    device = torch.device("cuda", local_rank) if isinstance(local_rank, int) else torch.device(local_rank)
    info, masks = pack_objects(obj_info_list_list, obj_mask_list_list, device)
    null = []
    if cfg_random_null_om:
        null = [i for i in range(info.shape[0]) if not (random.random() > cfg_random_null_om_ratio)]
    return traj_features_packed(info, masks, omcm, null)


def traj_features_packed(info, masks, omcm, null_clips=None):
    """get_traj_features_v2 on objects already packed by `pack_objects` (dense device tensors): the part of the call that
    is device work only, so a captured training step (train.GraphedStep) can contain it while the packing stays with the
    data loading."""
    from .. import train_engine
    null = list(null_clips or [])
    if train_engine.wants_training(omcm):
        # OMC training (train_cam_obj_ctrl.py:843): the ObjectEncoder runs on the tape; the returned features carry its graph
        if engine.precise():
            raise NotImplementedError("training runs in the bf16 mode")
        feat, mask = _scatter_inputs(info, masks, null)
        return train_engine.adapter_train_forward(omcm, feat, mask, info.shape[0], info.shape[1])
    return [f.to_reference() for f in traj_features_cl(info, masks, omcm, null)]
