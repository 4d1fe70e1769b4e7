"""ORACLE (test infrastructure, CPU fp32; parity unpinned -- see oracle/diffusers_restated.py).

Restates fmc/util.py:147-213 `get_traj_features_v2`: 6-D object pose broadcast x (Gaussian) mask scatter that
feeds the ObjectEncoder.  Objects are written in list order, so where masks overlap the LAST object with
mask > 0 wins (:178-182); the info channels end up as info * m * m and the mask channel as m * m (:176,:200).
"""
import random

import numpy as np
import torch


def build_traj_inputs(obj_info_list_list, obj_mask_list_list, device, dtype):
    """Returns (features [(b f), 13, H, W], mask [(b f), 1, H, W]) exactly as handed to the Adapter."""
    assert len(obj_info_list_list) == len(obj_mask_list_list)
    B = len(obj_info_list_list)
    Fn = len(obj_info_list_list[0])
    H, W = obj_mask_list_list[0][0].shape[-2:]
    traj = torch.zeros([B, Fn, H, W, 12], dtype=dtype).to(device)
    maskf = torch.zeros([B, Fn, H, W, 1], dtype=dtype).to(device)
    for b, (infos, masks) in enumerate(zip(obj_info_list_list, obj_mask_list_list)):
        for f, (obj_info, obj_mask) in enumerate(zip(infos, masks)):
            obj_mask = torch.as_tensor(obj_mask).permute(0, 2, 3, 1).to(device=device, dtype=dtype)  # [n, H, W, 1]
            info = torch.from_numpy(np.asarray(obj_info)).unsqueeze(1).unsqueeze(1).expand(-1, H, W, -1)
            info = info.to(device=device, dtype=dtype)
            masked_info = info * obj_mask
            for one_info, one_mask in zip(masked_info, obj_mask):
                sel = (one_mask > 0)[..., 0]
                traj[b][f][sel] = one_info[sel]
                maskf[b][f][sel] = one_mask[sel]
    return traj, maskf


def get_traj_features_v2(obj_info_list_list, obj_mask_list_list, omcm, cfg_random_null_om, cfg_random_null_om_ratio,
                         is_cm_condition_null_list, local_rank, dtype):
    traj, maskf = build_traj_inputs(obj_info_list_list, obj_mask_list_list, local_rank, dtype)
    features = torch.cat([traj, maskf], dim=-1)
    if cfg_random_null_om:
        for i in range(features.shape[0]):
            features[i] = features[i] if (random.random() > cfg_random_null_om_ratio) else torch.zeros_like(features[i])
    b, f, h, w, c = features.shape
    features = features * maskf
    features = features.permute(0, 1, 4, 2, 3).reshape(b * f, c, h, w)
    mask_in = maskf.permute(0, 1, 4, 2, 3).reshape(b * f, 1, h, w)
    outs = omcm(features, mask_in)
    return [o.reshape(b, f, *o.shape[1:]).permute(0, 2, 1, 3, 4) for o in outs]
