"""Mirror of fmc/models/unet_cam_obj.py: the CMC + OMC U-Net = UNet3DConditionModelPoseCond + `traj_features`
(passed to the cross-attention down blocks inside cross_attention_kwargs, unet_cam_obj.py:1215-1234)."""
from .unet import UNet3DConditionModel, UNet3DConditionModelPoseCond  # noqa: F401


class UNet3DConditionModelCamObjCond(UNet3DConditionModelPoseCond):
    _accepts_traj_features = True
