"""Mirror of fmc/models/unet_cam_obj.py: the CMC + OMC U-Net = UNet3DConditionModelPoseCond + `traj_features`
(passed to the cross-attention down blocks inside cross_attention_kwargs, unet_cam_obj.py:1215-1234)."""
from .unet import UNet3DConditionModel, UNet3DConditionModelPoseCond, _reject_unsupported  # noqa: F401


class UNet3DConditionModelCamObjCond(UNet3DConditionModelPoseCond):
    _accepts_traj_features = True

    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, attention_mask=None,
                cross_attention_kwargs=None, pose_embedding_features=None, traj_features=None, return_dict=True,
                down_block_additional_residuals=None, mid_block_additional_residual=None, motion_module_alphas=1.0,
                debug=False):
        """Parameter order and defaults of the reference's forward (unet_cam_obj.py:1107-1122)."""
        _reject_unsupported(down_block_additional_residuals, mid_block_additional_residual, motion_module_alphas, debug)
        return self._forward_impl(sample, timestep, encoder_hidden_states, class_labels, attention_mask,
                                  cross_attention_kwargs, pose_embedding_features, traj_features, return_dict)
