"""Mirror of fmc/modified_modules.py:52-185: the down-block forwards the OMC trainer binds over the stock ones
(train_cam_obj_ctrl.py:317-329).  After the block's last motion module the ObjectEncoder feature
`traj_features[self.traj_fea_idx]` is added and replaces the last skip tensor, before the downsampler (:115-117).
The add runs in fmc_add_bf16 on channels-last activations."""
from ..engine import CL


def Adapted_CrossAttnDownBlock3D_forward(self, hidden_states, temb=None, encoder_hidden_states=None,
                                         attention_mask=None, motion_module_alpha=1.0, cross_attention_kwargs=None,
                                         motion_cross_attention_kwargs=None):
    cross_attention_kwargs = dict(cross_attention_kwargs or {})
    traj_features = cross_attention_kwargs.pop("traj_features", None)
    hidden_states, output_states = self.run_layers(hidden_states, temb, encoder_hidden_states, cross_attention_kwargs,
                                                   motion_cross_attention_kwargs)
    if traj_features is not None:
        hidden_states = hidden_states + CL.from_reference(traj_features[self.traj_fea_idx])
        output_states = output_states[:-1] + (hidden_states,)
    return self.run_downsample(hidden_states, output_states)


def Adapted_DownBlock3D_forward(self, hidden_states, temb=None, encoder_hidden_states=None, motion_module_alpha=1.0,
                                motion_cross_attention_kwargs=None, **kwargs):
    # 'traj_features' never arrives as a direct keyword (the UNet nests it in cross_attention_kwargs), so ObjectEncoder
    # feature 3 is unused -- reference behaviour, unet_cam_obj.py:1227-1234 vs modified_modules.py:131
    traj_features = kwargs.pop("traj_features", None)
    hidden_states, output_states = self.run_layers(hidden_states, temb, encoder_hidden_states,
                                                   motion_cross_attention_kwargs)
    if traj_features is not None:
        hidden_states = hidden_states + CL.from_reference(traj_features[self.traj_fea_idx])
        output_states = output_states[:-1] + (hidden_states,)
    return self.run_downsample(hidden_states, output_states)


def bind_omcm_forwards(unet):
    """What train_cam_obj_ctrl.py:317-329 does: rebind `.forward`, number the blocks in named_modules() order."""
    idx = 0
    for _name, module in unet.down_blocks.named_modules():
        cls = module.__class__.__name__
        if cls == "CrossAttnDownBlock3D":
            module.forward = Adapted_CrossAttnDownBlock3D_forward.__get__(module, module.__class__)
        elif cls == "DownBlock3D":
            module.forward = Adapted_DownBlock3D_forward.__get__(module, module.__class__)
        else:
            continue
        module.traj_fea_idx = idx
        idx += 1
