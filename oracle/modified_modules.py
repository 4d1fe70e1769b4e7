"""ORACLE (test infrastructure, CPU fp32; parity unpinned -- see oracle/diffusers_restated.py).

Restates fmc/modified_modules.py:52-185: the down-block forwards the OMC trainer binds over the stock ones
(train_cam_obj_ctrl.py:317-329) to inject ObjectEncoder features after the block's last motion module and before
the downsampler; the last skip tensor is replaced by the sum (:115-117).
"""


def Adapted_CrossAttnDownBlock3D_forward(self, hidden_states, temb=None, encoder_hidden_states=None,
                                         attention_mask=None, motion_module_alpha=1.0, cross_attention_kwargs=None,
                                         motion_cross_attention_kwargs=None):
    cross_attention_kwargs = dict(cross_attention_kwargs or {})
    traj_features = cross_attention_kwargs.pop("traj_features", None)
    hidden_states, output_states = self.run_layers(hidden_states, temb, encoder_hidden_states,
                                                   cross_attention_kwargs, motion_cross_attention_kwargs)
    if traj_features is not None:
        hidden_states = hidden_states + traj_features[self.traj_fea_idx]
        output_states = output_states[:-1] + (hidden_states,)
    return self.run_downsample(hidden_states, output_states)


def Adapted_DownBlock3D_forward(self, hidden_states, temb=None, encoder_hidden_states=None, motion_module_alpha=1.0,
                                motion_cross_attention_kwargs=None, **kwargs):
    # 'traj_features' is never a direct keyword here: the UNet nests it inside cross_attention_kwargs
    # (unet_cam_obj.py:1227-1234 vs modified_modules.py:131), so ObjectEncoder feature 3 is dead
    traj_features = kwargs.pop("traj_features", None)
    hidden_states, output_states = self.run_layers(hidden_states, temb, encoder_hidden_states,
                                                   motion_cross_attention_kwargs)
    if traj_features is not None:
        hidden_states = hidden_states + traj_features[self.traj_fea_idx]
        output_states = output_states[:-1] + (hidden_states,)
    return self.run_downsample(hidden_states, output_states)


def bind_omcm_forwards(unet):
    """train_cam_obj_ctrl.py:317-329: rebind `.forward` and number the blocks in named_modules() order."""
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
