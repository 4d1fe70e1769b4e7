"""Mirror of fmc/models/pose_obj_adaptor.py:7-23."""
from torch import nn

from ... import engine
from .pose_adaptor import unshuffle8_to_cl


class CamObjPoseAdaptor(nn.Module):
    def __init__(self, unet, pose_encoder):
        super().__init__()
        self.unet = unet
        self.pose_encoder = pose_encoder

    def forward(self, noisy_latents, timesteps, encoder_hidden_states, pose_embedding, traj_features):
        assert pose_embedding.ndim == 5
        engine.require_no_grad(self, noisy_latents, encoder_hidden_states, pose_embedding, *(traj_features or ()))
        feats = self.pose_encoder.encode_cl(unshuffle8_to_cl(pose_embedding.float()))
        return self.unet(noisy_latents, timesteps, encoder_hidden_states, pose_embedding_features=feats,
                         traj_features=traj_features).sample
