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
        from ... import train_engine
        if train_engine.wants_training(self, noisy_latents, pose_embedding, traj_features):
            # training step (train_cam_obj_ctrl.py:843-866): the object features may carry the ObjectEncoder's graph
            return train_engine.pose_adaptor_train_forward(self.unet, self.pose_encoder, noisy_latents, timesteps,
                                                           encoder_hidden_states, pose_embedding, traj_features)
        feats = self.pose_encoder.encode_cl(unshuffle8_to_cl(pose_embedding.float()))
        return self.unet(noisy_latents, timesteps, encoder_hidden_states, pose_embedding_features=feats,
                         traj_features=traj_features).sample
