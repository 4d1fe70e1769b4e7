"""ORACLE (test infrastructure, CPU fp32; parity unpinned -- see oracle/diffusers_restated.py).

Restates fmc/models/pose_adaptor.py (CameraEncoder) and fmc/models/pose_obj_adaptor.py:
  PoseAdaptor :56-72, Downsample :75-99, ResnetBlock :102-135, CameraPoseEncoder :159-240,
  CamObjPoseAdaptor pose_obj_adaptor.py:7-23.
"""
import torch
from torch import nn

from .motion_module import TemporalTransformerBlock


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        assert dims == 2
        self.channels = channels
        self.out_channels = out_channels or channels
        if use_conv:
            self.op = nn.Conv2d(self.channels, self.out_channels, 3, stride=2, padding=padding)
        else:
            assert self.channels == self.out_channels
            self.op = nn.AvgPool2d(kernel_size=2, stride=2)

    def forward(self, x):
        assert x.shape[1] == self.channels
        return self.op(x)


class ResnetBlock(nn.Module):
    """(down) -> (in_conv if cin != cout or not sk) -> block1 3x3 -> ReLU -> block2 ksize -> + (skep(x) | x)."""

    def __init__(self, in_c, out_c, down, ksize=3, sk=False, use_conv=True):
        super().__init__()
        in_c, out_c = int(in_c), int(out_c)
        ps = ksize // 2
        self.in_conv = nn.Conv2d(in_c, out_c, ksize, 1, ps) if (in_c != out_c or not sk) else None
        self.block1 = nn.Conv2d(out_c, out_c, 3, 1, 1)
        self.act = nn.ReLU()
        self.block2 = nn.Conv2d(out_c, out_c, ksize, 1, ps)
        self.skep = nn.Conv2d(in_c, out_c, ksize, 1, ps) if not sk else None
        self.down = down
        if self.down:
            self.down_opt = Downsample(in_c, use_conv=use_conv)

    def forward(self, x):
        if self.down:
            x = self.down_opt(x)
        if self.in_conv is not None:
            x = self.in_conv(x)
        h = self.block2(self.act(self.block1(x)))
        return h + (self.skep(x) if self.skep is not None else x)


class CameraPoseEncoder(nn.Module):
    def __init__(self, downscale_factor, channels=(320, 640, 1280, 1280), nums_rb=3, cin=64, ksize=3, sk=False,
                 use_conv=True, compression_factor=1, temporal_attention_nhead=8,
                 attention_block_types=("Temporal_Self",), temporal_position_encoding=False,
                 temporal_position_encoding_max_len=16, rescale_output_factor=1.0):
        super().__init__()
        self.unshuffle = nn.PixelUnshuffle(downscale_factor)
        self.channels = list(channels)
        self.nums_rb = nums_rb
        self.encoder_down_conv_blocks = nn.ModuleList()
        self.encoder_down_attention_blocks = nn.ModuleList()
        for i in range(len(channels)):
            convs, attns = nn.ModuleList(), nn.ModuleList()
            for j in range(nums_rb):
                mid = int(channels[i] / compression_factor)
                if j == 0 and i != 0:
                    in_dim, out_dim, down = channels[i - 1], mid, True
                elif j == 0:
                    in_dim, out_dim, down = channels[0], mid, False
                elif j == nums_rb - 1:
                    in_dim, out_dim, down = mid, channels[i], False
                else:
                    in_dim, out_dim, down = mid, mid, False
                convs.append(ResnetBlock(in_dim, out_dim, down=down, ksize=ksize, sk=sk, use_conv=use_conv))
                attns.append(TemporalTransformerBlock(
                    dim=out_dim, num_attention_heads=temporal_attention_nhead,
                    attention_head_dim=int(out_dim / temporal_attention_nhead),
                    attention_block_types=tuple(attention_block_types), dropout=0.0, cross_attention_dim=None,
                    temporal_position_encoding=temporal_position_encoding,
                    temporal_position_encoding_max_len=temporal_position_encoding_max_len,
                    rescale_output_factor=rescale_output_factor))
            self.encoder_down_conv_blocks.append(convs)
            self.encoder_down_attention_blocks.append(attns)
        self.encoder_conv_in = nn.Conv2d(cin, channels[0], 3, 1, 1)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def forward(self, x):
        bs, c, f, H, W = x.shape
        x = self.unshuffle(x.permute(0, 2, 1, 3, 4).reshape(bs * f, c, H, W))
        x = self.encoder_conv_in(x)
        features = []
        for res_block, attention_block in zip(self.encoder_down_conv_blocks, self.encoder_down_attention_blocks):
            for res_layer, attention_layer in zip(res_block, attention_block):
                x = res_layer(x)
                ch, h, w = x.shape[1:]
                t = x.reshape(bs, f, ch, h, w).permute(0, 3, 4, 1, 2).reshape(bs * h * w, f, ch)
                t = attention_layer(t)
                x = t.reshape(bs, h, w, f, ch).permute(0, 3, 4, 1, 2).reshape(bs * f, ch, h, w)
            features.append(x)
        return features


def _to_bcfhw(feats, bs):
    return [x.reshape(bs, x.shape[0] // bs, *x.shape[1:]).permute(0, 2, 1, 3, 4) for x in feats]


class PoseAdaptor(nn.Module):
    def __init__(self, unet, pose_encoder):
        super().__init__()
        self.unet = unet
        self.pose_encoder = pose_encoder

    def forward(self, noisy_latents, timesteps, encoder_hidden_states, pose_embedding):
        assert pose_embedding.ndim == 5
        feats = _to_bcfhw(self.pose_encoder(pose_embedding), pose_embedding.shape[0])
        return self.unet(noisy_latents, timesteps, encoder_hidden_states, pose_embedding_features=feats).sample


class CamObjPoseAdaptor(nn.Module):
    def __init__(self, unet, pose_encoder):
        super().__init__()
        self.unet = unet
        self.pose_encoder = pose_encoder

    def forward(self, noisy_latents, timesteps, encoder_hidden_states, pose_embedding, traj_features):
        assert pose_embedding.ndim == 5
        feats = _to_bcfhw(self.pose_encoder(pose_embedding), pose_embedding.shape[0])
        return self.unet(noisy_latents, timesteps, encoder_hidden_states, pose_embedding_features=feats,
                         traj_features=traj_features).sample
