"""ORACLE (test infrastructure, CPU fp32; parity unpinned -- see oracle/diffusers_restated.py).

Restates fmc/adapter.py: the ObjectEncoder (`Adapter`, :109-192) = T2I-Adapter body + pre/post zero-convs +
per-level mask modulation.  Note the masked tensor is what flows into the next level (:177 rebinds x).
"""
import torch
import torch.nn.functional as F
from torch import nn


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
    """adapter.py:64-98 -- differs from the CameraEncoder one only in skep's input width (out_c, :78)."""

    def __init__(self, in_c, out_c, down, ksize=3, sk=False, use_conv=True):
        super().__init__()
        ps = ksize // 2
        self.in_conv = nn.Conv2d(in_c, out_c, ksize, 1, ps) if (in_c != out_c or not sk) else None
        self.block1 = nn.Conv2d(out_c, out_c, 3, 1, 1)
        self.act = nn.ReLU()
        self.block2 = nn.Conv2d(out_c, out_c, ksize, 1, ps)
        self.skep = nn.Conv2d(out_c, out_c, ksize, 1, ps) if not sk else None
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


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


class Adapter(nn.Module):
    def __init__(self, channels=(320, 640, 1280, 1280), nums_rb=3, cin=64, ksize=3, sk=False, use_conv=True,
                 align_training_size=0, use_pre_zero_conv=False, use_post_zero_conv=False):
        super().__init__()
        assert align_training_size == 0
        self.align_training_size = align_training_size
        self.unshuffle = nn.PixelUnshuffle(8)
        self.channels = list(channels)
        self.nums_rb = nums_rb
        body = []
        for i in range(len(channels)):
            for j in range(nums_rb):
                if i != 0 and j == 0:
                    body.append(ResnetBlock(channels[i - 1], channels[i], down=True, ksize=ksize, sk=sk, use_conv=use_conv))
                else:
                    body.append(ResnetBlock(channels[i], channels[i], down=False, ksize=ksize, sk=sk, use_conv=use_conv))
        self.body = nn.ModuleList(body)
        self.conv_in = nn.Conv2d(cin, channels[0], 3, 1, 1)
        self.zero_conv_in = zero_module(nn.Conv2d(cin, cin, 1)) if use_pre_zero_conv else nn.Identity()
        self.zero_conv_out_list = nn.ModuleList(
            [zero_module(nn.Conv2d(c, c, 1)) if use_post_zero_conv else nn.Identity() for c in channels])

    def forward(self, x, mask_feat):
        x = self.unshuffle(x)
        x = self.conv_in(self.zero_conv_in(x))
        features = []
        for i in range(len(self.channels)):
            for j in range(self.nums_rb):
                x = self.body[i * self.nums_rb + j](x)
            x = self.zero_conv_out_list[i](x)
            if mask_feat is not None:
                # iterated nearest resize: level l samples the previous level's mask (:176)
                mask_feat = F.interpolate(mask_feat, size=x.size()[-2:], mode="nearest")
                x = mask_feat * x
            features.append(x)
        return features
