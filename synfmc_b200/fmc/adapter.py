"""Mirror of fmc/adapter.py: the ObjectEncoder `Adapter` (:109-192) = T2I-Adapter body + pre/post zero-convs +
per-level mask modulation `x = interpolate(mask, nearest) * x` (:175-177; the masked tensor feeds the next level)."""
import torch
from torch import nn

from .. import engine, ops
from ..engine import CL
from .models.pose_adaptor import ResnetBlock as _PoseResnetBlock
from .models.pose_adaptor import cl_to_frames_nchw, unshuffle8_to_cl


class ResnetBlock(_PoseResnetBlock):
    skep_in_is_out = True  # adapter.py:78


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


class Adapter(nn.Module):
    def __init__(self, channels=(320, 640, 1280, 1280), nums_rb=3, cin=64, ksize=3, sk=False, use_conv=True,
                 align_training_size=0, use_pre_zero_conv=False, use_post_zero_conv=False):
        super().__init__()
        if align_training_size != 0:
            raise NotImplementedError("align_training_size > 0 ends in `assert False` in the reference (adapter.py:182)")
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
        self._plan = None

    def invalidate_plans(self):
        engine.invalidate_plans(self)

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self.invalidate_plans()
        return out

    def plan(self, device):
        engine.refresh_plans(self)
        if self._plan is None or self._plan["device"] != engine.plan_key(device):
            def maybe(m):
                return engine.ConvPlan(m, device) if isinstance(m, nn.Conv2d) else None
            self._plan = {"device": engine.plan_key(device), "conv_in": engine.ConvPlan(self.conv_in, device),
                          "zero_in": maybe(self.zero_conv_in), "zero_out": [maybe(m) for m in self.zero_conv_out_list]}
        return self._plan

    def encode_cl(self, x_cl, mask):
        """x_cl: unshuffled object features [N, H/8, W/8, cin] bf16; mask [N, H, W] fp32 or None.
        Returns 4 channels-last tensors [N, h_l, w_l, C_l]."""
        p = self.plan(x_cl.device)
        x = x_cl
        if p["zero_in"] is not None:
            x = p["zero_in"](x)
        x = p["conv_in"](x)
        features = []
        sizes_h, sizes_w = ([mask.shape[1]], [mask.shape[2]]) if mask is not None else (None, None)
        for i in range(len(self.channels)):
            for j in range(self.nums_rb):
                x = self.body[i * self.nums_rb + j].run(x)
            if p["zero_out"][i] is not None:
                x = p["zero_out"][i](x)
            if mask is not None:
                sizes_h.append(x.shape[1])
                sizes_w.append(x.shape[2])
                ry = engine.nearest_index_chain(sizes_h)[-1].to(x.device)
                rx = engine.nearest_index_chain(sizes_w)[-1].to(x.device)
                x = ops.mask_modulate(x, mask, ry, rx)
            features.append(x)
        return features

    def forward(self, x, mask_feat):
        """Reference signature (:154): x [(b f), 13, H, W], mask_feat [(b f), 1, H, W] -> 4 x [(b f), C_l, h_l, w_l]."""
        engine.require_no_grad(self, x, mask_feat)
        ops.require_cuda(x)
        n, c, H, W = x.shape
        x_cl = unshuffle8_to_cl(x.float().view(n, c, 1, H, W)).view(n, H // 8, W // 8, c * 64)
        mask = mask_feat.float().reshape(n, H, W).contiguous() if mask_feat is not None else None
        feats = self.encode_cl(x_cl, mask)
        return [cl_to_frames_nchw(CL(f.view(n, 1, *f.shape[1:]))) for f in feats]
