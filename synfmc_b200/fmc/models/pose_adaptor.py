"""Mirror of fmc/models/pose_adaptor.py: CameraEncoder (`CameraPoseEncoder`, :159-240) and the train-time wrapper
`PoseAdaptor` (:56-72).  PixelUnshuffle(8) -> conv3x3 -> 4 levels x 2 x [ResnetBlock -> TemporalTransformerBlock].
1x1 convs run as tcgen05 GEMMs, the temporal blocks on the same kernels as the U-Net motion modules; with cameras
given as (K, c2w) the Pluecker rays are generated directly in the unshuffled channels-last layout
(`encode_cameras`), skipping the CPU ray build + H2D copy + permute + unshuffle of train_cam_ctrl.py:77-90,584."""
import torch
from torch import nn

from ... import engine, ops
from ...engine import CL
from .motion_module import TemporalTransformerBlock


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        assert dims == 2
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        if use_conv:
            self.op = nn.Conv2d(self.channels, self.out_channels, 3, stride=2, padding=padding)
        else:
            assert self.channels == self.out_channels
            self.op = nn.AvgPool2d(kernel_size=2, stride=2)


class ResnetBlock(nn.Module):
    skep_in_is_out = False  # pose_adaptor.py:115 feeds skep with in_c; fmc/adapter.py:78 with out_c

    def __init__(self, in_c, out_c, down, ksize=3, sk=False, use_conv=True):
        super().__init__()
        in_c, out_c = int(in_c), int(out_c)
        ps = ksize // 2
        self.in_conv = nn.Conv2d(in_c, out_c, ksize, 1, ps) if (in_c != out_c or not sk) else None
        self.block1 = nn.Conv2d(out_c, out_c, 3, 1, 1)
        self.act = nn.ReLU()
        self.block2 = nn.Conv2d(out_c, out_c, ksize, 1, ps)
        self.skep = nn.Conv2d(out_c if self.skep_in_is_out else in_c, out_c, ksize, 1, ps) if not sk else None
        self.down = down
        if self.down:
            self.down_opt = Downsample(in_c, use_conv=use_conv)
        self._plan = None

    def plan(self, device):
        if self._plan is None or self._plan["device"] != engine.plan_key(device):
            self._plan = {
                "device": engine.plan_key(device),
                "in_conv": engine.ConvPlan(self.in_conv, device) if self.in_conv is not None else None,
                "block1": engine.ConvPlan(self.block1, device),
                "block2": engine.ConvPlan(self.block2, device),
                "skep": engine.ConvPlan(self.skep, device) if self.skep is not None else None,
                "down": engine.ConvPlan(self.down_opt.op, device) if (self.down and self.down_opt.use_conv) else None,
            }
        return self._plan

    def run(self, x):
        """x [N, h, w, Cin] bf16 channels-last -> [N, h', w', Cout]."""
        p = self.plan(x.device)
        if self.down:
            x = p["down"](x) if p["down"] is not None else ops.avgpool2(x)
        if p["in_conv"] is not None:
            x = p["in_conv"](x)
        h = p["block1"](x, relu=True)
        N, hh, ww, C = x.shape
        skip = p["skep"](x) if p["skep"] is not None else x
        b2 = p["block2"]
        if b2.linear is not None:
            out = b2.linear(h.view(-1, C), residual=skip.reshape(-1, C))
        else:
            out = ops.add(b2(h).view(-1, C), skip.reshape(-1, C))
        return out.view(N, hh, ww, C)


def unshuffle8_to_cl(x):
    """[b, c, f, H, W] -> PixelUnshuffle(8) -> channels-last bf16 [b, f, H/8, W/8, c*64] (data-layout plumbing for the
    drop-in entry point; the fused kernels fmc_plucker_unshuffle_bf16 / fmc_traj_scatter_unshuffle_bf16 avoid it)."""
    b, c, f, H, W = x.shape
    y = torch.nn.functional.pixel_unshuffle(x.permute(0, 2, 1, 3, 4).reshape(b * f, c, H, W), 8)
    return y.permute(0, 2, 3, 1).reshape(b, f, H // 8, W // 8, c * 64).to(engine.act_dtype()).contiguous()


def cl_to_frames_nchw(x):
    """CL [b, f, h, w, C] -> reference layout [(b f), C, h, w] fp32."""
    b, f, h, w, C = x.dims
    return ops.from_channels_last(x.t.view(b * f, 1, h, w, C)).view(b * f, C, h, w)


class CameraPoseEncoder(nn.Module):
    def __init__(self, downscale_factor, channels=(320, 640, 1280, 1280), nums_rb=3, cin=64, ksize=3, sk=False,
                 use_conv=True, compression_factor=1, temporal_attention_nhead=8,
                 attention_block_types=("Temporal_Self",), temporal_position_encoding=False,
                 temporal_position_encoding_max_len=16, rescale_output_factor=1.0):
        super().__init__()
        assert downscale_factor == 8, "the fused ray kernel and the reference configs use PixelUnshuffle(8)"
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
        self._plan = None

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def invalidate_plans(self):
        engine.invalidate_plans(self)

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self.invalidate_plans()
        return out

    def encode_cl(self, x_cl):
        """x_cl: unshuffled rays [b, f, H/8, W/8, 384] bf16 -> list of 4 CL features [b, f, h_l, w_l, C_l]."""
        b, f, h, w, cin = x_cl.shape
        engine.refresh_plans(self)
        if self._plan is None or self._plan.key != engine.plan_key(x_cl.device):
            self._plan = engine.ConvPlan(self.encoder_conv_in, x_cl.device)
        x = self._plan(x_cl.view(b * f, h, w, cin))
        features = []
        for res_block, attention_block in zip(self.encoder_down_conv_blocks, self.encoder_down_attention_blocks):
            for res_layer, attention_layer in zip(res_block, attention_block):
                x = res_layer.run(x)
                n, hh, ww, C = x.shape
                rows = attention_layer.run(x.view(-1, C), b, f, hh * ww)
                x = rows.view(n, hh, ww, C)
            features.append(CL(x.view(b, f, *x.shape[1:])))
        return features

    def encode_cameras(self, K, c2w, H, W):
        """K [b, f, 4] = (fx, fy, cx, cy), c2w [b, f, 3, 4] (device tensors) -> 4 CL features, rays built on the GPU."""
        b, f = K.shape[:2]
        rays = ops.plucker_unshuffle(K.reshape(b * f, 4), c2w.reshape(b * f, 3, 4), H, W)
        if engine.precise():  # the fused ray kernel writes bf16; the reference-precision mode takes the fp32 rays
            rays6 = ops.plucker(K.reshape(b * f, 4), c2w.reshape(b * f, 3, 4), H, W)  # [bf, H, W, 6]
            return self.encode_cl(unshuffle8_to_cl(rays6.view(b, f, H, W, 6).permute(0, 4, 1, 2, 3)))
        return self.encode_cl(rays.view(b, f, H // 8, W // 8, 384))

    def forward(self, x):
        """Reference signature (:224-240): x [b, 6, f, H, W] -> 4 tensors [(b f), C_l, h_l, w_l] (fp32)."""
        engine.require_no_grad(self, x)
        ops.require_cuda(x)
        return [cl_to_frames_nchw(f) for f in self.encode_cl(unshuffle8_to_cl(x.float()))]


class PoseAdaptor(nn.Module):
    def __init__(self, unet, pose_encoder):
        super().__init__()
        self.unet = unet
        self.pose_encoder = pose_encoder

    def forward(self, noisy_latents, timesteps, encoder_hidden_states, pose_embedding):
        assert pose_embedding.ndim == 5
        from ... import train_engine
        if train_engine.wants_training(self, noisy_latents, pose_embedding):
            # training step (train_cam_ctrl.py:586-648): forward on the tape, backward kernels behind loss.backward()
            return train_engine.pose_adaptor_train_forward(self.unet, self.pose_encoder, noisy_latents, timesteps,
                                                           encoder_hidden_states, pose_embedding)
        feats = self.pose_encoder.encode_cl(unshuffle8_to_cl(pose_embedding.float()))
        return self.unet(noisy_latents, timesteps, encoder_hidden_states, pose_embedding_features=feats).sample
