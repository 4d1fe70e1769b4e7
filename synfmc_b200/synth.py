"""Synthetic weights and inputs of the shapes BASELINE.json names (no checkpoints or datasets exist offline).

`synth_init_` re-draws every parameter of a module from a generator seeded by the PARAMETER NAME, so two
implementations with the same state-dict keys (this package's mirror of `fmc.models` and the test oracle) get
identical weights without sharing code.  The reference's zero-inits (qkv_merge attention_processor.py:191-192,
LoRA `up`, Adapter zero-convs adapter.py:129-146, proj_out when zero_initialize) are overridden with non-zero
draws, otherwise the CameraAdapter / Domain-LoRA / ObjectEncoder paths would be vacuous.
"""
import math
import zlib

import numpy as np
import torch


def _gen(name, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


@torch.no_grad()
def synth_init_(module, seed=0, bf16_exact=True):
    """In-place deterministic init.  bf16_exact rounds every value to a bf16-representable fp32 so the fp32 oracle and
    the bf16 kernels consume bit-identical weights."""
    for name, p in module.named_parameters():
        g = _gen(name, seed)
        if p.ndim >= 2:
            fan_in = int(np.prod(p.shape[1:]))
            std = 1.0 / math.sqrt(fan_in)
            if "lora" in name and ".up." in name:
                std *= 0.5
            if "merge" in name or "zero_conv" in name:
                std *= 0.5
            v = torch.randn(p.shape, generator=g) * std
        elif name.endswith("weight"):  # 1-D weight = norm scale
            v = 1.0 + 0.02 * torch.randn(p.shape, generator=g)
        else:
            v = 0.02 * torch.randn(p.shape, generator=g)
        if bf16_exact:
            v = v.to(torch.bfloat16).to(torch.float32)
        p.copy_(v.to(p.dtype))
    return module


def round_bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def synth_camera(b, f, H, W, seed=0):
    """Intrinsics [b,f,4] = (fx,fy,cx,cy) and relative c2w [b,f,3,4]: smooth random walk, frame 0 = identity
    (fmc/data/utils.py:161), rotation <= ~30 deg total, |translation| <= ~1 after the /1200 rescale."""
    g = _gen("camera", seed)
    K = torch.tensor([0.9 * W, 0.9 * W, W / 2.0, H / 2.0]).repeat(b, f, 1)
    c2w = torch.zeros(b, f, 3, 4)
    for bi in range(b):
        axis = torch.randn(3, generator=g)
        axis = axis / axis.norm()
        total_angle = math.radians(30.0) * float(torch.rand(1, generator=g))
        direction = torch.randn(3, generator=g)
        direction = direction / direction.norm() * float(torch.rand(1, generator=g))
        for fi in range(f):
            a = total_angle * fi / max(f - 1, 1)
            Kx = torch.tensor([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
            R = torch.eye(3) + math.sin(a) * Kx + (1 - math.cos(a)) * (Kx @ Kx)
            c2w[bi, fi, :, :3] = R
            c2w[bi, fi, :, 3] = direction * fi / max(f - 1, 1)
    return K, c2w


def synth_objects(b, f, H, W, n_obj, seed=0, gaussian=True):
    """obj_info: list[b] of list[f] of float64 numpy [n_obj, 12] (relative 3x4 RT, translation/1000);
    masks: list[b] of list[f] of float64 tensors [n_obj, 1, H, W] -- Gaussian discs exp(-0.5 (r/sigma)^2), sigma = R/2,
    clipped to the disc and normalised to max 1 (fmc/data/dataset.py:5350-5403), or {0,1} ellipses.  Centres follow a
    random walk and are drawn close together so masks overlap (order-dependent scatter)."""
    g = _gen("objects", seed)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    infos, masks = [], []
    for bi in range(b):
        centre0 = torch.tensor([H / 2.0, W / 2.0], dtype=torch.float64)
        radius = 20 + 60 * torch.rand(n_obj, generator=g, dtype=torch.float64)
        radius = torch.minimum(radius, torch.tensor(min(H, W) / 3.0, dtype=torch.float64))
        offs = (torch.rand(n_obj, 2, generator=g, dtype=torch.float64) - 0.5) * radius.mean() * 1.5
        vel = (torch.rand(n_obj, 2, generator=g, dtype=torch.float64) - 0.5) * 6.0
        rt0 = torch.randn(n_obj, 3, 4, generator=g, dtype=torch.float64)
        info_f, mask_f = [], []
        for fi in range(f):
            m = torch.zeros(n_obj, 1, H, W, dtype=torch.float64)
            info = np.zeros((n_obj, 12), dtype=np.float64)
            for o in range(n_obj):
                cy, cx = (centre0 + offs[o] + vel[o] * fi).tolist()
                r2 = (yy - cy) ** 2 + (xx - cx) ** 2
                R = float(radius[o])
                inside = r2 <= R * R
                if gaussian:
                    val = torch.exp(-0.5 * r2 / (R / 2.0) ** 2) * inside
                    val = val / val.max().clamp_min(1e-12)
                else:
                    val = inside.to(torch.float64)
                m[o, 0] = val
                rt = rt0[o].clone()
                rt[:, 3] = rt[:, 3] * (1.0 + 0.05 * fi) / 10.0
                rt[:, :3] = rt[:, :3] / rt[:, :3].norm(dim=1, keepdim=True)
                info[o] = rt.reshape(-1).numpy()
            info_f.append(info)
            mask_f.append(m)
        infos.append(info_f)
        masks.append(mask_f)
    return infos, masks


def synth_circles(b, f, H, W, n_obj, seed=0):
    """Object poses + minimum-enclosing circles (what the dataset's `use_sphere_mask` branch derives from the segmentation
    masks, fmc/data/dataset.py:5359): info [b, f, n_obj, 12] fp32 and circles [b, f, n_obj, 3] = (cx, cy, r) fp32 with
    sub-pixel centres on a random walk, radii 20 - 80 px, overlapping objects, and one absent object (r = 0) per clip
    when n_obj > 1."""
    g = _gen("circles", seed)
    info = torch.randn(b, f, n_obj, 12, generator=g)
    info[..., 3::4] *= 0.1
    circles = torch.zeros(b, f, n_obj, 3)
    for bi in range(b):
        radius = 20 + 60 * torch.rand(n_obj, generator=g)
        radius = torch.minimum(radius, torch.tensor(min(H, W) / 3.0))
        c0 = torch.tensor([W / 2.0, H / 2.0]) + (torch.rand(n_obj, 2, generator=g) - 0.5) * float(radius.mean()) * 1.5
        vel = (torch.rand(n_obj, 2, generator=g) - 0.5) * 6.0
        for fi in range(f):
            circles[bi, fi, :, :2] = c0 + vel * fi
            circles[bi, fi, :, 2] = radius * (1.0 + 0.01 * fi)
        if n_obj > 1:
            circles[bi, f // 2, n_obj - 1, 2] = 0.0  # the object leaves the frame: empty segmentation mask
    return info, circles


def synth_step_inputs(b, f, h, w, cfg=True, seed=0, text_len=77, text_dim=768):
    """Latents ~ N(0,1) [b,4,f,h,w]; text embeddings ~ 0.5 N(0,1) (CLIP-like scale), uncond = a different draw."""
    g = _gen("step_inputs", seed)
    latents = torch.randn(b, 4, f, h, w, generator=g)
    text = 0.5 * torch.randn(b, text_len, text_dim, generator=g)
    if cfg:
        uncond = 0.5 * torch.randn(b, text_len, text_dim, generator=g)
        text = torch.cat([uncond, text], dim=0)
    return latents, round_bf16(text)
