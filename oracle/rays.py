"""ORACLE (test infrastructure, CPU fp32; parity unpinned -- see oracle/diffusers_restated.py).

Restates fmc/data/dataset.py:922-972 `ray_condition` and train_cam_ctrl.py:77-90 `to_plucker_embedding`:
the Pluecker-ray embedding (o x d, d) of every pixel centre.  The reference's `torch.cross(o, d)` relies on the
default-dim heuristic and crosses the wrong axis when B == 3 or V == 3; the intended dim=-1 is
restated here.
"""
import torch


def ray_condition(K, c2w, H, W, device, flip_flag=None):
    B, V = K.shape[:2]
    j, i = torch.meshgrid(torch.linspace(0, H - 1, H, device=device, dtype=c2w.dtype),
                          torch.linspace(0, W - 1, W, device=device, dtype=c2w.dtype), indexing="ij")
    i = i.reshape(1, 1, H * W).expand(B, V, H * W) + 0.5
    j = j.reshape(1, 1, H * W).expand(B, V, H * W) + 0.5
    if flip_flag is not None and int(torch.sum(flip_flag).item()) > 0:
        raise NotImplementedError("horizontal flip is never enabled by the trainers (flip_flag = zeros)")
    fx, fy, cx, cy = K.chunk(4, dim=-1)
    zs = torch.ones_like(i)
    xs = (i - cx) / fx * zs
    ys = (j - cy) / fy * zs
    directions = torch.stack((xs, ys, zs), dim=-1).to(c2w)
    directions = directions / directions.norm(dim=-1, keepdim=True)
    rays_d = directions @ c2w[..., :3, :3].transpose(-1, -2)
    rays_o = c2w[..., :3, 3][:, :, None].expand_as(rays_d)
    rays_dxo = torch.cross(rays_o, rays_d, dim=-1)
    return torch.cat([rays_dxo, rays_d], dim=-1).reshape(B, c2w.shape[1], H, W, 6)


def to_plucker_embedding(c2w_rel_poses, intrinsics, sample_size):
    """c2w [B, f, 3, 4], intrinsics [B, f, 4] = (fx, fy, cx, cy) -> [B, f, 6, H, W]."""
    intrinsics = torch.as_tensor(intrinsics)
    c2w = torch.as_tensor(c2w_rel_poses)
    B, n_frame = c2w.shape[:2]
    bottom = torch.tensor([0, 0, 0, 1], dtype=c2w.dtype).view(1, 1, 1, 4).expand(B, n_frame, 1, 4)
    c2w = torch.cat([c2w, bottom], dim=2)
    return ray_condition(intrinsics, c2w, sample_size[0], sample_size[1], device="cpu").permute(0, 1, 4, 2, 3).contiguous()
