"""Mirror of the one hot-path function of fmc/data/dataset.py: `ray_condition` (:930-972), the Pluecker-ray embedding,
computed on the GPU by fmc_plucker_f32 instead of on the CPU (train_cam_ctrl.py:87 passes device='cpu').

Everything else the trainers import from this module (`UnrealTrajVideoDataset`, `UnrealTrajLoraDataset`,
train_cam_ctrl.py:37) is CPU data loading outside the hot path: those names resolve lazily to the reference's own
module when a checkout is known (synfmc_b200/dropin.py), and raise an explanatory AttributeError otherwise."""
import torch

from ... import dropin, ops


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return dropin.reference_attr("data.dataset", name)


def ray_condition(K, c2w, H, W, device, flip_flag=None):
    """K [B, V, 4] = (fx, fy, cx, cy); c2w [B, V, 4, 4] -> [B, V, H, W, 6] = (o x d, d), fp32 on `device`.
    (The reference's default-dim torch.cross misbehaves when B == 3 or V == 3; dim=-1 is implemented)"""
    if flip_flag is not None and int(torch.as_tensor(flip_flag).sum().item()) > 0:
        raise NotImplementedError("horizontal flip is never enabled by the trainers (flip_flag = zeros)")
    dev = torch.device(device)
    # `device` says where the RESULT lives.  The trainers ask for 'cpu' (train_cam_ctrl.py:87) and move the embedding
    # to the GPU afterwards: the rays are still built by the kernel on the current CUDA device and copied back -- there
    # is no CPU implementation behind this function.
    if dev.type == "cuda":
        compute = dev
    elif torch.cuda.is_available():
        compute = torch.device("cuda", torch.cuda.current_device())
    else:
        raise RuntimeError("synfmc_b200.fmc.data.dataset.ray_condition needs a CUDA device (no CPU fallback)")
    B, V = K.shape[:2]
    Kd = K.to(compute, torch.float32).reshape(B * V, 4)
    M = c2w.to(compute, torch.float32)[..., :3, :].reshape(B * V, 3, 4)
    rays = ops.plucker(Kd, M, H, W).view(B, V, H, W, 6)
    return rays if dev == compute else rays.to(dev)
