"""Camera / object pose bookkeeping the trainers and the dataset use from `fmc.data.utils` (train_cam_ctrl.py:38,
train_cam_obj_ctrl.py:38,57; fmc/data/dataset.py).  Host-side numpy on 16 small matrices per clip -- it feeds the
Plucker embedding (a9) and the ObjectEncoder pose features (a10), it is not on the device path.

Conventions of the reference (fmc/data/utils.py:148-200): a pose is a 3x4 `[R | T]` block (rows of a 4x4 matrix);
"relative to X" means R_rel = R^T R_X and T_rel = R^T (T_X - T) / scale_T."""
import numpy as np
import torch


def _rt34(m):
    m = m.numpy() if isinstance(m, torch.Tensor) else np.asarray(m)
    return np.array(m[:3], copy=True)


def create_relative_matrix_of_cam_list(cam_info, scale_T=1):
    """Poses of a clip relative to its first frame -> [n, 12] tensor (row-major 3x4), row 0 is exactly `eye(3, 4)`
    (fmc/data/utils.py:148-163)."""
    poses = [_rt34(rt) for rt in cam_info]
    r0, t0 = poses[0][:, :3].copy(), poses[0][:, 3].copy()
    rows = []
    for rt in poses:
        r_t = rt[:, :3].T
        rel = np.empty_like(rt)
        rel[:, :3] = r_t @ r0
        rel[:, 3] = (-(r_t @ rt[:, 3]) + r_t @ t0) / scale_T
        rows.append(torch.from_numpy(rel))
    rows[0] = torch.eye(3, 4, dtype=rows[0].dtype)
    return torch.stack([r.reshape(-1) for r in rows])


def create_absolute_matrix_from_ref_cam_list(first_cam_info, ref_cam_info_np, scale_T=1):
    """Inverse of the above for a 16-frame clip: absolute 3x4 poses from the first frame's 4x4 pose and the 16 relative
    3x4 poses (translations rescaled by scale_T first): `first @ inv([rel; 0 0 0 1])`, frame 0 = the first pose itself
    (fmc/data/utils.py:167-183; the frame count is asserted there too)."""
    assert len(ref_cam_info_np) == 16
    first = np.asarray(first_cam_info)
    out = [np.array(first[:3], copy=True)]
    bottom = np.array([[0.0, 0.0, 0.0, 1.0]])
    for rel in ref_cam_info_np[1:]:
        rel = np.array(rel, copy=True)
        rel[:, 3] = rel[:, 3] * scale_T
        out.append((first @ np.linalg.inv(np.concatenate([rel, bottom], axis=0)))[:3])
    return out


def create_relative_matrix_of_two_torch_matrix(RT1, RT2, scale_T=1):
    """Object poses RT2 [n, >=3, 4] relative to the camera pose RT1 -> numpy [n, 12] (fmc/data/utils.py:185-200), the
    `obj_info` rows of get_traj_features_v2.

    Reference quirk kept on purpose: its translation term is `np.dot(R^T, T)[..., 0, 0]` on STACKED arrays, which
    numpy evaluates as an outer product over the two batch axes and then slices at batch index 0 -- so every object's
    translation is computed from the translation of OBJECT 0: T_rel[i] = R_i^T (T1 - T_obj0) / scale_T.  With one
    object per clip (config 2) this is the intended formula."""
    cam = _rt34(RT1)
    obj = RT2.numpy() if isinstance(RT2, torch.Tensor) else np.asarray(RT2)
    obj = np.array(obj[:, :3], copy=True)
    r_t = obj[:, :, :3].transpose(0, 2, 1)
    t_obj0 = obj[0, :, 3].copy()
    out = np.empty_like(obj)
    out[:, :, :3] = r_t @ cam[:, :3]
    out[:, :, 3] = (-(r_t @ t_obj0) + r_t @ cam[:, 3]) / scale_T
    return out.reshape(out.shape[0], -1)


# --------------------------------------------------------------------------------------------------------------
# Device forms: the same three functions on poses that already live on the GPU, batched over clips /
# frames, fp64 like the numpy originals (csrc/pose.cu).  Results stay on the device and feed `ops.plucker*` /
# `fmc.util.pack_objects` without a round trip through numpy.
# --------------------------------------------------------------------------------------------------------------
def _poses_f64(x, what):
    from ... import ops
    ops.require_cuda(x)
    if x.dtype != torch.float64 or x.shape[-1] != 4 or x.shape[-2] not in (3, 4):
        raise ValueError(f"{what}: float64 poses [..., 3 or 4, 4] expected, got {tuple(x.shape)} {x.dtype}")
    return x.contiguous()


def relative_poses_to_first_frame(cams, scale_T=1.0):
    """create_relative_matrix_of_cam_list for a batch: cams [clips, frames, 3|4, 4] float64 (device) -> [clips, frames, 12]."""
    from ... import _cabi, ops
    cams = _poses_f64(cams, "relative_poses_to_first_frame")
    clips, frames, rows, _ = cams.shape
    out = torch.empty((clips, frames, 12), device=cams.device, dtype=torch.float64)
    ops._check_cuda(cams)
    _cabi.call("fmc_pose_relative_to_first_f64", cams.data_ptr(), rows * 4, out.data_ptr(), clips, frames, float(scale_T),
               ops._stream())
    return out


def absolute_poses_from_relative(first, rel, scale_T=1.0):
    """create_absolute_matrix_from_ref_cam_list for a batch: first [clips, 4, 4], rel [clips, frames, 12] (or [..., 3, 4])
    float64 (device) -> [clips, frames, 3, 4]."""
    from ... import _cabi, ops
    ops.require_cuda(first)
    if first.dtype != torch.float64 or tuple(first.shape[-2:]) != (4, 4):
        raise ValueError("absolute_poses_from_relative: first must be float64 [clips, 4, 4]")
    first = first.contiguous()
    clips = first.shape[0]
    rel = rel.reshape(clips, -1, 12).contiguous()
    if rel.dtype != torch.float64:
        raise ValueError("absolute_poses_from_relative: rel must be float64")
    frames = rel.shape[1]
    out = torch.empty((clips, frames, 3, 4), device=first.device, dtype=torch.float64)
    ops._check_cuda(first, rel)
    _cabi.call("fmc_pose_absolute_from_relative_f64", first.data_ptr(), rel.data_ptr(), out.data_ptr(), clips, frames,
               float(scale_T), ops._stream())
    return out


def relative_object_poses(cams, objs, scale_T=1.0):
    """create_relative_matrix_of_two_torch_matrix for a batch of (camera pose, n object poses) sets -- typically one set per
    frame: cams [sets, 3|4, 4], objs [sets, n, 3|4, 4] float64 (device) -> [sets, n, 12], the `obj_info` rows of
    get_traj_features_v2 (the reference's object-0 translation quirk included, see the host function above)."""
    from ... import _cabi, ops
    cams = _poses_f64(cams, "relative_object_poses")
    objs = _poses_f64(objs, "relative_object_poses")
    sets, n = objs.shape[0], objs.shape[1]
    assert cams.shape[0] == sets
    out = torch.empty((sets, n, 12), device=cams.device, dtype=torch.float64)
    ops._check_cuda(cams, objs)
    _cabi.call("fmc_pose_objects_relative_f64", cams.data_ptr(), cams.shape[-2] * 4, objs.data_ptr(), objs.shape[-2] * 4,
               out.data_ptr(), sets, n, float(scale_T), ops._stream())
    return out
