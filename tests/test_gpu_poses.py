"""the relative-pose algebra on the device (csrc/pose.cu, fmc.data.utils.relative_* / absolute_*) against the
reference's own outputs (tests/golden/fmc_reference_utils.pt, produced by running fmc/data/utils.py) and against the host
functions on batches of random poses.  fp64 on both sides: 1e-12."""
import os

import numpy as np
import pytest
import torch

from synfmc_b200.fmc.data import utils as du
from tests.golden.make_golden_utils import utils_inputs

pytestmark = pytest.mark.gpu
GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fmc_reference_utils.pt"),
                  weights_only=False)


def _random_poses(n, seed):
    g = torch.Generator().manual_seed(seed)
    q, _ = torch.linalg.qr(torch.randn(n, 3, 3, generator=g, dtype=torch.float64))
    out = torch.zeros(n, 4, 4, dtype=torch.float64)
    out[:, :3, :3] = q
    out[:, :3, 3] = torch.randn(n, 3, generator=g, dtype=torch.float64) * 500
    out[:, 3, 3] = 1.0
    return out


def test_reference_golden_values(cuda_device):
    inp = utils_inputs()
    cams = torch.stack(list(inp["cams"])).double()                      # [16, 4, 4]
    with torch.cuda.device(cuda_device):
        rel = du.relative_poses_to_first_frame(cams[None].to(cuda_device), scale_T=1200)
        assert rel.shape == (1, 16, 12) and rel.dtype == torch.float64
        assert torch.equal(rel[0, 0].cpu(), torch.eye(3, 4, dtype=torch.float64).reshape(-1))
        np.testing.assert_allclose(rel[0].cpu().numpy(), GOLD["relative_cam_list"].numpy(), rtol=1e-12, atol=1e-12)
        absol = du.absolute_poses_from_relative(cams[:1].to(cuda_device), rel, scale_T=1200)
        np.testing.assert_allclose(absol[0].cpu().numpy(), GOLD["absolute_from_ref"], rtol=1e-10, atol=1e-9)
        objs = inp["objs"].double()
        got = du.relative_object_poses(cams[3:4].to(cuda_device), objs[None].to(cuda_device), scale_T=1000)
        np.testing.assert_allclose(got[0].cpu().numpy(), GOLD["relative_two"], rtol=1e-12, atol=1e-12)
        one = du.relative_object_poses(cams[5:6].to(cuda_device), objs[None, :1].to(cuda_device), scale_T=1000)
        np.testing.assert_allclose(one[0].cpu().numpy(), GOLD["relative_two_single"], rtol=1e-12, atol=1e-12)


def test_batches_match_the_host_functions(cuda_device):
    clips, frames, n = 3, 16, 4
    cams = _random_poses(clips * frames, 1).view(clips, frames, 4, 4)
    objs = _random_poses(frames * n, 2).view(frames, n, 4, 4)
    with torch.cuda.device(cuda_device):
        rel = du.relative_poses_to_first_frame(cams.to(cuda_device), scale_T=1200).cpu()
        rel34 = du.relative_poses_to_first_frame(cams[:, :, :3].contiguous().to(cuda_device), scale_T=1200).cpu()
        assert torch.equal(rel, rel34)                                   # 3x4 and 4x4 storage
        absol = du.absolute_poses_from_relative(cams[:, 0].to(cuda_device), rel.to(cuda_device), scale_T=1200).cpu()
        obj_rel = du.relative_object_poses(cams[0].to(cuda_device), objs.to(cuda_device), scale_T=1000).cpu()
    for c in range(clips):
        want = du.create_relative_matrix_of_cam_list(list(cams[c]), scale_T=1200)
        np.testing.assert_allclose(rel[c].numpy(), want.numpy(), rtol=1e-12, atol=1e-12)
        want_abs = du.create_absolute_matrix_from_ref_cam_list(cams[c, 0].numpy(), want.reshape(16, 3, 4).numpy(), scale_T=1200)
        np.testing.assert_allclose(absol[c].numpy(), np.stack(want_abs), rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(absol[c].numpy(), cams[c, :, :3].numpy(), rtol=1e-9, atol=1e-8)   # round trip
    for f in range(frames):
        want = du.create_relative_matrix_of_two_torch_matrix(cams[0, f], objs[f], scale_T=1000)
        np.testing.assert_allclose(obj_rel[f].numpy(), want, rtol=1e-12, atol=1e-12)


def test_rejects_host_and_fp32_poses(cuda_device):
    with pytest.raises(Exception):
        du.relative_poses_to_first_frame(torch.zeros(1, 16, 4, 4, dtype=torch.float64))
    with pytest.raises(ValueError):
        du.relative_poses_to_first_frame(torch.zeros(1, 16, 4, 4, device=cuda_device))
