"""Host-side helpers of the mirror package (`fmc.data.utils`, `fmc.utils.util` -- the last row of SURVEY 8b) against
values produced by running the reference's own functions (tests/golden/make_golden_utils.py)."""
import logging
import os

import numpy as np
import torch

from synfmc_b200.fmc.data import utils as du
from synfmc_b200.fmc.utils import util as uu
from tests.golden.make_golden_utils import utils_inputs

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fmc_reference_utils.pt"),
                  weights_only=False)


def test_relative_and_absolute_camera_matrices():
    inp = utils_inputs()
    rel = du.create_relative_matrix_of_cam_list(inp["cams"], scale_T=1200)
    assert rel.dtype == GOLD["relative_cam_list"].dtype and rel.shape == (16, 12)
    assert torch.equal(rel[0], torch.eye(3, 4, dtype=rel.dtype).reshape(-1))
    np.testing.assert_allclose(rel.numpy(), GOLD["relative_cam_list"].numpy(), rtol=1e-12, atol=1e-12)
    absol = du.create_absolute_matrix_from_ref_cam_list(inp["cams"][0].numpy(), rel.reshape(16, 3, 4).numpy(), scale_T=1200)
    assert isinstance(absol, list) and len(absol) == 16 and absol[0].shape == (3, 4)
    np.testing.assert_allclose(np.stack(absol), GOLD["absolute_from_ref"], rtol=1e-10, atol=1e-9)
    try:
        du.create_absolute_matrix_from_ref_cam_list(inp["cams"][0].numpy(), rel.reshape(16, 3, 4).numpy()[:8])
    except AssertionError:
        pass
    else:
        raise AssertionError("a clip that is not 16 frames long must be rejected like the reference does")


def test_object_pose_relative_to_camera_keeps_the_reference_quirk():
    inp = utils_inputs()
    got = du.create_relative_matrix_of_two_torch_matrix(inp["cams"][3], inp["objs"], scale_T=1000)
    assert isinstance(got, np.ndarray) and got.shape == (3, 12)
    np.testing.assert_allclose(got, GOLD["relative_two"], rtol=1e-12, atol=1e-12)
    one = du.create_relative_matrix_of_two_torch_matrix(inp["cams"][5], inp["objs"][:1], scale_T=1000)
    np.testing.assert_allclose(one, GOLD["relative_two_single"], rtol=1e-12, atol=1e-12)
    # the quirk: every object's translation is derived from object 0's translation (np.dot over stacked arrays)
    cam = inp["cams"][3].numpy()[:3]
    obj = inp["objs"].numpy()[:, :3]
    for i in range(3):
        want_t = obj[i, :, :3].T @ (cam[:, 3] - obj[0, :, 3]) / 1000
        np.testing.assert_allclose(got[i].reshape(3, 4)[:, 3], want_t, rtol=1e-12, atol=1e-12)
    # inputs are not modified
    assert torch.equal(inp["objs"], utils_inputs()["objs"])


def test_format_time():
    inp = utils_inputs()
    assert [uu.format_time(t) for t in inp["times"]] == GOLD["format_time"]


def test_video_grid_frames_match_torchvision_grid_of_the_reference(tmp_path):
    inp = utils_inputs()
    grids = GOLD["grids"]
    got = np.stack(uu.video_grid_frames(inp["videos"], n_rows=2))
    assert got.dtype == np.uint8 and np.array_equal(got, grids[0])
    got = np.stack(uu.video_grid_frames(inp["videos_signed"], rescale=True))
    assert np.array_equal(got, grids[1])
    got = np.stack(uu.video_grid_frames(inp["videos"]))
    assert np.array_equal(got, grids[2])
    path = tmp_path / "sub" / "grid.gif"
    uu.save_videos_grid(inp["videos"], str(path), n_rows=2, fps=4)
    assert path.exists() and path.stat().st_size > 0


def test_setup_logger_files_and_idempotence(tmp_path, capsys):
    out = str(tmp_path / "run")
    lg = uu.setup_logger(out, 0, color=False, name="fmc_test_logger")
    assert uu.setup_logger(out, 0, color=False, name="fmc_test_logger") is lg and len(lg.handlers) == 2
    assert lg.level == logging.DEBUG and lg.propagate is False
    lg.info("hello rank 0")
    for h in lg.handlers:
        h.flush()
    assert "hello rank 0" in capsys.readouterr().out
    assert "INFO: hello rank 0" in open(os.path.join(out, "log.txt")).read()
    lg1 = uu.setup_logger(str(tmp_path / "run" / "train.log"), 3, name="fmc_test_logger_rank3")
    assert len(lg1.handlers) == 1   # no console handler off rank 0
    lg1.warning("from rank 3")
    lg1.handlers[0].flush()
    assert "from rank 3" in open(str(tmp_path / "run" / "train.log.rank3")).read()


def test_instantiate_from_config():
    obj = uu.instantiate_from_config({"target": "collections.OrderedDict", "kwargs": {"a": 1}})
    assert obj == {"a": 1}
    assert uu.instantiate_from_config("__is_unconditional__") is None
    try:
        uu.instantiate_from_config({"kwargs": {}})
    except KeyError:
        pass
    else:
        raise AssertionError("missing target must raise KeyError")
