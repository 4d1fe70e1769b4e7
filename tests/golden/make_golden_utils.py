#!/usr/bin/env python
"""Golden values for the host-side helpers of fmc.data.utils / fmc.utils.util, produced by EXECUTING THE REFERENCE'S OWN
sources from /root/reference (read-only; `imageio` and `termcolor`, which are not installed here, are stubbed -- the
frames handed to imageio.mimsave are what gets recorded).  Writes tests/golden/fmc_reference_utils.pt (a few KB).

    python tests/golden/make_golden_utils.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"


def utils_inputs():
    """seeded inputs shared with tests/test_utils_golden.py"""
    g = np.random.default_rng(7)

    def rot():
        q, _ = np.linalg.qr(g.standard_normal((3, 3)))
        return q * np.sign(np.linalg.det(q))

    def pose44():
        m = np.eye(4)
        m[:3, :3] = rot()
        m[:3, 3] = g.standard_normal(3) * 300.0
        return m

    cams = [torch.from_numpy(pose44()) for _ in range(16)]
    objs = torch.from_numpy(np.stack([pose44() for _ in range(3)]))
    gen = torch.Generator().manual_seed(5)
    videos = torch.rand(3, 3, 2, 5, 7, generator=gen)
    videos_signed = torch.rand(1, 3, 2, 4, 4, generator=gen) * 2 - 1
    times = [0.0, 0.004, 12.3456, 59.999, 60.0, 61.5, 3600.0, 3725.25, 86400.0, 90061.7, 2 * 86400 + 59.0]
    return {"cams": cams, "objs": objs, "videos": videos, "videos_signed": videos_signed, "times": times}


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    captured = []
    sys.modules["imageio"] = types.SimpleNamespace(mimsave=lambda path, frames, fps=8: captured.append([np.array(f) for f in frames]))
    sys.modules["termcolor"] = types.SimpleNamespace(colored=lambda s, *a, **k: s)
    du = _load(os.path.join(REFERENCE, "fmc", "data", "utils.py"), "ref_fmc_data_utils")
    uu = _load(os.path.join(REFERENCE, "fmc", "utils", "util.py"), "ref_fmc_utils_util")
    inp = utils_inputs()
    out = {}
    rel = du.create_relative_matrix_of_cam_list(inp["cams"], scale_T=1200)
    out["relative_cam_list"] = rel
    out["absolute_from_ref"] = np.stack(du.create_absolute_matrix_from_ref_cam_list(
        inp["cams"][0].numpy(), rel.reshape(16, 3, 4).numpy(), scale_T=1200))
    out["relative_two"] = du.create_relative_matrix_of_two_torch_matrix(inp["cams"][3], inp["objs"], scale_T=1000)
    out["relative_two_single"] = du.create_relative_matrix_of_two_torch_matrix(inp["cams"][5], inp["objs"][:1], scale_T=1000)
    out["format_time"] = [uu.format_time(t) for t in inp["times"]]
    uu.save_videos_grid(inp["videos"], "/tmp/_golden_utils/a.gif", n_rows=2)
    uu.save_videos_grid(inp["videos_signed"], "/tmp/_golden_utils/b.gif", rescale=True)
    uu.save_videos_grid(inp["videos"], "/tmp/_golden_utils/c.gif")
    out["grids"] = [np.stack(c) for c in captured]
    torch.save(out, os.path.join(HERE, "fmc_reference_utils.pt"))
    print({k: (v.shape if hasattr(v, "shape") else len(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
