"""Level-1 integration: `import fmc` must bind the B200 mirror even when the script being run lives in
a reference checkout that has its own `fmc/` (sys.path[0]), and the names of `fmc` outside the hot path must come
from that checkout, unmodified.  Run in subprocesses against a miniature stand-in checkout (the real one needs
diffusers / decord / cv2 / nltk at import time)."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write(path, text):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        fh.write(textwrap.dedent(text))


def _fake_checkout(tmp_path):
    ref = str(tmp_path / "SynFMC")
    _write(f"{ref}/fmc/__init__.py", "")
    _write(f"{ref}/fmc/data/__init__.py", "")
    _write(f"{ref}/fmc/data/utils.py", """
        def helper():
            return "reference utils"
    """)
    _write(f"{ref}/fmc/data/dataset.py", """
        from .utils import *
        class UnrealTrajVideoDataset:
            origin = helper()
        class UnrealTrajLoraDataset:
            pass
        def ray_condition(*a, **k):
            return "reference ray_condition"
    """)
    _write(f"{ref}/fmc/models/__init__.py", "")
    _write(f"{ref}/fmc/models/unet.py", """
        class UNet3DConditionModelPoseCond:
            origin = "reference"
    """)
    _write(f"{ref}/fmc/extras.py", 'VALUE = "reference-only module"')
    _write(f"{ref}/train_x.py", """
        import json, sys
        from fmc.utils.util import setup_logger, format_time, save_videos_grid
        from fmc.models.unet import UNet3DConditionModelPoseCond
        from fmc.models.attention_processor import AttnProcessor as CustomizedAttnProcessor
        from fmc.data.dataset import UnrealTrajVideoDataset, UnrealTrajLoraDataset, ray_condition
        from fmc.data.utils import create_absolute_matrix_from_ref_cam_list
        from fmc.adapter import Adapter
        from fmc.util import get_traj_features_v2
        import fmc.extras
        import fmc, synfmc_b200.fmc, synfmc_b200.fmc.models.unet as real_unet
        print(json.dumps({
            "unet": UNet3DConditionModelPoseCond.__module__,
            "same_module_object": sys.modules["fmc.models.unet"] is real_unet and fmc is synfmc_b200.fmc,
            "ray": ray_condition.__module__,
            "dataset": UnrealTrajVideoDataset.__module__, "dataset_origin": UnrealTrajVideoDataset.origin,
            "abs": create_absolute_matrix_from_ref_cam_list.__module__,
            "extras": fmc.extras.VALUE, "argv": sys.argv[1:], "time": format_time(61.5),
            "name": __name__}))
    """)
    return ref


def _run(args, cwd, env_extra=None):
    env = dict(os.environ, PYTHONPATH=ROOT)
    env.pop("FMC_REFERENCE_ROOT", None)
    env.update(env_extra or {})
    return subprocess.run([sys.executable] + args, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          text=True, timeout=300)


def test_launcher_binds_the_mirror_and_falls_through_to_the_checkout(tmp_path):
    ref = _fake_checkout(tmp_path)
    proc = _run(["-m", "synfmc_b200.launch", os.path.join(ref, "train_x.py"), "--config", "configs/obj.yaml"], cwd=ref)
    assert proc.returncode == 0, proc.stderr[-2000:]
    got = json.loads(proc.stdout.strip().splitlines()[-1])
    assert got["unet"] == "synfmc_b200.fmc.models.unet" and got["same_module_object"] is True
    assert got["ray"] == "synfmc_b200.fmc.data.dataset"
    assert got["dataset"] == "_fmc_reference.data.dataset" and got["dataset_origin"] == "reference utils"
    assert got["abs"] == "synfmc_b200.fmc.data.utils"
    assert got["extras"] == "reference-only module"
    assert got["argv"] == ["--config", "configs/obj.yaml"] and got["name"] == "__main__"
    assert got["time"] == "1 minutes 1.50 seconds"


def test_without_a_checkout_the_missing_names_say_what_to_do(tmp_path):
    code = ("from synfmc_b200 import dropin; dropin.install(); import fmc.data.dataset as d\n"
            "from fmc.data.dataset import ray_condition\n"
            "try:\n    d.UnrealTrajVideoDataset\nexcept AttributeError as e:\n    print('MSG', e)\n"
            "try:\n    import fmc.extras\nexcept ModuleNotFoundError as e:\n    print('MISSING', e)\n")
    proc = _run(["-c", code], cwd=str(tmp_path))
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert "FMC_REFERENCE_ROOT" in proc.stdout and "MISSING" in proc.stdout


def test_install_after_a_foreign_fmc_import_is_refused(tmp_path):
    ref = _fake_checkout(tmp_path)
    code = ("import fmc\nfrom synfmc_b200 import dropin\n"
            "try:\n    dropin.install()\nexcept RuntimeError as e:\n    print('REFUSED', e)\n")
    proc = _run(["-c", code], cwd=ref, env_extra={"PYTHONPATH": ROOT + os.pathsep + ref})
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert "REFUSED" in proc.stdout
