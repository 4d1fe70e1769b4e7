"""Build-container check (needs /root/reference; not part of the test suite, like make_golden.py): execute the import block + definitions of the REAL trainers under the drop-in
(third-party packages that are not installable offline are stubbed), without running main()."""
import runpy, sys, types
import torch, torchvision, transformers
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

class _Any:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Any()
    def __getattr__(self, n): return _Any()
    def __mro_entries__(self, bases): return (object,)
def stub(name, **attrs):
    m = types.ModuleType(name); m.__dict__.update(attrs); m.__file__ = "<stub>"
    def ga(n):
        if n.startswith("__"): raise AttributeError(n)
        return _Any()
    m.__getattr__ = ga; sys.modules[name] = m; return m
for n in ["omegaconf", "diffusers", "diffusers.optimization", "diffusers.utils", "diffusers.models", "diffusers.models.attention_processor",
          "diffusers.utils.import_utils", "decord", "cv2", "imageio", "nltk", "nltk.stem", "wandb", "termcolor"]:
    if n not in sys.modules:
        try:
            __import__(n)
        except Exception:
            stub(n)
from synfmc_b200 import dropin
dropin.install(reference_root="/root/reference")
for script in ("train_cam_ctrl.py", "train_cam_obj_ctrl.py"):
    ns = runpy.run_path("/root/reference/" + script, run_name="imported_not_main")
    names = ["setup_logger", "format_time", "save_videos_grid", "ray_condition", "create_absolute_matrix_from_ref_cam_list",
             "UnrealTrajVideoDataset", "UnrealTrajLoraDataset", "CustomizedAttnProcessor"]
    names += ["CameraCtrlPipeline", "UNet3DConditionModelPoseCond", "CameraPoseEncoder", "PoseAdaptor"] if script == "train_cam_ctrl.py" else \
             ["CameraObjCtrlPipeline", "UNet3DConditionModelCamObjCond", "CameraPoseEncoder", "CamObjPoseAdaptor", "Adapter", "get_traj_features_v2",
              "Adapted_CrossAttnDownBlock3D_forward", "Adapted_DownBlock3D_forward"]
    print(script)
    for n in names:
        print("   ", n, "<-", getattr(ns[n], "__module__", "?"))
