"""Input builders for the profiles/ scripts that must not touch oracle/ (only tests, smoke() and the CPU-baseline legs may):
the Pluecker embedding of the trainers (train_cam_ctrl.py:77-90) from the product's own ray kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def plucker_embedding(K, c2w, H, W):
    """K [b, f, 4], c2w [b, f, 3, 4] (device) -> [b, 6, f, H, W] fp32, the layout PoseAdaptor.forward takes."""
    from synfmc_b200 import ops
    b, f = K.shape[:2]
    rays = ops.plucker(K.reshape(b * f, 4), c2w.reshape(b * f, 3, 4), H, W)  # [bf, H, W, 6]
    return rays.view(b, f, H, W, 6).permute(0, 4, 1, 2, 3).contiguous()
