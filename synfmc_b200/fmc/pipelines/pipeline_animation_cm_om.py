"""Mirror of fmc/pipelines/pipeline_animation_cm_om.py: `CameraObjCtrlPipeline` (:442-738) = CameraCtrlPipeline +
ObjectEncoder features: zeros for the unconditional CFG half (:671-676), dropped once t < omcm_min_step (:682-685),
single window only (:690)."""
from .pipeline_animation import AnimationPipelineOutput, CameraCtrlPipeline  # noqa: F401


class CameraObjCtrlPipeline(CameraCtrlPipeline):
    _accepts_traj = True
