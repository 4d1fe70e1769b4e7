"""Mirror of fmc/pipelines/pipeline_animation_cm_om.py: `CameraObjCtrlPipeline` (:442-738) = CameraCtrlPipeline +
ObjectEncoder features: zeros for the unconditional CFG half (:671-676), dropped once t < omcm_min_step (:682-685),
single window only (:690)."""
import torch

from .pipeline_animation import AnimationPipelineOutput, CameraCtrlPipeline  # noqa: F401


class CameraObjCtrlPipeline(CameraCtrlPipeline):
    _accepts_traj = True

    @torch.no_grad()
    def __call__(self, prompt, pose_embedding, video_length, traj_features=None, height=None, width=None,
                 num_inference_steps=50, guidance_scale=7.5, negative_prompt=None, num_videos_per_prompt=1, eta=0.0,
                 generator=None, latents=None, output_type="tensor", return_dict=True, callback=None,
                 callback_steps=1, multidiff_total_steps=1, multidiff_overlaps=12, **kwargs):
        """Parameter order and defaults of the reference's CameraObjCtrlPipeline.__call__
        (pipeline_animation_cm_om.py:570-595); `omcm_min_step` (and `prompt_embeds`, `max_steps`) ride in **kwargs."""
        return self._sample(prompt, pose_embedding, video_length, traj_features, height, width, num_inference_steps,
                            guidance_scale, negative_prompt, num_videos_per_prompt, eta, generator, latents, output_type,
                            return_dict, callback, callback_steps, multidiff_total_steps, multidiff_overlaps, **kwargs)
