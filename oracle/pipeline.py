"""ORACLE (test infrastructure, CPU fp32; parity unpinned -- see oracle/diffusers_restated.py).

Restates the denoising loops of fmc/pipelines/pipeline_animation.py:661-707 (CameraCtrlPipeline) and
fmc/pipelines/pipeline_animation_cm_om.py:661-726 (CameraObjCtrlPipeline): CFG batch doubling, per-step U-Net,
CFG combine, multidiff window averaging, DDIM update.  Prompt encoding (CLIP) and VAE decode are outside the
hot path: text embeddings and latents are inputs/outputs here.
"""
import torch

from .pose_adaptor import _to_bcfhw


@torch.no_grad()
def denoise(unet, scheduler, pose_encoder, latents, text_embeddings, pose_embedding, video_length,
            traj_features=None, num_inference_steps=25, guidance_scale=8.0, multidiff_total_steps=1,
            multidiff_overlaps=12, omcm_min_step=None, max_steps=None, return_noise_preds=False, windowed_objects=False):
    """latents [b,4,F,h,w] with F = n*(L-overlap)+overlap; text_embeddings [2b,77,768] (uncond ++ cond) when
    guidance_scale > 1 else [b,77,768]; pose_embedding [b,6,F,H,W] or a LIST with one [b,6,L,H,W] per window
    (pipeline_animation.py:644-651,678-679); traj_features list of 4 [b,C_l,f,h_l,w_l].
    windowed_objects=True is NOT reference behaviour (pipeline_animation_cm_om.py:690 asserts one window): the object
    features are sliced per window exactly like the pose features of :680-681 -- the definition BASELINE config 5 (64
    frames with objects) is run under."""
    do_cfg = guidance_scale > 1.0
    scheduler.set_timesteps(num_inference_steps)
    timesteps = scheduler.timesteps
    L = video_length
    if traj_features is not None and not windowed_objects:
        assert multidiff_total_steps == 1  # pipeline_animation_cm_om.py:690
    per_window = isinstance(pose_embedding, list)
    if per_window:
        bs = pose_embedding[0].shape[0]
        feats = [_to_bcfhw(pose_encoder(pe), bs) for pe in pose_embedding]
        if do_cfg:
            feats = [[torch.cat([x, x], dim=0) for x in f] for f in feats]
    else:
        bs = pose_embedding.shape[0]
        feats = _to_bcfhw(pose_encoder(pose_embedding), bs)
        if do_cfg:
            feats = [torch.cat([x, x], dim=0) for x in feats]
    if do_cfg and traj_features is not None:
        traj_features = [torch.cat([torch.zeros_like(t), t], dim=0) for t in traj_features]
    preds = []
    for i, t in enumerate(timesteps):
        if max_steps is not None and i >= max_steps:
            break
        step_traj = traj_features
        if omcm_min_step is not None and traj_features is not None and omcm_min_step > 0 and t < omcm_min_step:
            step_traj = None
        noise_full = torch.zeros_like(latents)
        count = torch.zeros_like(latents)
        window_preds = []
        for k in range(multidiff_total_steps):
            s = k * (L - multidiff_overlaps)
            part = latents[:, :, s:s + L].contiguous()
            count[:, :, s:s + L] += 1
            window_feats = feats[k] if per_window else [x[:, :, s:s + L] for x in feats]
            if windowed_objects and step_traj is not None and multidiff_total_steps > 1:
                window_traj = [x[:, :, s:s + L] for x in step_traj]
            else:
                window_traj = step_traj
            x_in = torch.cat([part] * 2) if do_cfg else part
            x_in = scheduler.scale_model_input(x_in, t)
            kw = {"traj_features": window_traj} if getattr(unet, "_accepts_traj_features", False) else {}
            eps = unet(x_in, t, encoder_hidden_states=text_embeddings, pose_embedding_features=window_feats,
                       **kw).sample.to(latents.dtype)
            if do_cfg:
                e_u, e_c = eps.chunk(2)
                eps = e_u + guidance_scale * (e_c - e_u)
            window_preds.append(eps)
        for k, eps in enumerate(window_preds):
            s = k * (L - multidiff_overlaps)
            noise_full[:, :, s:s + L] += eps / count[:, :, s:s + L]
        preds.append(noise_full)
        latents = scheduler.step(noise_full, t, latents).prev_sample
    return (latents, preds) if return_noise_preds else latents
