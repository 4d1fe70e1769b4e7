#!/usr/bin/env python
"""Golden vectors for the denoising LOOP, produced by EXECUTING THE REFERENCE'S OWN pipelines
(/root/reference/fmc/pipelines/pipeline_animation_cm_om.py `CameraObjCtrlPipeline.__call__`, :570-740, and
pipeline_animation.py `CameraCtrlPipeline.__call__` with two multidiff windows) on the tiny U-Net of make_golden.py.

Shimmed: the diffusers layer (as in make_golden.py; `DiffusionPipeline` reduced to register_modules / progress_bar /
device, `DDIMScheduler` = the restated one), and the frozen third-party nets around the loop -- a hash tokenizer, an
embedding-table "text encoder" and a nearest-upsampling "VAE" (stand-ins defined below and re-used by the tests, so the
prompt -> embeddings and latents -> video edges are identical on both sides).  Everything between them -- CFG batch
doubling, zero object features for the unconditional half, the omcm_min_step gate, per-window U-Net calls, window
averaging, scheduler stepping, callback protocol, video post-processing -- is the reference's code.

    python tests/golden/make_golden_pipeline.py      # needs /root/reference; writes tests/golden/fmc_reference_pipeline.pt
"""
import contextlib
import importlib
import os
import sys
import types
from types import SimpleNamespace

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if HERE not in sys.path:
    sys.path.insert(0, HERE)


# ---------------------------------------------------------------------------------------------------------------
# stand-ins for the frozen nets around the loop (shared with tests/test_golden.py)
# ---------------------------------------------------------------------------------------------------------------
class HashTokenizer:
    model_max_length = 77

    def __call__(self, texts, padding=None, max_length=None, truncation=None, return_tensors=None):
        texts = [texts] if isinstance(texts, str) else list(texts)
        n = max_length or self.model_max_length
        rows = []
        for s in texts:
            ids = [1 + (sum(ord(c) * (i + 1) for i, c in enumerate(w)) % 997) for w in s.split()][: n - 2]
            rows.append([998] + ids + [999] + [0] * (n - 2 - len(ids)))
        return SimpleNamespace(input_ids=torch.tensor(rows, dtype=torch.long))

    def batch_decode(self, ids):
        return ["?"] * len(ids)


class TableTextEncoder(nn.Module):
    def __init__(self, dim=768, seed=21):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.table = nn.Parameter(0.5 * torch.randn(1000, dim, generator=g), requires_grad=False)
        self.pos = nn.Parameter(0.1 * torch.randn(77, dim, generator=g), requires_grad=False)
        self.config = SimpleNamespace()

    def forward(self, input_ids, attention_mask=None):
        return (self.table[input_ids] + self.pos[: input_ids.shape[1]],)


class UpsampleVAE(nn.Module):
    """decode: 4 latent channels -> 3 'colours' by a fixed mix, nearest x8"""

    def __init__(self):
        super().__init__()
        self.config = SimpleNamespace(block_out_channels=(1, 1, 1, 1))
        self.mix = nn.Parameter(torch.tensor([[0.6, -0.2, 0.1, 0.3], [-0.1, 0.5, 0.4, -0.2], [0.2, 0.3, -0.5, 0.4]]),
                                requires_grad=False)

    def decode(self, z):
        x = torch.einsum("oc,bchw->bohw", self.mix, z)
        return SimpleNamespace(sample=torch.nn.functional.interpolate(x, scale_factor=8, mode="nearest"))

    def enable_slicing(self):
        pass


def pipeline_inputs():
    from make_golden import golden_inputs
    from synfmc_b200 import synth
    inp = golden_inputs()
    g = torch.Generator().manual_seed(4321)
    b, f, h, w = inp["b"], inp["f"], inp["h"], inp["w"]
    inp["latents"] = torch.randn(b, 4, f, h, w, generator=g)
    # two multidiff windows of 4 frames with overlap 2 -> 6 frames in total
    K6, c2w6 = synth.synth_camera(b, 6, inp["H"], inp["W"], seed=13)
    inp["K6"], inp["c2w6"] = K6, c2w6
    inp["latents6"] = torch.randn(b, 4, 6, h, w, generator=g)
    # a training-style batch of two clips with per-sample timesteps (train_cam_obj_ctrl.py:840-866)
    inp["train_latents"] = torch.randn(2, 4, f, h, w, generator=g)
    inp["train_text"] = 0.5 * torch.randn(2, 77, 768, generator=g)
    inp["train_timesteps"] = torch.tensor([961, 41])
    inp["train_K"], inp["train_c2w"] = synth.synth_camera(2, f, inp["H"], inp["W"], seed=14)
    inp["train_obj_infos"], inp["train_obj_masks"] = synth.synth_objects(2, f, inp["H"], inp["W"], 2, seed=15, gaussian=True)
    inp["prompt"] = ["a red car drives along the coast road"]
    inp["negative_prompt"] = ["blurry low quality"]
    return inp


def _plucker(ray_condition, K, c2w, H, W):
    b, f = K.shape[:2]
    bottom = torch.tensor([0, 0, 0, 1.0]).view(1, 1, 1, 4).expand(b, f, 1, 4)
    rays = ray_condition(K, torch.cat([c2w, bottom], dim=2), H, W, device="cpu", flip_flag=torch.zeros(f, dtype=torch.bool))
    return rays.permute(0, 4, 1, 2, 3).contiguous()  # b 6 f H W


def main():
    import make_golden as mg
    from oracle import diffusers_restated as R
    from oracle import harness
    from synfmc_b200.synth import synth_init_
    ref = mg.reference_modules()

    class DiffusionPipeline:
        def register_modules(self, **kw):
            for k, v in kw.items():
                setattr(self, k, v)

        @property
        def device(self):
            return torch.device("cpu")

        @contextlib.contextmanager
        def progress_bar(self, total=None):
            yield SimpleNamespace(update=lambda *a, **k: None)

    class DDIMScheduler(R.DDIMScheduler):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.config = SimpleNamespace(steps_offset=self.steps_offset, clip_sample=self.clip_sample)

    mg._module("diffusers.pipelines", __path__=[])
    mg._module("diffusers.pipelines.pipeline_utils", DiffusionPipeline=DiffusionPipeline)
    mg._module("diffusers.schedulers", DDIMScheduler=DDIMScheduler, DPMSolverMultistepScheduler=mg._Anything,
               EulerAncestralDiscreteScheduler=mg._Anything, EulerDiscreteScheduler=mg._Anything,
               LMSDiscreteScheduler=mg._Anything, PNDMScheduler=mg._Anything)
    p_obj = importlib.import_module("fmc.pipelines.pipeline_animation_cm_om")
    p_cam = importlib.import_module("fmc.pipelines.pipeline_animation")

    inp = pipeline_inputs()
    cfg = harness.unet_config(True)
    ray_condition = ref.dataset.ray_condition if ref.dataset is not None else mg._ray_condition_from_source()
    out = {}

    # ---- CameraObjCtrlPipeline: cam + object features, CFG 8.0, 25 steps, object features only while t >= 700
    unet_o = ref.unet_obj.UNet3DConditionModelCamObjCond(**cfg)
    harness.set_processors(unet_o, cfg["block_out_channels"])
    synth_init_(unet_o, seed=0)
    idx = 0
    for _n, m in unet_o.down_blocks.named_modules():   # train_cam_obj_ctrl.py:317-329
        if m.__class__.__name__ == "CrossAttnDownBlock3D":
            m.forward = ref.modified.Adapted_CrossAttnDownBlock3D_forward.__get__(m, m.__class__)
        elif m.__class__.__name__ == "DownBlock3D":
            m.forward = ref.modified.Adapted_DownBlock3D_forward.__get__(m, m.__class__)
        else:
            continue
        m.traj_fea_idx = idx
        idx += 1
    unet_o.eval()
    enc = ref.pose.CameraPoseEncoder(channels=list(inp["channels"]), **harness.POSE_ENCODER_KWARGS)
    synth_init_(enc, seed=1)
    enc.eval()
    pipe = p_obj.CameraObjCtrlPipeline(UpsampleVAE(), TableTextEncoder(), HashTokenizer(), unet_o, DDIMScheduler(), enc)
    trace = []
    plucker = _plucker(ray_condition, inp["K"], inp["c2w"], inp["H"], inp["W"])
    res = pipe(prompt=inp["prompt"], pose_embedding=plucker, video_length=inp["f"], traj_features=inp["traj_feats"],
               height=inp["H"], width=inp["W"], num_inference_steps=25, guidance_scale=8.0,
               negative_prompt=inp["negative_prompt"], latents=inp["latents"].clone(), omcm_min_step=700,
               callback=lambda i, t, lat: trace.append((int(i), int(t), lat.clone())))
    out["obj_steps"] = [(i, t) for i, t, _ in trace]
    out["obj_latents"] = torch.stack([lat for _, _, lat in trace])
    out["obj_videos_stride8"] = res.videos[..., ::8, ::8].clone()   # the stand-in VAE upsamples x8 by repetition
    out["obj_videos_shape"] = tuple(res.videos.shape)
    emb = pipe._encode_prompt(inp["prompt"], torch.device("cpu"), 1, True, inp["negative_prompt"])
    out["text_embeddings_head"] = emb[:, :12, :8].clone()            # order check: [uncond (negative prompt) ++ cond]

    # ---- CameraCtrlPipeline: cam only, two overlapping windows (4 frames, overlap 2 -> 6 frames), 6 steps, CFG 7.5
    unet_c = ref.unet.UNet3DConditionModelPoseCond(**cfg)
    harness.set_processors(unet_c, cfg["block_out_channels"])
    synth_init_(unet_c, seed=0)
    unet_c.eval()
    pipe_c = p_cam.CameraCtrlPipeline(UpsampleVAE(), TableTextEncoder(), HashTokenizer(), unet_c, DDIMScheduler(), enc)
    trace_c = []
    plucker6 = _plucker(ray_condition, inp["K6"], inp["c2w6"], inp["H"], inp["W"])
    res_c = pipe_c(prompt=inp["prompt"], pose_embedding=plucker6, video_length=4, height=inp["H"], width=inp["W"],
                   num_inference_steps=6, guidance_scale=7.5, latents=inp["latents6"].clone(), multidiff_total_steps=2,
                   multidiff_overlaps=2, callback=lambda i, t, lat: trace_c.append((int(i), int(t), lat.clone())))
    out["cam_steps"] = [(i, t) for i, t, _ in trace_c]
    out["cam_latents"] = torch.stack([lat for _, _, lat in trace_c])
    out["cam_videos_stride8"] = res_c.videos[..., ::8, ::8].clone()
    out["cam_videos_shape"] = tuple(res_c.videos.shape)
    # ---- the same windows with the pose embedding given as a per-window LIST (:644-651, :678-679): the only form the
    # reference can run beyond 16 frames (the CameraEncoder's positional encoding has max_len 16).  Each window's
    # embedding is encoded separately, so its temporal attention sees that window's 4 frames only.
    trace_l = []
    windows = [plucker6[:, :, 0:4].contiguous(), plucker6[:, :, 2:6].contiguous()]
    pipe_c(prompt=inp["prompt"], pose_embedding=windows, video_length=4, height=inp["H"], width=inp["W"],
           num_inference_steps=6, guidance_scale=7.5, latents=inp["latents6"].clone(), multidiff_total_steps=2,
           multidiff_overlaps=2, callback=lambda i, t, lat: trace_l.append((int(i), int(t), lat.clone())))
    out["cam_list_latents"] = torch.stack([lat for _, _, lat in trace_l])
    # ---- the training forward: get_traj_features_v2 with random nulling (python `random`, seeded) -> CamObjPoseAdaptor
    import random
    omcm = ref.adapter.Adapter(channels=list(inp["channels"]), **harness.OMCM_KWARGS)
    synth_init_(omcm, seed=2)
    omcm.eval()
    wrapper = importlib.import_module("fmc.models.pose_obj_adaptor").CamObjPoseAdaptor(unet_o, enc)
    plucker2 = _plucker(ray_condition, inp["train_K"], inp["train_c2w"], inp["H"], inp["W"])
    with torch.no_grad():
        random.seed(1)   # first draw 0.134 -> clip 0 loses its object features, second 0.847 -> clip 1 keeps them
        trajs = ref.util.get_traj_features_v2(inp["train_obj_infos"], inp["train_obj_masks"], omcm, True, 0.5, [False, False],
                                              "cpu", torch.float32)
        out["train_traj_features_c8"] = [t[:, ::8].clone() for t in trajs]   # every 8th channel
        out["train_noise_pred"] = wrapper(inp["train_latents"], inp["train_timesteps"], inp["train_text"], plucker2,
                                          trajs).clone()
        wrapper_c = ref.pose.PoseAdaptor(unet_c, enc)
        out["train_noise_pred_cam"] = wrapper_c(inp["train_latents"], inp["train_timesteps"], inp["train_text"],
                                                plucker2).clone()

    import inspect

    def signature(fn):
        return [(n, None if q.default is inspect.Parameter.empty else q.default, q.kind.name)
                for n, q in inspect.signature(fn).parameters.items() if n != "self"]
    out["call_signature_obj"] = signature(p_obj.CameraObjCtrlPipeline.__call__)
    out["call_signature_cam"] = signature(p_cam.CameraCtrlPipeline.__call__)
    out["init_signature_obj"] = [n for n, _, _ in signature(p_obj.CameraObjCtrlPipeline.__init__)]
    path = os.path.join(HERE, "fmc_reference_pipeline.pt")
    torch.save(out, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KB)", out["obj_steps"][:3], out["obj_latents"].shape,
          out["obj_videos_shape"], out["cam_steps"], out["cam_latents"].shape, out["cam_videos_shape"])


if __name__ == "__main__":
    main()
