"""The call surface of the mirror pipelines against the reference's (signatures recorded by
tests/golden/make_golden_pipeline.py from the reference classes), and the routing of every argument of `__call__` into
the loop -- with the three stages that need the GPU (`_pose_features`, `_traj_features`, `denoise`) replaced by recorders."""
import inspect
import os

import pytest
import torch

from synfmc_b200.fmc._blocks import DDIMScheduler
from synfmc_b200.fmc.pipelines.pipeline_animation import CameraCtrlPipeline
from synfmc_b200.fmc.pipelines.pipeline_animation_cm_om import CameraObjCtrlPipeline
from tests.golden.make_golden_pipeline import HashTokenizer, TableTextEncoder, UpsampleVAE

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fmc_reference_pipeline.pt"),
                  weights_only=False)


def _signature(fn):
    return [(n, None if q.default is inspect.Parameter.empty else q.default, q.kind.name)
            for n, q in inspect.signature(fn).parameters.items() if n != "self"]


def test_call_and_init_signatures_equal_the_reference():
    assert _signature(CameraObjCtrlPipeline.__call__) == GOLD["call_signature_obj"]
    assert _signature(CameraCtrlPipeline.__call__) == GOLD["call_signature_cam"]
    assert [n for n, _, _ in _signature(CameraObjCtrlPipeline.__init__)] == GOLD["init_signature_obj"]


class _Unet:
    in_channels = 4
    config = type("C", (), {"sample_size": 8})()


def _recording(cls):
    class Rec(cls):
        def _pose_features(self, pose_embedding, do_cfg):
            self.seen = {"pose": pose_embedding, "pose_cfg": do_cfg}
            return ["pose-features"]

        def _traj_features(self, traj_features, do_cfg):
            self.seen["traj_in"] = traj_features
            return None if traj_features is None else ["traj-features"]

        def denoise(self, latents, text_embeddings, pose_features, video_length, **kw):
            self.seen.update(latents=latents, text=text_embeddings, pose_features=pose_features, L=video_length, **kw)
            if kw.get("callback") is not None:
                kw["callback"](0, 961, latents)
            return latents + 1.0
    return Rec(UpsampleVAE(), TableTextEncoder(), HashTokenizer(), _Unet(), DDIMScheduler(), object())


def test_positional_call_of_the_camera_pipeline_binds_like_the_reference():
    pipe = _recording(CameraCtrlPipeline)
    pose = torch.zeros(1, 6, 6, 64, 64)
    lat = torch.zeros(1, 4, 6, 8, 8)
    calls = []
    # reference order: prompt, pose_embedding, video_length, height, width, num_inference_steps, guidance_scale
    out = pipe(["a cat"], pose, 4, 64, 64, 6, 7.5, latents=lat, multidiff_total_steps=2, multidiff_overlaps=2,
               callback=lambda i, t, x: calls.append((i, t)), max_steps=3)
    s = pipe.seen
    assert s["pose"] is pose and s["pose_cfg"] is True and s["traj_in"] is None
    assert s["L"] == 4 and s["num_inference_steps"] == 6 and s["guidance_scale"] == 7.5
    assert s["multidiff_total_steps"] == 2 and s["multidiff_overlaps"] == 2 and s["max_steps"] == 3
    assert s["traj_features"] is None and s["omcm_min_step"] is None
    assert s["text"].shape == (2, 77, 768)                      # [uncond ("") ++ cond] from the attached text encoder
    assert s["latents"].shape == (1, 4, 6, 8, 8) and calls == [(0, 961)]
    assert torch.equal(out.latents, lat + 1.0) and out.videos.shape == (1, 3, 6, 64, 64)
    assert out.videos.dtype == torch.float32 and out.videos.device.type == "cpu"
    with pytest.raises(TypeError):
        pipe(["a cat"], pose, 4, traj_features=[torch.zeros(1)])
    with pytest.raises(ValueError):
        pipe(["a cat"], pose, 4, 64, 64, latents=torch.zeros(1, 4, 5, 8, 8))   # 5 frames where 4 are announced


def test_positional_call_of_the_object_pipeline_and_its_kwargs():
    pipe = _recording(CameraObjCtrlPipeline)
    pose = torch.zeros(1, 6, 4, 64, 64)
    trajs = [torch.zeros(1, 320, 4, 8, 8)]
    embeds = torch.randn(2, 77, 768)
    # reference order: prompt, pose_embedding, video_length, traj_features, height, width
    out = pipe(None, pose, 4, trajs, 64, 64, num_inference_steps=25, guidance_scale=8.0, omcm_min_step=700,
               prompt_embeds=embeds, return_dict=False)
    s = pipe.seen
    assert s["traj_in"] is trajs and s["traj_features"] == ["traj-features"] and s["omcm_min_step"] == 700
    assert s["text"] is embeds or torch.equal(s["text"], embeds)
    assert s["L"] == 4 and s["multidiff_total_steps"] == 1
    assert isinstance(out, torch.Tensor) and out.shape == (1, 3, 4, 64, 64)   # return_dict=False -> the video
    # latents default to seeded noise of the announced size, scaled by init_noise_sigma = 1
    g = torch.Generator().manual_seed(3)
    pipe(None, pose, 4, None, 64, 64, guidance_scale=1.0, generator=g, prompt_embeds=embeds[:1])
    assert pipe.seen["latents"].shape == (1, 4, 4, 8, 8) and pipe.seen["pose_cfg"] is False
    want = torch.randn((1, 4, 4, 8, 8), generator=torch.Generator().manual_seed(3))
    assert torch.equal(pipe.seen["latents"], want)


def test_a_diffusers_style_scheduler_object_drives_the_update():
    """The trainers build diffusers' own DDIMScheduler and pass it in: the pipeline reads alpha-bar from its public
    attributes (alphas_cumprod, final_alpha_cumprod, num_inference_steps, config.num_train_timesteps)."""
    from types import SimpleNamespace

    from synfmc_b200.fmc.pipelines.pipeline_animation import ddim_alphas
    ours = DDIMScheduler()
    ours.set_timesteps(25)

    class DiffusersLike:   # what diffusers.DDIMScheduler exposes, nothing of the mirror's
        def __init__(self):
            self.config = SimpleNamespace(num_train_timesteps=1000, prediction_type="epsilon", clip_sample=False,
                                          thresholding=False, steps_offset=1)
            betas = torch.linspace(0.00085, 0.012, 1000, dtype=torch.float32)
            self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
            self.final_alpha_cumprod = torch.tensor(1.0)
            self.num_inference_steps = 25
    theirs = DiffusersLike()
    for t in ours.timesteps.tolist():
        assert ddim_alphas(theirs, t) == ours.alphas_for(t)
    assert ddim_alphas(theirs, 1)[1] == 1.0    # last step: alpha_prev = final_alpha_cumprod
    theirs.config.prediction_type = "v_prediction"
    with pytest.raises(NotImplementedError):
        ddim_alphas(theirs, 961)
    with pytest.raises(NotImplementedError):
        ddim_alphas(SimpleNamespace(timesteps=[1]), 1)


def test_prompt_encoding_and_video_decoding_equal_the_reference_pipeline():
    """The two frozen-network edges of `__call__` on the stand-in CLIP / VAE: [negative ++ prompt] embedding order
    (pipeline_animation_cm_om.py:480-567) and decode_latents (per-frame decode, /0.18215, (x/2+0.5).clamp, fp32 on the
    CPU, :465-478), against what the reference pipeline produced from the same stand-ins."""
    from tests.golden.make_golden_pipeline import pipeline_inputs
    inp = pipeline_inputs()
    pipe = CameraObjCtrlPipeline(UpsampleVAE(), TableTextEncoder(), HashTokenizer(), _Unet(), DDIMScheduler(), object())
    emb = pipe._encode_prompt(inp["prompt"], torch.device("cpu"), True, inp["negative_prompt"])
    assert emb.shape == (2, 77, 768) and torch.equal(emb[:, :12, :8], GOLD["text_embeddings_head"])
    assert pipe._encode_prompt(inp["prompt"], torch.device("cpu"), False, None).shape == (1, 77, 768)
    empty = pipe._encode_prompt(inp["prompt"], torch.device("cpu"), True, None)       # negative prompt defaults to ""
    assert torch.equal(empty[1], emb[1]) and not torch.equal(empty[0], emb[0])
    video = pipe.decode_latents(GOLD["obj_latents"][-1])
    assert tuple(video.shape) == GOLD["obj_videos_shape"] and video.dtype == torch.float32
    assert torch.equal(video[..., ::8, ::8], GOLD["obj_videos_stride8"])
    video6 = pipe.decode_latents(GOLD["cam_latents"][-1])
    assert torch.equal(video6[..., ::8, ::8], GOLD["cam_videos_stride8"])
