"""Mirror of fmc/pipelines/pipeline_animation.py: `CameraCtrlPipeline` (:442-719).  The denoising loop (:661-707) --
CFG batch doubling, per-window U-Net, CFG combine, multidiff window averaging, DDIM update -- runs on the libfmc_b200
kernels; prompt encoding (CLIP) and VAE decode are frozen third-party networks outside the hot path (SURVEY.md
section 2, rows 13) and are used as passed in (or skipped: pass `prompt_embeds`, read `.latents`)."""
from types import SimpleNamespace

import torch

from ... import _cabi, engine, ops
from ...engine import CL, TextCtx
from ..models.pose_adaptor import unshuffle8_to_cl


class AnimationPipelineOutput(SimpleNamespace):
    pass


def _slice_frames(feat, start, length):
    t = feat.t[:, start:start + length]
    return feat if t.shape[1] == feat.t.shape[1] else CL(t.contiguous())


def ddim_alphas(scheduler, t):
    """(alpha_bar_t, alpha_bar_prev) of the DDIM update (eta = 0) that fmc_cfg_ddim_step_f32 applies, from the mirror's
    own DDIMScheduler or from ANY object with diffusers' DDIMScheduler attributes -- the trainers construct diffusers'
    scheduler themselves and hand it to the pipeline (train_cam_ctrl.py:220, :480).  Same arithmetic as
    DDIMScheduler.step: prev = t - num_train_timesteps // num_inference_steps; alpha_prev = final_alpha_cumprod
    below 0.  Anything that is not an epsilon-prediction DDIM without sample clipping is refused."""
    if hasattr(scheduler, "alphas_for"):
        return scheduler.alphas_for(t)
    cfg = getattr(scheduler, "config", None)

    def conf(name, default=None):
        if cfg is not None:
            if isinstance(cfg, dict) and name in cfg:
                return cfg[name]
            if hasattr(cfg, name):
                return getattr(cfg, name)
        return getattr(scheduler, name, default)
    if not all(hasattr(scheduler, a) for a in ("alphas_cumprod", "final_alpha_cumprod", "num_inference_steps")):
        raise NotImplementedError(f"{type(scheduler).__name__}: the denoising step fuses the DDIM update; pass a DDIMScheduler")
    if conf("prediction_type", "epsilon") != "epsilon" or conf("clip_sample", False) or conf("thresholding", False):
        raise NotImplementedError("only epsilon-prediction DDIM without clipping / thresholding is on the hot path")
    t = int(t)
    prev_t = t - int(conf("num_train_timesteps")) // int(scheduler.num_inference_steps)
    a_t = float(scheduler.alphas_cumprod[t])
    a_prev = float(scheduler.alphas_cumprod[prev_t]) if prev_t >= 0 else float(scheduler.final_alpha_cumprod)
    return a_t, a_prev


class _StepGraph:
    """The CFG-doubled U-Net forward of one denoising step, captured once into a CUDA graph for fixed shapes.

    A step is ~830 launches of 5-200 us kernels; issued one by one from Python the deep (small-token) levels are
    launch-bound.  The graph owns static input buffers (latents, timestep, text, pose / object features); `run` copies
    the caller's tensors in (device-to-device, < 0.3 % of a step), replays, and returns the static noise prediction.
    Tensor maps and scalar kernel arguments are frozen at capture time, which is why everything that changes per step
    lives in device memory."""

    def __init__(self, unet, latents, text, feats, traj, do_cfg, accepts_traj):
        dev = latents.device
        self.unet, self.do_cfg, self.accepts_traj = unet, do_cfg, accepts_traj
        self.lat = torch.empty_like(latents)
        self.text_ctx = None   # engine.TextCtx in static buffers: text rows + the K | V of every cross-attention
        self._text_seen = None
        self.t = torch.zeros(1, device=dev, dtype=torch.float32)
        self.feats = [CL(torch.empty_like(f.t)) for f in feats]
        self.traj = None if traj is None else [CL(torch.empty_like(f.t)) for f in traj]
        self._loaded = {}
        self._load(latents, text, feats, traj)
        self.t.fill_(1.0)
        # eager warm-up on a side stream: builds the weight plans, sets kernel attributes, lets cuDNN pick algorithms
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._forward()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        n0 = _cabi.launch_count
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.eps = self._forward()
        self.launches = _cabi.launch_count - n0
        _cabi.launch_count = n0  # capture launched nothing; replays are counted in run()

    def _forward(self):
        x_in = torch.cat([self.lat] * 2) if self.do_cfg else self.lat
        kw = {"traj_features": self.traj} if self.accepts_traj else {}
        return self.unet(x_in, self.t, encoder_hidden_states=self.text_ctx, pose_embedding_features=self.feats, **kw).sample

    def _load(self, latents, text, feats, traj):
        self.lat.copy_(latents)
        # the text is constant over a denoising loop: its 16 K | V projections run when it CHANGES (another tensor, or an
        # in-place write since the last step), outside the graph, into the buffers the graph reads
        if self._text_seen is None or self._text_seen[0] is not text or self._text_seen[1] != text._version:
            if self.text_ctx is None:
                self.text_ctx = self.unet.text_context(text, text.device)
            else:
                self.text_ctx.reload(text)
            self._text_seen = (text, text._version)
        # features are constant over a denoising loop: copy only when the caller hands in a different tensor object or
        # has modified it in place since the last step (torch bumps `_version` on every in-place write).  The source
        # tensor is kept referenced, so its address cannot be recycled for other data behind our back.
        pairs = list(zip(self.feats, feats)) + (list(zip(self.traj, traj)) if self.traj is not None else [])
        for slot, (dst, src) in enumerate(pairs):
            if dst.t.data_ptr() == src.t.data_ptr():
                continue
            seen = self._loaded.get(slot)
            if seen is None or seen[0] is not src.t or seen[1] != src.t._version:
                dst.t.copy_(src.t)
                self._loaded[slot] = (src.t, src.t._version)

    def run(self, latents, t, text, feats, traj):
        self._load(latents, text, feats, traj)
        self.t.fill_(float(t))
        self.graph.replay()
        _cabi.launch_count += self.launches
        return self.eps


class CameraCtrlPipeline:
    _accepts_traj = False
    use_cuda_graph = True   # single-window steps replay a captured graph (see _StepGraph); False = launch kernel by kernel
    max_graphs = 4

    def __init__(self, vae, text_encoder, tokenizer, unet, scheduler, pose_encoder):
        self.vae, self.text_encoder, self.tokenizer = vae, text_encoder, tokenizer
        self.unet, self.scheduler, self.pose_encoder = unet, scheduler, pose_encoder
        self.vae_scale_factor = 8
        self._graphs = {}

    def _step_graph(self, latents, text, feats, traj, do_cfg):
        # a graph freezes device pointers into the U-Net's weight plans: the key carries the U-Net's identity, its plan
        # generation (bumped when refresh_plans sees changed parameters / processors) and the precision mode
        engine.refresh_plans(self.unet, min_interval_s=0.25)  # exact at the start of every loop (denoise), throttled per step
        key = (tuple(latents.shape), tuple(text.shape), do_cfg, tuple(f.dims for f in feats),
               None if traj is None else tuple(f.dims for f in traj), latents.device, id(self.unet),
               engine.generation(self.unet), engine.get_precision())
        for stale in [k for k in self._graphs if k[-3] == id(self.unet) and k[-2] != engine.generation(self.unet)]:
            del self._graphs[stale]  # graphs of dropped plans must not be replayed, nor keep their buffers alive
        g = self._graphs.get(key)
        if g is None:
            while len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            g = self._graphs[key] = _StepGraph(self.unet, latents, text, feats, traj, do_cfg, self._accepts_traj)
        return g

    def reset_graphs(self):
        """Drop captured graphs (never required: graphs are keyed on the U-Net's plan generation and re-captured when its
        weights or processors change; kept for callers that want the memory back)."""
        self._graphs.clear()

    def enable_vae_slicing(self):
        if self.vae is not None and hasattr(self.vae, "enable_slicing"):
            self.vae.enable_slicing()

    def to(self, device):
        for m in (self.vae, self.text_encoder, self.unet, self.pose_encoder):
            if m is not None and hasattr(m, "to"):
                m.to(device)
        return self

    # ---- frozen third-party stages (outside the hot path) ----
    def _encode_prompt(self, prompt, device, do_cfg, negative_prompt):
        if self.text_encoder is None or self.tokenizer is None:
            raise RuntimeError("no text encoder attached: pass prompt_embeds=[(2)b, 77, 768]")
        def enc(texts):
            ids = self.tokenizer(texts, padding="max_length", max_length=self.tokenizer.model_max_length,
                                 truncation=True, return_tensors="pt").input_ids
            return self.text_encoder(ids.to(device))[0]
        cond = enc(prompt)
        if not do_cfg:
            return cond
        neg = negative_prompt if negative_prompt is not None else [""] * len(prompt)
        return torch.cat([enc(neg), cond])

    def decode_latents(self, latents):
        if self.vae is None:
            raise RuntimeError("no VAE attached: read `.latents` from the output instead of `.videos`")
        if hasattr(self.vae, "decode_video"):
            # the B200 VAE (synfmc_b200/edge/autoencoder_kl.py): all frames of a clip per pass, rearrange + /2 + 0.5 + clamp fused
            return self.vae.decode_video(latents, scaling_factor=0.18215).cpu()
        f = latents.shape[2]
        latents = (1 / 0.18215 * latents).permute(0, 2, 1, 3, 4).flatten(0, 1)
        frames = [self.vae.decode(latents[i:i + 1]).sample for i in range(latents.shape[0])]
        video = torch.cat(frames)
        video = video.reshape(-1, f, *video.shape[1:]).permute(0, 2, 1, 3, 4)
        return (video / 2 + 0.5).clamp(0, 1).cpu().float()

    def prepare_latents(self, batch_size, channels, video_length, height, width, dtype, device, generator, latents=None):
        shape = (batch_size, channels, video_length, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if latents is None:
            rand_device = "cpu" if generator is not None and generator.device.type == "cpu" else device
            latents = torch.randn(shape, generator=generator, device=rand_device, dtype=dtype).to(device)
        else:
            if latents.shape != shape:
                raise ValueError(f"Unexpected latents shape, got {latents.shape}, expected {shape}")
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    # ---- the hot loop ----
    @torch.no_grad()
    def denoise(self, latents, text_embeddings, pose_features, video_length, traj_features=None,
                num_inference_steps=25, guidance_scale=8.0, multidiff_total_steps=1, multidiff_overlaps=12,
                omcm_min_step=None, max_steps=None, callback=None, callback_steps=1, cfg_pair=None):
        """latents [b, 4, F, h, w] fp32 (device); text_embeddings [(2)b, 77, 768]; pose_features: 4 CL features over all
        F frames (already duplicated for CFG); traj_features: 4 CL features (CFG: zeros ++ features) or None.
        cfg_pair = (which, group) from synfmc_b200.shard.cfg_pair(): single-clip latency mode, this rank evaluates only
        half `which` of the CFG pair (0 = unconditional) and exchanges noise predictions once per step."""
        self.scheduler.set_timesteps(num_inference_steps)
        engine.refresh_plans(self.unet)  # weights / processors changed since the last loop -> new plans, new graphs
        L = video_length
        latents = latents.float().contiguous()
        if cfg_pair is not None:
            if not guidance_scale > 1.0 or multidiff_total_steps != 1:
                raise ValueError("cfg_pair needs classifier-free guidance (guidance_scale > 1) and a single window")
            which, b = cfg_pair[0], latents.shape[0]
            half = slice(which * b, (which + 1) * b)
            text_embeddings = text_embeddings[half].contiguous()
            pose_features = [CL(CL.from_reference(f).t[half].contiguous()) for f in pose_features]
            if traj_features is not None:  # the unconditional half carries zeros, i.e. no object features at all
                traj_features = None if which == 0 else [CL(CL.from_reference(f).t[half].contiguous()) for f in traj_features]
        if multidiff_total_steps > 1:
            # slice the per-frame features into their windows ONCE per loop (the reference re-slices every step, :678-681)
            pose_features = self._window_features(pose_features, L, multidiff_total_steps, multidiff_overlaps)
            if traj_features is not None:
                traj_features = self._window_features(traj_features, L, multidiff_total_steps, multidiff_overlaps)
        for i, t in enumerate(self.scheduler.timesteps.tolist()):
            if max_steps is not None and i >= max_steps:
                break
            step_traj = traj_features
            if omcm_min_step is not None and traj_features is not None and omcm_min_step > 0 and t < omcm_min_step:
                step_traj = None
            latents = self.denoise_step(latents, t, text_embeddings, pose_features, L, traj_features=step_traj,
                                        guidance_scale=guidance_scale, multidiff_total_steps=multidiff_total_steps,
                                        multidiff_overlaps=multidiff_overlaps, cfg_pair=cfg_pair)
            if callback is not None and i % callback_steps == 0:
                callback(i, t, latents)
        return latents

    @staticmethod
    def _is_window_list(feats):
        return isinstance(feats, (list, tuple)) and len(feats) > 0 and isinstance(feats[0], (list, tuple))

    def _window_features(self, feats, L, n_windows, overlaps):
        """4 features over all frames -> list (one entry per window) of 4 CL features over that window's L frames; a
        per-window list (the reference's `pose_embedding` list form, pipeline_animation.py:644-651,678-679) passes through."""
        if self._is_window_list(feats):
            if len(feats) != n_windows:
                raise ValueError(f"{len(feats)} per-window feature sets for {n_windows} windows")
            return [[CL.from_reference(f) for f in w] for w in feats]
        feats = [CL.from_reference(f) for f in feats]
        return [[_slice_frames(f, k * (L - overlaps), L) for f in feats] for k in range(n_windows)]

    def _window_eps(self, part, t, text_embeddings, feats, traj, do_cfg):
        """CFG-doubled U-Net on one window: graph replay when possible, kernel by kernel otherwise."""
        if (self.use_cuda_graph and _cabi.trace is None and not torch.is_tensor(t) and torch.is_tensor(text_embeddings)
                and part.dtype == torch.float32):
            return self._step_graph(part, text_embeddings, feats, traj, do_cfg).run(part, t, text_embeddings, feats, traj)
        x_in = torch.cat([part] * 2) if do_cfg else part
        kw = {"traj_features": traj} if self._accepts_traj else {}
        return self.unet(x_in, t, encoder_hidden_states=text_embeddings, pose_embedding_features=feats, **kw).sample

    @torch.no_grad()
    def denoise_step(self, latents, t, text_embeddings, pose_features, video_length, traj_features=None,
                     guidance_scale=8.0, multidiff_total_steps=1, multidiff_overlaps=12, cfg_pair=None):
        """One iteration of the loop at pipeline_animation.py:669-707 / pipeline_animation_cm_om.py:678-726: per-window
        U-Net on the CFG-doubled latents, CFG combine, window averaging, DDIM update.  `t` is a Python int taken from
        `scheduler.timesteps` after `set_timesteps`; returns the new fp32 latents.
        With cfg_pair = (which, group) the text / feature arguments are THIS RANK'S HALF of the CFG pair (as `denoise`
        prepares them): the U-Net runs at batch b, the two ranks all-gather their predictions, both apply the update."""
        do_cfg = guidance_scale > 1.0
        L = video_length
        b = latents.shape[0]
        a_t, a_prev = ddim_alphas(self.scheduler, t)
        if cfg_pair is not None:
            from ... import shard
            feats = [CL.from_reference(f) for f in pose_features]
            traj = traj_features if self._accepts_traj else None
            if traj is not None:
                traj = [CL.from_reference(f) for f in traj]
            if self.use_cuda_graph and _cabi.trace is None and not torch.is_tensor(t):
                eps_half = self._step_graph(latents, text_embeddings, feats, traj, False).run(latents, t, text_embeddings,
                                                                                               feats, traj)
            else:
                kw = {"traj_features": traj} if self._accepts_traj else {}
                eps_half = self.unet(latents, t, encoder_hidden_states=text_embeddings, pose_embedding_features=feats,
                                     **kw).sample
            e_u, e_c = shard.cfg_exchange(eps_half, cfg_pair[1])
            return ops.cfg_ddim_step(e_u, e_c, guidance_scale, latents, a_t, a_prev)
        if (self.use_cuda_graph and multidiff_total_steps == 1 and _cabi.trace is None and not torch.is_tensor(t)
                and torch.is_tensor(text_embeddings) and latents.dtype == torch.float32):
            feats = [CL.from_reference(f) for f in pose_features]
            traj = traj_features if self._accepts_traj else None
            if traj is not None:
                traj = [CL.from_reference(f) for f in traj]
            eps = self._step_graph(latents, text_embeddings, feats, traj, do_cfg).run(latents, t, text_embeddings, feats,
                                                                                     traj)
            e_u, e_c = (eps[:b], eps[b:]) if do_cfg else (eps, None)
            return ops.cfg_ddim_step(e_u, e_c, guidance_scale, latents, a_t, a_prev)
        if multidiff_total_steps == 1:
            feats = [CL.from_reference(f) for f in pose_features]
            x_in = torch.cat([latents] * 2) if do_cfg else latents
            kw = {"traj_features": traj_features} if self._accepts_traj else {}
            eps = self.unet(x_in, t, encoder_hidden_states=text_embeddings, pose_embedding_features=feats, **kw).sample
            e_u, e_c = (eps[:b], eps[b:]) if do_cfg else (eps, None)
            return ops.cfg_ddim_step(e_u, e_c, guidance_scale, latents, a_t, a_prev)
        # overlapping windows (:669-702): every window through the same captured graph, then ONE kernel that averages the
        # guided predictions per frame in the reference's order and applies the DDIM update
        n_win, stride = multidiff_total_steps, L - multidiff_overlaps
        if latents.shape[2] != (n_win - 1) * stride + L:
            raise ValueError(f"{latents.shape[2]} frames are not {n_win} windows of {L} with overlap {multidiff_overlaps}")
        win_feats = self._window_features(pose_features, L, n_win, multidiff_overlaps)
        win_traj = None
        if self._accepts_traj and traj_features is not None:
            win_traj = self._window_features(traj_features, L, n_win, multidiff_overlaps)
        eps_all = torch.empty((n_win, (2 if do_cfg else 1) * b) + tuple(latents.shape[1:2]) + (L,) + tuple(latents.shape[3:]),
                              device=latents.device, dtype=torch.float32)
        for k in range(n_win):
            part = latents[:, :, k * stride:k * stride + L].contiguous()
            eps_all[k].copy_(self._window_eps(part, t, text_embeddings, win_feats[k],
                                              None if win_traj is None else win_traj[k], do_cfg))
        return ops.window_combine_ddim(eps_all, do_cfg, guidance_scale, latents, L, stride, a_t, a_prev)

    def _pose_features(self, pose_embedding, do_cfg):
        if isinstance(pose_embedding, list):
            # one embedding PER WINDOW (pipeline_animation.py:644-651): the only form the reference can run beyond 16 frames,
            # because the CameraEncoder's positional encoding has max_len 16 (configs/cam.yaml:120)
            assert all(x.ndim == 5 for x in pose_embedding)
            return [self._pose_features(pe, do_cfg) for pe in pose_embedding]
        assert pose_embedding.ndim == 5
        feats = self.pose_encoder.encode_cl(unshuffle8_to_cl(pose_embedding.float()))
        if do_cfg:
            feats = [CL(torch.cat([f.t, f.t], dim=0)) for f in feats]
        return feats

    def _traj_features(self, traj_features, do_cfg):
        if traj_features is None:
            return None
        feats = [CL.from_reference(f) for f in traj_features]
        if do_cfg:
            feats = [CL(torch.cat([torch.zeros_like(f.t), f.t], dim=0)) for f in feats]  # uncond half gets zeros (:671-676)
        return feats

    @torch.no_grad()
    def __call__(self, prompt, pose_embedding, video_length, height=None, width=None, num_inference_steps=50,
                 guidance_scale=7.5, negative_prompt=None, num_videos_per_prompt=1, eta=0.0, generator=None,
                 latents=None, output_type="tensor", return_dict=True, callback=None, callback_steps=1,
                 multidiff_total_steps=1, multidiff_overlaps=12, **kwargs):
        """Parameter order and defaults of the reference's CameraCtrlPipeline.__call__ (pipeline_animation.py:570-593);
        extras ride in **kwargs: `prompt_embeds` ([(2)b, 77, 768], skips the text encoder), `max_steps`."""
        if "traj_features" in kwargs:
            raise TypeError("traj_features needs CameraObjCtrlPipeline")
        return self._sample(prompt, pose_embedding, video_length, None, height, width, num_inference_steps,
                            guidance_scale, negative_prompt, num_videos_per_prompt, eta, generator, latents, output_type,
                            return_dict, callback, callback_steps, multidiff_total_steps, multidiff_overlaps, **kwargs)

    @torch.no_grad()
    def _sample(self, prompt, pose_embedding, video_length, traj_features=None, height=None, width=None,
                num_inference_steps=50, guidance_scale=7.5, negative_prompt=None, num_videos_per_prompt=1, eta=0.0,
                generator=None, latents=None, output_type="tensor", return_dict=True, callback=None,
                callback_steps=1, multidiff_total_steps=1, multidiff_overlaps=12, prompt_embeds=None, **kwargs):
        assert eta == 0.0 and num_videos_per_prompt == 1
        if traj_features is not None and not self._accepts_traj:
            raise TypeError("traj_features needs CameraObjCtrlPipeline")
        device = pose_embedding[0].device if isinstance(pose_embedding, list) else pose_embedding.device
        height = height or self.unet.config.sample_size * self.vae_scale_factor
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        batch_size = latents.shape[0] if latents is not None else 1
        if isinstance(prompt, list):
            batch_size = len(prompt)
        do_cfg = guidance_scale > 1.0
        if prompt_embeds is None:
            prompt = prompt if isinstance(prompt, list) else [prompt] * batch_size
            if negative_prompt is not None and not isinstance(negative_prompt, list):
                negative_prompt = [negative_prompt] * batch_size
            prompt_embeds = self._encode_prompt(prompt, device, do_cfg, negative_prompt)
        single_len = video_length
        total_len = multidiff_total_steps * (video_length - multidiff_overlaps) + multidiff_overlaps
        latents = self.prepare_latents(batch_size, self.unet.in_channels, total_len, height, width, torch.float32,
                                       device, generator, latents)
        pose_features = self._pose_features(pose_embedding, do_cfg)
        traj = self._traj_features(traj_features, do_cfg)
        if traj is not None and multidiff_total_steps != 1 and not kwargs.get("windowed_objects"):
            # pipeline_animation_cm_om.py:690 asserts a single window.  windowed_objects=True is this package's extension
            # for BASELINE config 5 (64 frames with objects): the object features are sliced per window like the pose
            # features of :680-681
            raise AssertionError("CameraObjCtrlPipeline: multidiff_total_steps must be 1 (pass windowed_objects=True for the "
                                 "per-window extension)")
        latents = self.denoise(latents, prompt_embeds.to(device), pose_features, single_len, traj_features=traj,
                               num_inference_steps=num_inference_steps, guidance_scale=guidance_scale,
                               multidiff_total_steps=multidiff_total_steps, multidiff_overlaps=multidiff_overlaps,
                               omcm_min_step=kwargs.get("omcm_min_step"), max_steps=kwargs.get("max_steps"),
                               callback=callback,
                               callback_steps=callback_steps)
        videos = self.decode_latents(latents) if self.vae is not None else None
        out = AnimationPipelineOutput(videos=videos, latents=latents)
        return out if return_dict else (videos if videos is not None else latents)
