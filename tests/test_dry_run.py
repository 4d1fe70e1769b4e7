"""Host-logic dry run on CPU: the mirror modules issue their C-ABI calls against a recorder instead of the library
(ops.DRY_RUN; no kernel executes, outputs are uninitialised), so the Python plumbing of BOTH precision modes -- plan
building, dtype routing, buffer shapes, pointer / stride arguments -- is exercised without a GPU.  What the kernels
compute is the business of the `-m gpu` parity tests."""
import ctypes

import pytest
import torch

from oracle import harness as helpers
from synfmc_b200 import _cabi, engine, ops


class Recorder:
    def __init__(self):
        self.calls = []

    def __call__(self, name, *args):
        sig = _cabi.SIGNATURES[name]
        assert len(args) == len(sig), (name, len(args), len(sig))
        for a, ty in zip(args, sig):  # every argument must convert to its declared ctypes type
            ty(a)
        self.calls.append((name, args))

    def names(self):
        return [n for n, _ in self.calls]


@pytest.fixture
def recorder(monkeypatch):
    rec = Recorder()
    monkeypatch.setattr(ops, "DRY_RUN", True)
    monkeypatch.setattr(_cabi, "call", rec)
    yield rec
    engine.set_precision("bf16")


def _inputs(b=2, f=4, h=8, w=8, channels=(320, 640), traj=False):
    g = torch.Generator().manual_seed(0)
    sample = torch.randn(b, 4, f, h, w, generator=g)
    text = torch.randn(b, 77, 768, generator=g)
    feats = [torch.randn(b, C, f, h >> l, w >> l, generator=g) for l, C in enumerate(channels)]
    trajs = [torch.randn(b, C, f, h >> l, w >> l, generator=g) for l, C in enumerate(channels)] if traj else None
    return sample, text, feats, trajs


@pytest.mark.parametrize("mode", ["bf16", "tf32", "reference"])
@pytest.mark.parametrize("obj", [False, True])
def test_tiny_unet_call_sequence(recorder, mode, obj):
    o_unet = helpers.build_oracle_unet(tiny=True, obj=obj)
    unet = helpers.build_product_unet(o_unet, tiny=True, obj=obj, device="cpu")
    sample, text, feats, trajs = _inputs(traj=obj)
    kw = {"traj_features": trajs} if obj else {}
    with engine.precision(mode):
        out = unet(sample, 961, text, pose_embedding_features=feats, **kw).sample
    assert out.shape == sample.shape and out.dtype == torch.float32
    names = set(recorder.names())
    if mode == "bf16":
        assert "fmc_gemm_bf16" in names and "fmc_spatial_attn_bf16" in names and "fmc_groupnorm_bf16" in names
        assert not any(n in names for n in ("fmc_gemm_tf32", "fmc_attention_f32", "fmc_groupnorm_f32"))
    else:
        # reference precision: NO bf16 kernel may appear anywhere in the step
        assert not any(n.endswith("_bf16") or n.endswith("_vf16") for n in names), sorted(names)
        assert {"fmc_gemm_tf32", "fmc_attention_f32", "fmc_groupnorm_f32", "fmc_layernorm_f32",
                "fmc_im2col3x3_f32"} <= names
        splits = {a[16] for n, a in recorder.calls if n == "fmc_gemm_tf32"}
        assert splits == ({3} if mode == "reference" else {1})
        assert ("fmc_split_tf32" in names) == (mode == "reference")


@pytest.mark.parametrize("mode", ["bf16", "reference"])
def test_encoders_call_sequence(recorder, mode):
    from synfmc_b200 import synth
    from synfmc_b200.fmc.util import get_traj_features_v2
    channels = (320, 640, 1280, 1280)
    enc = helpers.build_product_pose_encoder(helpers.build_oracle_pose_encoder(channels), channels, device="cpu")
    omcm = helpers.build_product_omcm(helpers.build_oracle_omcm(channels), channels, device="cpu")
    infos, masks = synth.synth_objects(1, 2, 64, 64, 2, seed=1)
    with engine.precision(mode):
        feats = enc(torch.randn(1, 6, 2, 64, 64))
        trajs = get_traj_features_v2(infos, masks, omcm, False, 0.0, None, "cpu", torch.float32)
    assert [tuple(f.shape) for f in feats] == [(2, 320, 8, 8), (2, 640, 4, 4), (2, 1280, 2, 2), (2, 1280, 1, 1)]
    assert [tuple(t.shape) for t in trajs] == [(1, 320, 2, 8, 8), (1, 640, 2, 4, 4), (1, 1280, 2, 2, 2), (1, 1280, 2, 1, 1)]
    names = set(recorder.names())
    if mode == "reference":
        assert not any(n.endswith("_bf16") for n in names), sorted(names)
        assert "fmc_mask_modulate_f32" in names and "fmc_traj_scatter_f32" in names


def test_plans_follow_parameter_changes(recorder):
    """in-place parameter updates, submodule load_state_dict and processor edits must drop the folded
    weights (and the generation CUDA graphs are keyed on must move)."""
    o_unet = helpers.build_oracle_unet(tiny=True)
    unet = helpers.build_product_unet(o_unet, tiny=True, device="cpu")
    sample, text, feats, _ = _inputs()
    unet(sample, 961, text, pose_embedding_features=feats)
    gen0 = engine.generation(unet)
    attn = unet.down_blocks[0].attentions[0].transformer_blocks[0].attn1
    plan0 = unet.down_blocks[0].attentions[0]._plan
    unet(sample, 961, text, pose_embedding_features=feats)
    assert unet.down_blocks[0].attentions[0]._plan is plan0 and engine.generation(unet) == gen0  # nothing changed
    with torch.no_grad():
        attn.to_q.weight.mul_(1.5)                                   # optimizer.step()-style in-place update
    unet(sample, 961, text, pose_embedding_features=feats)
    assert unet.down_blocks[0].attentions[0]._plan is not plan0 and engine.generation(unet) == gen0 + 1
    plan1 = unet.down_blocks[0].attentions[0]._plan
    attn.load_state_dict(attn.state_dict())                          # submodule load_state_dict
    unet(sample, 961, text, pose_embedding_features=feats)
    assert unet.down_blocks[0].attentions[0]._plan is not plan1
    plan2 = unet.down_blocks[0].attentions[0]._plan
    attn.processor.lora_scale = 0.5                                  # processor scalar edit
    unet(sample, 961, text, pose_embedding_features=feats)
    assert unet.down_blocks[0].attentions[0]._plan is not plan2


def test_spatial_pose_processors_are_refused(recorder):
    """add_spatial=True would install pose-adaptor processors the spatial kernels ignore -- refuse."""
    o_unet = helpers.build_oracle_unet(tiny=True)
    unet = helpers.build_product_unet(o_unet, tiny=True, device="cpu")
    kw = dict(helpers.ATTN_PROC_KWARGS, add_spatial=True)
    unet.set_all_attn_processor(add_spatial_lora=True, add_motion_lora=False, lora_kwargs={"lora_rank": 2, "lora_scale": 1.0},
                                motion_lora_kwargs={"lora_rank": -1, "lora_scale": 1.0},
                                pose_feature_dimensions=[320, 640], **kw)
    unet.requires_grad_(False)
    sample, text, feats, _ = _inputs()
    with pytest.raises(NotImplementedError, match="spatial"):
        unet(sample, 961, text, pose_embedding_features=feats)


def test_direct_module_calls_under_autograd_fail_loudly():
    """a mirror module called DIRECTLY under autograd with something requiring grad must raise a clear
    error instead of returning a tensor without grad_fn.  (The trainers' entry points -- PoseAdaptor / CamObjPoseAdaptor /
    get_traj_features_v2 -- do train: test_training_tape_plumbing, tests/test_gpu_training.py.)"""
    from synfmc_b200.fmc.models.pose_adaptor import CameraPoseEncoder, PoseAdaptor
    o_unet = helpers.build_oracle_unet(tiny=True)
    unet = helpers.build_product_unet(o_unet, tiny=True, device="cpu")
    enc = CameraPoseEncoder(channels=[320, 640], **helpers.POSE_ENCODER_KWARGS)
    sample, text, feats, _ = _inputs()
    enc.requires_grad_(True)
    with pytest.raises(RuntimeError, match="forward pass only"):
        enc(torch.zeros(2, 6, 4, 64, 64))
    for n, p in unet.named_parameters():
        if "merge" in n:
            p.requires_grad_(True)
    with pytest.raises(RuntimeError, match="forward pass only"):
        unet(sample, 961, text, pose_embedding_features=feats)
    unet.requires_grad_(False)
    with pytest.raises(RuntimeError, match="forward pass only"):
        unet(sample.requires_grad_(True), 961, text, pose_embedding_features=feats)
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA tensors only"):  # no_grad: passes the guard, hits the CPU refusal
        unet(sample, 961, text, pose_embedding_features=feats)
    # the wrapper under autograd takes the training path -- which, like everything else, refuses CPU tensors
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        PoseAdaptor(unet, enc)(sample.detach(), torch.tensor([961]), text, torch.zeros(2, 6, 4, 64, 64))


@pytest.mark.parametrize("form", ["sliced", "list"])
@pytest.mark.parametrize("obj", [False, True])
def test_windowed_pipeline_call_sequence(recorder, form, obj):
    """Multidiff windows (pipeline_animation.py:669-702): n_win U-Net passes per step + ONE window-average/DDIM kernel;
    per-window pose lists are encoded window by window; the obj pipeline keeps the reference's single-window assert
    unless windowed_objects=True."""
    from synfmc_b200.fmc._blocks import DDIMScheduler
    from synfmc_b200.fmc.pipelines.pipeline_animation import CameraCtrlPipeline
    from synfmc_b200.fmc.pipelines.pipeline_animation_cm_om import CameraObjCtrlPipeline
    channels = (320, 640)
    o_unet = helpers.build_oracle_unet(tiny=True, obj=obj)
    unet = helpers.build_product_unet(o_unet, tiny=True, obj=obj, device="cpu")
    enc = helpers.build_product_pose_encoder(helpers.build_oracle_pose_encoder(channels), channels, device="cpu")
    L, ov, n_win, H, W = 4, 2, 3, 64, 64
    F_total = n_win * (L - ov) + ov
    pose = torch.randn(1, 6, F_total, H, W)
    if form == "list":
        pose = [pose[:, :, k * (L - ov):k * (L - ov) + L].contiguous() for k in range(n_win)]
    pipe = (CameraObjCtrlPipeline if obj else CameraCtrlPipeline)(None, None, None, unet, DDIMScheduler(), enc)
    pipe.use_cuda_graph = False
    kw = dict(height=H, width=W, num_inference_steps=5, guidance_scale=7.5, latents=torch.randn(1, 4, F_total, 8, 8),
              prompt_embeds=torch.randn(2, 77, 768), multidiff_total_steps=n_win, multidiff_overlaps=ov, max_steps=2)
    if obj:
        kw["traj_features"] = [torch.randn(1, C, F_total, 8 >> l, 8 >> l) for l, C in enumerate(channels)]
        with pytest.raises(AssertionError):          # pipeline_animation_cm_om.py:690
            pipe(None, pose, L, **kw)
        kw["windowed_objects"] = True
    recorder.calls.clear()
    encoded, encode_cl = [], enc.encode_cl
    enc.encode_cl = lambda x: (encoded.append(x.shape[1]), encode_cl(x))[1]   # frames per CameraEncoder call
    out = pipe(None, pose, L, **kw)
    assert out.latents.shape == (1, 4, F_total, 8, 8)
    names = recorder.names()
    assert names.count("fmc_window_combine_ddim_f32") == 2 and "fmc_cfg_ddim_step_f32" not in names
    combine = next(a for n, a in recorder.calls if n == "fmc_window_combine_ddim_f32")
    assert combine[1] == n_win and combine[2] == 1 and combine[8] == F_total and combine[10] == L and combine[11] == L - ov
    conv_in = [a for n, a in recorder.calls if n == "fmc_conv3x3_bf16" and a[8] == 64 and a[9] == 320]
    assert len(conv_in) == 2 * n_win                # one U-Net pass per window per step
    assert encoded == ([L] * n_win if form == "list" else [F_total])


def test_circle_driven_object_path_call_sequence(recorder):
    """object features straight from the objects' circles -- the fused sphere-mask scatter in bf16 mode
    (no mask tensor), sphere masks -> bit-exact fp32 scatter in the reference-precision mode."""
    from synfmc_b200 import synth
    from synfmc_b200.fmc.util import traj_features_from_circles
    channels = (320, 640, 1280, 1280)
    omcm = helpers.build_product_omcm(helpers.build_oracle_omcm(channels), channels, device="cpu")
    info, circles = synth.synth_circles(1, 2, 64, 64, 3, seed=1)
    for mode, must, never in (("bf16", "fmc_traj_scatter_circles_unshuffle_bf16", "fmc_sphere_mask_f32"),
                              ("reference", "fmc_sphere_mask_f32", "fmc_traj_scatter_circles_unshuffle_bf16")):
        recorder.calls.clear()
        with engine.precision(mode):
            feats = traj_features_from_circles(info, circles, omcm, 64, 64)
        assert [f.dims for f in feats] == [(1, 2, 8, 8, 320), (1, 2, 4, 4, 640), (1, 2, 2, 2, 1280), (1, 2, 1, 1, 1280)]
        names = recorder.names()
        assert must in names and never not in names
        assert not any(n in ("fmc_traj_scatter_unshuffle_bf16",) for n in names)  # no mask tensor is ever read in bf16 mode


@pytest.mark.parametrize("stage", ["cmc", "omc"])
def test_training_tape_plumbing(recorder, stage):
    """Training forward + backward through the trainers' entry points (kernels replaced by the recorder): every trainable
    parameter of the stage receives a gradient of its own shape through torch.autograd, the frozen U-Net none; the forward
    issues no fused GEGLU / LN-folded / fused temporal kernel (their intermediates are needed by backward)."""
    from synfmc_b200 import synth
    from synfmc_b200.fmc.models.pose_adaptor import PoseAdaptor
    from synfmc_b200.fmc.models.pose_obj_adaptor import CamObjPoseAdaptor
    from synfmc_b200.fmc.util import get_traj_features_v2
    channels = (320, 640)
    obj = stage == "omc"
    o_unet = helpers.build_oracle_unet(tiny=True, obj=obj)
    unet = helpers.build_product_unet(o_unet, tiny=True, obj=obj, device="cpu")
    enc = helpers.build_product_pose_encoder(helpers.build_oracle_pose_encoder(channels), channels, device="cpu")
    omcm = helpers.build_product_omcm(helpers.build_oracle_omcm(channels), channels, device="cpu") if obj else None
    if stage == "cmc":   # train_cam_ctrl.py:259-284: the pose encoder and the qkv_merge layers
        enc.requires_grad_(True)
        for n, p in unet.named_parameters():
            if "merge" in n and "lora" not in n:
                p.requires_grad_(True)
    else:                # train_cam_obj_ctrl.py:386-391: the ObjectEncoder only
        omcm.requires_grad_(True)
    b, f, H, W = 1, 4, 64, 64
    latents = torch.randn(b, 4, f, 8, 8)
    text = torch.randn(b, 77, 768)
    pose = torch.randn(b, 6, f, H, W)
    t = torch.tensor([961])
    if obj:
        infos, masks = synth.synth_objects(b, f, H, W, 2, seed=1)
        trajs = get_traj_features_v2(infos, masks, omcm, False, 0.0, None, "cpu", torch.float32)
        assert [tuple(x.shape) for x in trajs] == [(1, 320, 4, 8, 8), (1, 640, 4, 4, 4)] and all(x.requires_grad for x in trajs)
        out = CamObjPoseAdaptor(unet, enc)(latents, t, text, pose, trajs)
    else:
        out = PoseAdaptor(unet, enc)(latents, t, text, pose)
    assert out.shape == latents.shape and out.requires_grad
    names = set(recorder.names())
    assert "fmc_gemm_ln_bf16" not in names and "fmc_temporal_qkv_attn_bf16" not in names and "fmc_geglu_fwd_bf16" in names
    assert not any(a[15] & 1 for n, a in recorder.calls if n == "fmc_gemm_bf16")   # no fused-GEGLU GEMM in training
    recorder.calls.clear()
    out.square().mean().backward()
    bnames = set(recorder.names())
    assert {"fmc_attention_bwd_bf16", "fmc_layernorm_bwd_bf16", "fmc_groupnorm_bwd_bf16", "fmc_geglu_bwd_bf16",
            "fmc_wgrad_bf16"} <= bnames   # weight gradients straight from the row-major operands: no transposed copies
    trainable = [(n, p) for m in (unet, enc) + ((omcm,) if obj else ()) for n, p in m.named_parameters() if p.requires_grad]
    assert trainable
    for n, p in trainable:
        # tiny U-Net: only its first down block has cross-attention, so only ObjectEncoder feature 0 is injected
        # (modified_modules.py:52-127) and the deeper ObjectEncoder levels are off the graph, as under torch autograd
        used = not obj or n.startswith(("conv_in", "zero_conv_in", "body.0.", "body.1.", "zero_conv_out_list.0."))
        if used:
            assert p.grad is not None and p.grad.shape == p.shape and p.grad.dtype == torch.float32, n
        else:
            assert p.grad is None, n
    assert all(p.grad is None for p in unet.parameters() if not p.requires_grad)


def test_edge_models_call_sequence(recorder):
    """SURVEY 8 f3 host logic: VAE decode / encode / decode_video and the CLIP text encoder issue the expected entry points
    with convertible arguments; the reference-precision mode is refused for both (bf16-mode components)."""
    from synfmc_b200.edge import AutoencoderKL, CLIPTextModel
    vae = AutoencoderKL(block_out_channels=(32, 64, 64, 64))
    out = vae.decode(torch.randn(2, 4, 4, 4)).sample
    assert out.shape == (2, 3, 32, 32) and out.dtype == torch.float32
    names = recorder.names()
    assert {"fmc_groupnorm_bf16", "fmc_gemm_bf16", "fmc_softmax_rows", "fmc_transpose_bf16", "fmc_resize_nearest_bf16",
            "fmc_cl_to_video_f32"} <= set(names)
    assert names.count("fmc_softmax_rows") == 2 and names.count("fmc_resize_nearest_bf16") == 3
    assert vae.decode_video(torch.randn(1, 4, 3, 4, 4), chunk=2).shape == (1, 3, 3, 32, 32)
    dist = vae.encode(torch.randn(1, 3, 32, 32)).latent_dist
    assert dist.sample(noise=torch.randn(1, 4, 4, 4), scale=0.18215).shape == (1, 4, 4, 4) and dist.mode().shape == (1, 4, 4, 4)
    assert "fmc_vae_sample_f32" in recorder.names()
    clip = CLIPTextModel(vocab_size=100, hidden_size=64, num_hidden_layers=2, num_attention_heads=2, intermediate_size=128)
    del recorder.calls[:]
    hidden = clip(torch.randint(0, 100, (2, 77)))[0]
    assert hidden.shape == (2, 77, 64) and hidden.dtype == torch.float32
    names = recorder.names()
    assert names[0] == "fmc_embed_tokens" and names.count("fmc_small_mha") == 2 and names.count("fmc_quick_gelu") == 2
    assert names.count("fmc_layernorm_bf16") == 5 and names.count("fmc_gemm_bf16") == 8
    mha = [a for n, a in recorder.calls if n == "fmc_small_mha"][0]
    assert mha[8:14] == (2, 77, 2, 32, 1.0, 1)   # sequences, tokens, heads, head_dim, scale (folded into q), causal
    with engine.precision("reference"):
        with pytest.raises(NotImplementedError):
            vae.decode(torch.randn(1, 4, 4, 4))
        with pytest.raises(NotImplementedError):
            clip(torch.randint(0, 100, (1, 77)))


def test_edge_models_refuse_cpu_tensors_without_the_dry_run():
    from synfmc_b200.edge import AutoencoderKL, CLIPTextModel
    with pytest.raises(Exception, match="CUDA"):
        AutoencoderKL(block_out_channels=(32, 32, 32, 32)).decode(torch.randn(1, 4, 4, 4))
    with pytest.raises(Exception, match="CUDA"):
        CLIPTextModel(vocab_size=10, hidden_size=64, num_hidden_layers=1, num_attention_heads=2, intermediate_size=64)(
            torch.zeros(1, 77, dtype=torch.long))


def test_linear_wgrad_routes(recorder):
    """bwd_ops.linear_wgrad: 16-byte-aligned widths go to fmc_wgrad_bf16 (row-major operands, no transposes); odd widths fall
    back to transposed copies + the K-major GEMM with fp32 output."""
    from synfmc_b200 import bwd_ops
    dy, x = torch.zeros(300, 328, dtype=torch.bfloat16), torch.zeros(300, 72, dtype=torch.bfloat16)
    dw = bwd_ops.linear_wgrad(dy, x)
    assert dw.shape == (328, 72) and dw.dtype == torch.float32 and recorder.names() == ["fmc_wgrad_bf16"]
    args = recorder.calls[0][1]
    assert args[7:11] == (300, 328, 72, 0)                                  # T, M, N, accumulate
    acc = torch.zeros(328, 72)
    bwd_ops.linear_wgrad(dy, x, out=acc, accumulate=True)
    assert recorder.calls[1][1][10] == 1
    del recorder.calls[:]
    dw = bwd_ops.linear_wgrad(torch.zeros(64, 20, dtype=torch.bfloat16), torch.zeros(64, 12, dtype=torch.bfloat16))
    assert dw.shape == (20, 12) and recorder.names() == ["fmc_transpose_bf16", "fmc_transpose_bf16", "fmc_gemm_bf16"]


def test_graphed_step_needs_a_gpu():
    from synfmc_b200.train import GraphedStep
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError, match="CUDA"):
        GraphedStep(lambda: None)


def test_edge_model_plans_follow_the_weights(recorder):
    """The VAE / text-encoder plans are rebuilt when a weight changes (load_state_dict after the first call), like the
    U-Net's (engine.fingerprint)."""
    from synfmc_b200.edge import AutoencoderKL, CLIPTextModel
    vae = AutoencoderKL(block_out_channels=(32, 32, 32, 32))
    vae.decode(torch.randn(1, 4, 4, 4))
    first = vae._plans
    vae.decode(torch.randn(1, 4, 4, 4))
    assert vae._plans is first                                   # unchanged weights: plans reused
    with torch.no_grad():
        vae.decoder.conv_out.weight.add_(1.0)
    vae.decode(torch.randn(1, 4, 4, 4))
    assert vae._plans is not first
    clip = CLIPTextModel(vocab_size=50, hidden_size=64, num_hidden_layers=1, num_attention_heads=2, intermediate_size=64)
    ids = torch.randint(0, 50, (1, 77))
    clip(ids)
    first = clip._plans
    clip.load_state_dict({k: v + 0.01 for k, v in clip.state_dict().items()})
    clip(ids)
    assert clip._plans is not first
