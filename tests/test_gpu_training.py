"""Training forward + backward on the tape (synfmc_b200/train_engine.py) against torch autograd of the fp32 CPU oracle
(= the reference's eager modules): gradients of the trainable subsets of both stages through the frozen U-Net.

  CMC (train_cam_ctrl.py:259-284, :586-648): CameraEncoder + the CameraAdapter qkv_merge layers
  OMC (train_cam_obj_ctrl.py:386-391, :843-862): the ObjectEncoder

Tolerances: the backward runs on bf16 activations / gradients with fp32 accumulation; one gradient tensor is held to 8e-2
rel-L2 (many are ~1e-2; the bound is for the small-norm ones at the far end of the chain), the whole trainable gradient
vector to 3e-2, the cosine between the two flat gradients to 0.999."""
import pytest
import torch

from oracle import harness as helpers
from oracle.harness import rel_l2

pytestmark = pytest.mark.gpu


def _compare(named_pairs, what):
    flat_g, flat_w = [], []
    worst = (0.0, None)
    for name, got, want in named_pairs:
        assert got is not None, f"{what}: no gradient for {name}"
        assert got.shape == want.shape, name
        e = rel_l2(got, want)
        if e > worst[0]:
            worst = (e, name)
        flat_g.append(got.detach().float().cpu().reshape(-1))
        flat_w.append(want.detach().float().reshape(-1))
    g, w = torch.cat(flat_g), torch.cat(flat_w)
    total = float((g - w).norm() / w.norm())
    cos = float(torch.dot(g, w) / (g.norm() * w.norm()))
    print(f"[parity] {what}: {len(flat_g)} gradient tensors, whole-vector rel-L2 {total:.3e}, cosine {cos:.6f}, "
          f"worst tensor {worst[1]} {worst[0]:.3e}")
    assert total < 3e-2 and cos > 0.999 and worst[0] < 8e-2, (total, cos, worst)


def test_motion_module_backward(cuda_device):
    """CameraAdapter motion module (attention_processor.py:255-293, motion_module.py:349-373): gradients w.r.t. the input,
    the pose feature and qkv_merge, at the three widths."""
    from oracle.attention_processor import AttnProcessor as OA, PoseAdaptorAttnProcessor as OP
    from oracle.motion_module import get_motion_module as o_get
    from oracle.unet import FMC_UNET_ADDITIONAL_KWARGS as KW
    from synfmc_b200 import ops, train_engine as te
    from synfmc_b200.fmc.models.attention_processor import AttnProcessor, PoseAdaptorAttnProcessor
    from synfmc_b200.fmc.models.motion_module import get_motion_module
    from synfmc_b200.synth import round_bf16, synth_init_
    dev = cuda_device
    for C, (b, f, h, w) in ((320, (2, 16, 6, 10)), (640, (1, 16, 5, 4)), (1280, (1, 8, 3, 3))):
        om = o_get(C, "Vanilla", dict(KW["motion_module_kwargs"]))
        pm = get_motion_module(C, "Vanilla", dict(KW["motion_module_kwargs"]))
        kw = dict(hidden_size=C, pose_feature_dim=C, query_condition=True, key_value_condition=True, scale=1.0)
        bo = om.temporal_transformer.transformer_blocks[0].attention_blocks
        bp = pm.temporal_transformer.transformer_blocks[0].attention_blocks
        bo[0].set_processor(OP(**kw)); bo[1].set_processor(OA())
        bp[0].set_processor(PoseAdaptorAttnProcessor(**kw)); bp[1].set_processor(AttnProcessor())
        synth_init_(om, seed=C)
        pm.load_state_dict(om.state_dict(), strict=True)
        pm.to(dev).requires_grad_(False)
        om.requires_grad_(False)
        for m in (om, pm):
            for n, p in m.named_parameters():
                if "merge" in n:
                    p.requires_grad_(True)
        g = torch.Generator().manual_seed(C)
        x = round_bf16(torch.randn(b, C, f, h, w, generator=g)).requires_grad_(True)
        pose = round_bf16(torch.randn(b, C, f, h, w, generator=g)).requires_grad_(True)
        dy = round_bf16(torch.randn(b, C, f, h, w, generator=g))
        om(x, None, None, None, cross_attention_kwargs={"pose_feature": pose}).backward(dy)
        tape = te.Tape()
        xv = te.Var(ops.to_channels_last(x.detach().to(dev)).view(-1, C))
        pv = te.Var(ops.to_channels_last(pose.detach().to(dev)).view(-1, C))
        with torch.no_grad():
            out = te.motion_module(tape, pm, xv, (b, f, h, w, C), pv, dev)
            out.grad = ops.to_channels_last(dy.to(dev)).view(-1, C)
            tape.backward()
        dx = ops.from_channels_last(xv.grad.view(b, f, h, w, C))
        dpose = ops.from_channels_last(pv.grad.view(b, f, h, w, C))
        # judge the gradient of the BRANCH (the module is input + branch: dx = dy + branch gradient)
        pairs = [("d input (branch)", dx.cpu() - dy, x.grad - dy), ("d pose", dpose, pose.grad)]
        o_params = dict(om.named_parameters())
        for n, p in pm.named_parameters():
            if p.requires_grad:
                pairs.append((n, tape.param_grads.get(p), o_params[n].grad))
        _compare(pairs, f"motion module backward C={C}")


def _cmc_trainable(unet, enc):
    unet.requires_grad_(False)
    enc.requires_grad_(True)
    for n, p in unet.named_parameters():
        if "merge" in n and "lora" not in n:
            p.requires_grad_(True)


@pytest.mark.timeout(900)
def test_cmc_training_step_gradients(cuda_device):
    """PoseAdaptor.forward + loss.backward() as train_cam_ctrl.py:586-648 runs them: tiny U-Net (every block type) with
    LoRA + CameraAdapter processors, CameraEncoder on Pluecker rays; the gradients of all 150 + 2 trainable tensors."""
    from oracle.pose_adaptor import PoseAdaptor as OPA
    from oracle.rays import to_plucker_embedding
    from synfmc_b200 import synth
    from synfmc_b200.fmc.models.pose_adaptor import PoseAdaptor
    channels = (320, 640)
    o_unet = helpers.build_oracle_unet(tiny=True)
    p_unet = helpers.build_product_unet(o_unet, tiny=True, device=cuda_device)
    o_enc = helpers.build_oracle_pose_encoder(channels)
    p_enc = helpers.build_product_pose_encoder(o_enc, channels, device=cuda_device)
    _cmc_trainable(o_unet, o_enc)
    _cmc_trainable(p_unet, p_enc)
    b, f, H, W = 2, 8, 64, 96
    K, c2w = synth.synth_camera(b, f, H, W, seed=6)
    plucker = to_plucker_embedding(c2w, K, (H, W)).permute(0, 2, 1, 3, 4).contiguous()
    latents, text = synth.synth_step_inputs(b, f, H // 8, W // 8, cfg=False, seed=6)
    g = torch.Generator().manual_seed(9)
    target = torch.randn(latents.shape, generator=g)
    t = torch.tensor([961, 41])
    loss_o = torch.nn.functional.mse_loss(OPA(o_unet, o_enc)(latents, t, text, plucker).float(), target)
    loss_o.backward()
    dev = cuda_device
    pred = PoseAdaptor(p_unet, p_enc)(latents.to(dev), t.to(dev), text.to(dev), plucker.to(dev))
    assert pred.requires_grad
    loss_p = torch.nn.functional.mse_loss(pred.float(), target.to(dev))
    loss_p.backward()
    assert abs(float(loss_p) - float(loss_o)) < 2e-2 * abs(float(loss_o))
    pairs = []
    for (n, po), (n2, pp) in zip(list(o_unet.named_parameters()) + list(o_enc.named_parameters()),
                                 list(p_unet.named_parameters()) + list(p_enc.named_parameters())):
        assert n == n2
        if po.requires_grad:
            pairs.append((n, pp.grad, po.grad))
        else:
            assert pp.grad is None, n
    _compare(pairs, "CMC training step (tiny U-Net + CameraEncoder)")


@pytest.mark.timeout(1800)
def test_config3_full_depth_gradients(cuda_device):
    """BASELINE config 3's step on the FULL 4-level U-Net (all 16 Transformer2D + 20 motion modules, mid block) with the
    full CameraEncoder, one clip at config 1's size (256x256x16f: the fp32 oracle's autograd keeps every attention matrix,
    0.5 GB each at this size): all 152 trainable tensors (218 M parameters)."""
    from oracle.pose_adaptor import PoseAdaptor as OPA
    from oracle.rays import to_plucker_embedding
    from synfmc_b200 import synth
    from synfmc_b200.fmc.models.pose_adaptor import PoseAdaptor
    channels = (320, 640, 1280, 1280)
    o_unet = helpers.build_oracle_unet(tiny=False)
    p_unet = helpers.build_product_unet(o_unet, tiny=False, device=cuda_device)
    o_enc = helpers.build_oracle_pose_encoder(channels)
    p_enc = helpers.build_product_pose_encoder(o_enc, channels, device=cuda_device)
    _cmc_trainable(o_unet, o_enc)
    _cmc_trainable(p_unet, p_enc)
    b, f, H, W = 1, 16, 256, 256
    K, c2w = synth.synth_camera(b, f, H, W, seed=16)
    plucker = to_plucker_embedding(c2w, K, (H, W)).permute(0, 2, 1, 3, 4).contiguous()
    latents, text = synth.synth_step_inputs(b, f, H // 8, W // 8, cfg=False, seed=16)
    target = torch.randn(latents.shape, generator=torch.Generator().manual_seed(17))
    t = torch.tensor([501])
    torch.nn.functional.mse_loss(OPA(o_unet, o_enc)(latents, t, text, plucker).float(), target).backward()
    dev = cuda_device
    pred = PoseAdaptor(p_unet, p_enc)(latents.to(dev), t.to(dev), text.to(dev), plucker.to(dev))
    torch.nn.functional.mse_loss(pred.float(), target.to(dev)).backward()
    pairs = []
    for (n, po), (n2, pp) in zip(list(o_unet.named_parameters()) + list(o_enc.named_parameters()),
                                 list(p_unet.named_parameters()) + list(p_enc.named_parameters())):
        if po.requires_grad:
            pairs.append((n, pp.grad, po.grad))
    assert len(pairs) >= 150
    _compare(pairs, "config 3 gradients, full-depth U-Net + CameraEncoder at 256x256x16f")


@pytest.mark.timeout(900)
def test_omc_training_step_gradients(cuda_device):
    """train_cam_obj_ctrl.py:843-862: get_traj_features_v2 (3 overlapping Gaussian objects) -> ObjectEncoder (trainable) ->
    CamObjPoseAdaptor -> loss.backward(); the gradients reach the ObjectEncoder through the frozen U-Net."""
    from oracle.pose_adaptor import CamObjPoseAdaptor as OCA
    from oracle.rays import to_plucker_embedding
    from oracle.util import get_traj_features_v2 as o_get
    from synfmc_b200 import synth
    from synfmc_b200.fmc.models.pose_obj_adaptor import CamObjPoseAdaptor
    from synfmc_b200.fmc.util import get_traj_features_v2
    channels = (320, 640)
    o_unet = helpers.build_oracle_unet(tiny=True, obj=True)
    p_unet = helpers.build_product_unet(o_unet, tiny=True, obj=True, device=cuda_device)
    o_enc = helpers.build_oracle_pose_encoder(channels)
    p_enc = helpers.build_product_pose_encoder(o_enc, channels, device=cuda_device)
    o_m = helpers.build_oracle_omcm(channels)
    p_m = helpers.build_product_omcm(o_m, channels, device=cuda_device)
    for m in (o_unet, o_enc, p_unet, p_enc):
        m.requires_grad_(False)
    o_m.requires_grad_(True)
    p_m.requires_grad_(True)
    b, f, H, W = 1, 8, 64, 96
    K, c2w = synth.synth_camera(b, f, H, W, seed=7)
    plucker = to_plucker_embedding(c2w, K, (H, W)).permute(0, 2, 1, 3, 4).contiguous()
    infos, masks = synth.synth_objects(b, f, H, W, 3, seed=8, gaussian=True)
    latents, text = synth.synth_step_inputs(b, f, H // 8, W // 8, cfg=False, seed=7)
    g = torch.Generator().manual_seed(10)
    target = torch.randn(latents.shape, generator=g)
    t = torch.tensor([801])
    o_traj = o_get(infos, masks, o_m, False, 0.0, None, "cpu", torch.float32)
    torch.nn.functional.mse_loss(OCA(o_unet, o_enc)(latents, t, text, plucker, o_traj).float(), target).backward()
    dev = cuda_device
    p_traj = get_traj_features_v2(infos, masks, p_m, False, 0.0, None, dev, torch.float32)
    pred = CamObjPoseAdaptor(p_unet, p_enc)(latents.to(dev), t.to(dev), text.to(dev), plucker.to(dev), p_traj)
    torch.nn.functional.mse_loss(pred.float(), target.to(dev)).backward()
    pairs = []
    for (n, po), (n2, pp) in zip(o_m.named_parameters(), p_m.named_parameters()):
        if po.grad is None or float(po.grad.abs().max()) == 0.0:
            assert pp.grad is None or float(pp.grad.abs().max()) == 0.0, n   # levels whose feature is never injected
            continue
        pairs.append((n, pp.grad, po.grad))
    assert len(pairs) >= 10
    _compare(pairs, "OMC training step (tiny U-Net + ObjectEncoder)")


@pytest.mark.timeout(900)
def test_graphed_training_step_matches_eager(cuda_device):
    """train.GraphedStep: the whole CMC step (zero_grad, tape forward, loss, backward, fused AdamW with the step count on
    the device) captured once and replayed gives the parameters eager stepping gives, step for step."""
    import copy
    from oracle.rays import to_plucker_embedding
    from synfmc_b200 import synth
    from synfmc_b200.fmc.models.pose_adaptor import PoseAdaptor
    from synfmc_b200.train import FlatParams, FusedAdamW, GraphedStep
    channels = (320, 640)
    dev = cuda_device
    o_unet = helpers.build_oracle_unet(tiny=True)
    o_enc = helpers.build_oracle_pose_encoder(channels)
    b, f, H, W = 1, 8, 64, 96
    K, c2w = synth.synth_camera(b, f, H, W, seed=6)
    plucker = to_plucker_embedding(c2w, K, (H, W)).permute(0, 2, 1, 3, 4).contiguous().to(dev)
    latents, text = synth.synth_step_inputs(b, f, H // 8, W // 8, cfg=False, seed=6)
    latents, text = latents.to(dev), text.to(dev)
    target = torch.randn(latents.shape, generator=torch.Generator().manual_seed(9)).to(dev)
    t = torch.tensor([441], device=dev)
    results = {}
    for mode in ("eager", "graph"):
        unet = helpers.build_product_unet(o_unet, tiny=True, device=dev)
        enc = helpers.build_product_pose_encoder(o_enc, channels, device=dev)
        _cmc_trainable(unet, enc)
        wrapper = PoseAdaptor(unet, enc)
        flat = FlatParams(list(enc.parameters()) + [p for p in unet.parameters() if p.requires_grad])
        opt = FusedAdamW(flat, lr=1e-3, max_grad_norm=1.0)
        start = flat.values.clone()

        def step():
            opt.zero_grad()
            loss = torch.nn.functional.mse_loss(wrapper(latents, t, text, plucker).float(), target)
            loss.backward()
            opt.step(device_state=(mode == "graph"))
            return loss.detach()
        losses = []
        if mode == "eager":
            for _ in range(4):
                losses.append(float(step()))
        else:
            graphed = GraphedStep(step, warmup=2)
            for _ in range(2):
                losses.append(float(graphed()))
        assert opt.steps_taken() == 4 and not opt.found_inf()
        results[mode] = (flat.values - start, losses)
    d_e, l_e = results["eager"]
    d_g, l_g = results["graph"]
    # Adam turns every gradient into a step of about lr, whatever its size: bf16-level differences between two runs (cuDNN
    # picks its convolution algorithms per run) move individual small-gradient weights by up to 2 lr, so the two
    # 4-step parameter displacements are compared as vectors, not element by element
    cos = float(torch.dot(d_e, d_g) / (d_e.norm() * d_g.norm()))
    ratio = float(d_g.norm() / d_e.norm())
    print(f"[parity] graphed vs eager training, 4 steps: losses {l_e} / {l_g}, displacement cosine {cos:.4f}, "
          f"norm ratio {ratio:.4f}")
    assert l_e[0] > l_e[-1]                                   # it trains
    assert abs(l_e[-1] - l_g[-1]) < 3e-2 * abs(l_e[-1])       # the 4th step of both runs
    assert cos > 0.9 and 0.9 < ratio < 1.1


@pytest.mark.timeout(900)
def test_early_gradient_delivery_matches_end_of_backward_delivery(cuda_device):
    """GradAllReduce.install_hooks: from the second step on the tape hands every parameter gradient to the reducer the
    moment it is complete (so a bucket's all-reduce overlaps the rest of backward) instead of returning all of them through
    its single autograd node at the end -- same gradients either way, every trainable parameter delivered early."""
    from oracle.rays import to_plucker_embedding
    from synfmc_b200 import synth, train_engine
    from synfmc_b200.fmc.models.pose_adaptor import PoseAdaptor
    from synfmc_b200.train import FlatParams, GradAllReduce
    channels = (320, 640)
    dev = cuda_device
    unet = helpers.build_product_unet(helpers.build_oracle_unet(tiny=True), tiny=True, device=dev)
    enc = helpers.build_product_pose_encoder(helpers.build_oracle_pose_encoder(channels), channels, device=dev)
    _cmc_trainable(unet, enc)
    wrapper = PoseAdaptor(unet, enc)
    flat = FlatParams(list(enc.parameters()) + [p for p in unet.parameters() if p.requires_grad])
    b, f, H, W = 1, 8, 64, 96
    K, c2w = synth.synth_camera(b, f, H, W, seed=6)
    plucker = to_plucker_embedding(c2w, K, (H, W)).permute(0, 2, 1, 3, 4).contiguous().to(dev)
    latents, text = synth.synth_step_inputs(b, f, H // 8, W // 8, cfg=False, seed=6)
    latents, text = latents.to(dev), text.to(dev)
    target = torch.randn(latents.shape, generator=torch.Generator().manual_seed(9)).to(dev)
    t = torch.tensor([441], device=dev)
    red = GradAllReduce(flat, bucket_bytes=8 << 20).install_hooks()
    launched, early = [], []
    red._launch = lambda bkt: launched.append(bkt)       # world size 1: record the bucket order instead of a collective
    sink = train_engine.EARLY_GRAD_SINK
    train_engine.EARLY_GRAD_SINK = lambda p, g: early.append(id(p)) or sink(p, g)
    try:
        grads = []
        for step in range(2):
            flat.zero_grad()
            red.reset()
            del launched[:], early[:]
            loss = torch.nn.functional.mse_loss(wrapper(latents, t, text, plucker).float(), target)
            loss.backward()
            assert sorted(launched) == list(range(len(red.buckets)))    # every bucket exactly once
            assert len(early) == (0 if step == 0 else len(flat.params))
            grads.append(flat.grads.clone())
        assert float(grads[0].norm()) > 0
        assert rel_l2(grads[1], grads[0]) < 1e-5
        assert launched[0] != len(red.buckets) - 1                       # buckets go out DURING backward, not in flat order
    finally:
        red.remove_hooks()
    assert train_engine.EARLY_GRAD_SINK is None
