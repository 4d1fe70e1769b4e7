"""CPU checks of the host-side weight preparation that feeds the CUDA kernels (no kernel is launched here): the layouts
below are contracts between synfmc_b200/engine.py and the kernels of include/fmc_b200.h."""
import torch
import torch.nn.functional as Fn
from torch import nn

from synfmc_b200 import engine, ops


def _bf(x):
    return x.to(torch.bfloat16).float()


def test_fused_temporal_weight_layout():
    """fmc_temporal_qkv_attn_bf16 wants, per head h, the rows [to_q(h) 40 | to_k(h) 40 | to_v(h) 40 | 8 zero rows]."""
    from synfmc_b200.fmc.models.attention_processor import AttnProcessor
    from synfmc_b200.fmc.models.motion_module import TemporalSelfAttention
    torch.manual_seed(0)
    attn = TemporalSelfAttention(attention_mode="Temporal_Self", cross_attention_dim=None, query_dim=320, heads=8,
                                 dim_head=40, temporal_position_encoding=True)
    attn.set_processor(AttnProcessor())
    plan = engine.AttnPlan(attn, torch.device("cpu"), fused_temporal=True)
    w = plan.w_head_major.float()
    assert w.shape == (8 * 128, 320)
    for h in (0, 3, 7):
        blk = w[h * 128:(h + 1) * 128]
        assert torch.equal(blk[0:40], _bf(attn.to_q.weight.detach()[h * 40:(h + 1) * 40]))
        assert torch.equal(blk[40:80], _bf(attn.to_k.weight.detach()[h * 40:(h + 1) * 40]))
        assert torch.equal(blk[80:120], _bf(attn.to_v.weight.detach()[h * 40:(h + 1) * 40]))
        assert torch.count_nonzero(blk[120:]) == 0
    # the un-fused layout (q | k heads padded 40 -> 48, then v) stays available for the other widths
    assert plan.qkv.w.shape == (8 * 48 * 2 + 320, 320)


def test_conv3x3_weight_layout_is_tap_major():
    """fmc_conv3x3_bf16 walks K as (ky, kx, cin): an explicit im2col product with ConvPlan.w2d must equal conv2d."""
    torch.manual_seed(1)
    conv = nn.Conv2d(64, 32, 3, padding=1)
    plan = engine.ConvPlan(conv, torch.device("cpu"))
    assert plan.fast3x3 and plan.w2d.shape == (32, 9 * 64)
    x = _bf(torch.randn(2, 64, 6, 8))
    want = Fn.conv2d(x, _bf(conv.weight.detach()), None, padding=1)
    xp = Fn.pad(x, (1, 1, 1, 1)).permute(0, 2, 3, 1)  # NHWC, zero border
    cols = torch.cat([xp[:, ky:ky + 6, kx:kx + 8, :] for ky in range(3) for kx in range(3)], dim=-1)  # [N, H, W, 9 * Cin]
    got = (cols.reshape(-1, 9 * 64) @ plan.w2d.float().t()).view(2, 6, 8, 32).permute(0, 3, 1, 2)
    assert torch.allclose(got, want, atol=1e-4, rtol=1e-4)
    assert torch.equal(plan.b32, conv.bias.detach())


def test_conv3x3_geometry_predicate():
    assert ops.conv3x3_supported(40, 64, 320, 320, 1)       # level 0 of BASELINE config 2
    assert ops.conv3x3_supported(40, 64, 320, 320, 2)       # Downsample2D
    assert ops.conv3x3_supported(5, 8, 1280, 1280, 1)       # level 3: tiles span images
    assert ops.conv3x3_supported(32, 32, 64, 320, 1)        # conv_in of config 1 (latent channels padded to 64)
    assert not ops.conv3x3_supported(40, 40, 320, 320, 1)   # 128 % 40 != 0 -> cuDNN path
    assert not ops.conv3x3_supported(40, 64, 8, 320, 1)     # Cin not a multiple of 64
    assert not ops.conv3x3_supported(41, 64, 320, 320, 2)   # stride does not divide the image


def test_geglu_and_lora_plans_are_device_agnostic():
    """LinearPlan interleaves GEGLU rows in blocks of 16 (value | gate) and keeps N, K; _fold_lora is exact in fp32."""
    w, b = torch.randn(64, 16), torch.randn(64)
    plan = engine.LinearPlan(w, b, torch.device("cpu"), geglu=True)
    assert (plan.N, plan.K) == (64, 16)
    val, gate = w[:32], w[32:]
    assert torch.equal(plan.w[:16].float(), _bf(val[:16])) and torch.equal(plan.w[16:32].float(), _bf(gate[:16]))
    assert torch.equal(plan.b[:16], b[:16]) and torch.equal(plan.b[16:32], b[32:48])


def test_layernorm_fold_algebra():
    """LinearPlan(pre_norm=...): rstd * (x W'^T - mean * colsum) + bias' == Linear(LayerNorm(x)) with W' = W * gamma,
    colsum = row sums of the bf16 W', bias' = bias + W beta."""
    torch.manual_seed(2)
    C, N = 64, 48
    norm = nn.LayerNorm(C)
    norm.weight.data = 1 + 0.2 * torch.randn(C)
    norm.bias.data = 0.3 * torch.randn(C)
    w, b = torch.randn(N, C) / 8, torch.randn(N)
    plan = engine.LinearPlan(w, b, torch.device("cpu"), pre_norm=norm)
    x = _bf(torch.randn(10, C) + 0.5)
    mean, rstd = x.mean(1, keepdim=True), (x.var(1, unbiased=False, keepdim=True) + norm.eps).rsqrt()
    got = rstd * (x @ plan.w.float().t() - mean * plan.colsum[None, :]) + plan.b[None, :]
    want = Fn.layer_norm(x, (C,), norm.weight.data, norm.bias.data, norm.eps) @ w.t() + b
    assert torch.allclose(got, want, atol=2e-2, rtol=2e-2)  # bf16 weights
    exact = rstd * (x @ (w * norm.weight.data).t() - mean * (w * norm.weight.data).sum(1)[None, :]) + (b + w @ norm.bias.data)
    assert torch.allclose(exact, want, atol=1e-4, rtol=1e-4)


def test_groupnorm_kernel_plan(monkeypatch):
    """fmc_groupnorm_launches is pure host logic (the shape plan of norm.cu): the single-pass cluster kernel takes every
    GroupNorm shape of the config-2 U-Net except C = 960 at 2560 tokens (120-channel chunks x 2560 rows do not fit a
    cluster of 8), FMC_GN_FUSED=0 forces the three-kernel form, mode 1 leaves the 120-channel chunks to it."""
    from synfmc_b200 import _cabi
    fn = _cabi.lib().fmc_groupnorm_launches
    unet_shapes = [(2560, 320), (2560, 640), (2560, 960), (640, 320), (640, 640), (640, 960), (640, 1280), (640, 1920),
                   (160, 640), (160, 1280), (160, 1920), (160, 2560), (40, 1280), (40, 2560)]
    monkeypatch.delenv("FMC_GN_FUSED", raising=False)
    assert {s: fn(s[0], s[1], 32) for s in unet_shapes} == {s: (3 if s == (2560, 960) else 1) for s in unet_shapes}
    monkeypatch.setenv("FMC_GN_FUSED", "0")
    assert all(fn(hw, c, 32) == 3 for hw, c in unet_shapes)
    monkeypatch.setenv("FMC_GN_FUSED", "1")
    assert fn(640, 960, 32) == 3 and fn(640, 1920, 32) == 3 and fn(640, 640, 32) == 1
    monkeypatch.delenv("FMC_GN_FUSED", raising=False)
    assert fn(64, 128, 32) == 3      # 4 channels per group: a 16-byte vector would span more than two groups
    assert fn(64, 320, 0) == 3 and fn(64, 321, 32) == 3   # nonsense arguments never fail, they just say "three-kernel"


def test_domain_lora_fold_and_the_no_lora_configuration():
    """engine.AttnPlan for the spatial attentions: with a LoRAAttnProcessor the projection the kernels see is
    W + s * up @ down (attention_processor.py:138-157: q, k, v, out; text cross-attention has k / v of width 768) with q / k
    heads padded 40 -> 48; with the plain processor (`add_spatial_lora=False`, what the trainers configure when no image
    LoRA checkpoint is given, train_cam_ctrl.py:230) it is W itself."""
    from synfmc_b200.fmc._blocks import Attention
    from synfmc_b200.fmc.models.attention_processor import AttnProcessor, LoRAAttnProcessor
    torch.manual_seed(3)
    for cross_dim in (None, 768):
        attn = Attention(query_dim=320, cross_attention_dim=cross_dim, heads=8, dim_head=40)
        proc = LoRAAttnProcessor(hidden_size=320, cross_attention_dim=cross_dim, rank=160, lora_scale=0.7)
        for lin in (proc.to_q_lora, proc.to_k_lora, proc.to_v_lora, proc.to_out_lora):
            nn.init.normal_(lin.up.weight, std=0.02)
        attn.set_processor(proc)
        plan = engine.AttnPlan(attn, torch.device("cpu"))

        def want(name):
            lin = attn.to_out[0] if name == "to_out" else getattr(attn, name)
            lora = getattr(proc, name + "_lora")
            return lin.weight.detach() + 0.7 * (lora.up.weight.detach() @ lora.down.weight.detach())
        q = want("to_q").view(8, 40, 320)
        got_q = (plan.q.w if cross_dim else plan.qkv.w)[:8 * 48].float().view(8, 48, 320)
        assert torch.equal(got_q[:, :40], _bf(q)) and torch.count_nonzero(got_q[:, 40:]) == 0
        kv = plan.kv.w.float() if cross_dim else plan.qkv.w.float()[8 * 48:]
        assert torch.equal(kv[:8 * 48].view(8, 48, -1)[:, :40], _bf(want("to_k")).view(8, 40, -1))
        assert torch.equal(kv[8 * 48:], _bf(want("to_v")))
        assert torch.equal(plan.out.w.float(), _bf(want("to_out")))
        attn.set_processor(AttnProcessor())
        plain = engine.AttnPlan(attn, torch.device("cpu"))
        assert torch.equal(plain.out.w.float(), _bf(attn.to_out[0].weight.detach()))
        got_q = (plain.q.w if cross_dim else plain.qkv.w)[:8 * 48].float().view(8, 48, 320)
        assert torch.equal(got_q[:, :40], _bf(attn.to_q.weight.detach()).view(8, 40, 320))
