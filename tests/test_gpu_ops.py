"""GPU parity of every C-ABI entry point (include/fmc_b200.h) against the oracle / a plain torch fp32 statement of
the same arithmetic, on identical bf16-exact inputs.

Tolerances (DESIGN.md "numerics"): a bf16 output carries a rounding floor of ~1.7e-3 rel-L2, so bf16-output
ops are held to 4e-3; ops with an fp32 output (FMC_GEMM_OUT_F32, rays, DDIM) are held to 1e-5 -- far inside the 1e-3
north-star bound; index / byte work (object-mask scatter) is bit-exact."""
import math

import pytest
import torch
import torch.nn.functional as Fn

pytestmark = pytest.mark.gpu

BF16_TOL = 4e-3
FUSED_TOL = 6e-3  # kernels that chain two tensor-core stages through bf16 operands (projection -> attention)
F32_TOL = 1e-5


def within_one_bf16_ulp(got, want):
    """Fraction of elements whose bf16 result is within ONE bf16 ulp of the exact (fp64) value -- SURVEY H1's per-kernel
    form of the 1e-3 bound for a bf16-output kernel that is fp32 inside: the only error left is the output rounding."""
    got, want = got.double().cpu(), want.double().cpu()
    ulp = torch.exp2(torch.floor(torch.log2(want.abs().clamp_min(2.0 ** -120))) - 7)
    return float(((got - want).abs() <= ulp).double().mean())


def bf(x):
    return x.to(torch.bfloat16)


def rel(got, want):
    got, want = got.float().cpu(), want.float().cpu()
    return float((got - want).norm() / want.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def ops(cuda_device):
    from synfmc_b200 import ops as o
    return o


def randn(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).float()


# ---------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 320, 320), (77, 640, 768), (2, 1280, 320), (5000, 960, 320),
                                   (4096, 320, 1280), (1000, 1280, 5120), (20480, 640, 640)])
def test_gemm_f32_out(ops, cuda_device, M, N, K):
    a, w = randn(M, K, seed=1), randn(N, K, seed=2, scale=K ** -0.5)
    bias = randn(N, seed=3)
    got = ops.gemm(bf(a).to(cuda_device), bf(w).to(cuda_device), bias=bias.to(cuda_device), out_f32=True)
    want = a.double() @ w.double().t() + bias.double()
    assert rel(got, want) < F32_TOL


@pytest.mark.parametrize("M,N,K", [(256, 320, 320), (1000, 1280, 320), (81920 // 8, 320, 1280)])
def test_gemm_residual_rowbias_bf16(ops, cuda_device, M, N, K):
    a, w = randn(M, K, seed=1), randn(N, K, seed=2, scale=K ** -0.5)
    bias, res = randn(N, seed=3), randn(M, N, seed=4)
    rpg = 100
    rowbias = randn((M + rpg - 1) // rpg, N, seed=5)
    got = ops.gemm(bf(a).to(cuda_device), bf(w).to(cuda_device), bias=bias.to(cuda_device),
                   residual=bf(res).to(cuda_device), rowbias=rowbias.to(cuda_device), rows_per_group=rpg)
    want = a @ w.t() + bias + res + rowbias.repeat_interleave(rpg, 0)[:M]
    assert got.dtype == torch.bfloat16
    assert rel(got, want) < BF16_TOL


@pytest.mark.parametrize("M,N,K,res", [(300, 1280, 1280, True), (1000, 1920, 640, False), (129, 1088, 1280, False),
                                       (5120, 1280, 5120, True), (257, 3840, 1280, False)])
def test_gemm_cta_pair(ops, cuda_device, M, N, K, res):
    """Shapes that take the CTA-pair MMA path (tcgen05 cta_group::2: K >= 1280, or K >= 640 with N > 640): odd numbers
    of 128-row tiles (the last pair's odd CTA recomputes the last tile), ragged M, N that is not a tile multiple."""
    a, w = randn(M, K, seed=1), randn(N, K, seed=2, scale=K ** -0.5)
    bias = randn(N, seed=3)
    r = randn(M, N, seed=4) if res else None
    got = ops.gemm(bf(a).to(cuda_device), bf(w).to(cuda_device), bias=bias.to(cuda_device),
                   residual=None if r is None else bf(r).to(cuda_device))
    want = a @ w.t() + bias + (r if res else 0)
    assert rel(got, want) < BF16_TOL


@pytest.mark.parametrize("M,C", [(1000, 320), (333, 1280)])
def test_gemm_geglu(ops, cuda_device, M, C):
    """diffusers FeedForward GEGLU (motion_module.py:297): proj -> chunk(2) -> value * gelu_erf(gate)."""
    from synfmc_b200.engine import LinearPlan
    x = randn(M, C, seed=1)
    w, b = randn(8 * C, C, seed=2, scale=C ** -0.5), randn(8 * C, seed=3)
    plan = LinearPlan(w, b, cuda_device, geglu=True)
    got = plan(bf(x).to(cuda_device))
    val, gate = (x @ w.t() + b).chunk(2, dim=-1)
    want = val * Fn.gelu(gate)
    assert got.shape == (M, 4 * C)
    assert rel(got, want) < BF16_TOL


@pytest.mark.parametrize("M,C,N,geglu", [(1000, 320, 1088, False), (333, 640, 384, False), (1000, 320, 2560, True),
                                         (5120, 1280, 10240, True)])
def test_gemm_with_folded_layernorm(ops, cuda_device, M, C, N, geglu):
    """LayerNorm folded into the consuming GEMM (fmc_rowstats_bf16 + fmc_gemm_ln_bf16) against LayerNorm -> Linear
    (-> GEGLU) in fp32 torch, on rows with a non-zero mean (the term the epilogue has to cancel)."""
    import torch.nn as nn
    from synfmc_b200.engine import LinearPlan
    x = randn(M, C, seed=1) * 1.5 + 0.7
    norm = nn.LayerNorm(C)
    norm.weight.data = 1 + 0.2 * randn(C, seed=2)
    norm.bias.data = 0.3 * randn(C, seed=3)
    w, b = randn(N, C, seed=4, scale=C ** -0.5), randn(N, seed=5)
    plan = LinearPlan(w, b, cuda_device, geglu=geglu, pre_norm=norm)
    xd = bf(x).to(cuda_device)
    stats = ops.rowstats(xd, norm.eps)
    xf = bf(x).float()
    assert rel(stats[:, 0], xf.mean(1)) < 1e-5
    assert rel(stats[:, 1], (xf.var(1, unbiased=False) + norm.eps).rsqrt()) < 1e-5
    got = plan(xd, ln_stats=stats)
    y = Fn.layer_norm(xf, (C,), norm.weight.data, norm.bias.data, norm.eps) @ w.t() + b
    if geglu:
        val, gate = y.chunk(2, dim=-1)
        y = val * Fn.gelu(gate)
    assert rel(got, y) < BF16_TOL


def test_gemm_rejects_bad_shapes(ops, cuda_device):
    from synfmc_b200._cabi import FmcError
    a = torch.zeros(16, 30, dtype=torch.bfloat16, device=cuda_device)  # K not a multiple of 8 -> unaligned row stride
    w = torch.zeros(32, 30, dtype=torch.bfloat16, device=cuda_device)
    with pytest.raises(FmcError):
        ops.gemm(a, w)


# ---------------------------------------------------------------- attention cores
def _pad_heads(x, heads, d, hs):
    rows = x.shape[0]
    out = torch.zeros(rows, heads, hs)
    out[:, :, :d] = x.view(rows, heads, d)
    return out.view(rows, heads * hs)


@pytest.mark.parametrize("images,d,nq,nk,kv_div", [(2, 40, 256, 256, 1), (2, 40, 300, 300, 1), (2, 80, 640, 640, 1),
                                                   (3, 160, 160, 160, 1), (2, 160, 40, 40, 1), (4, 40, 200, 77, 2),
                                                   (4, 160, 160, 77, 4), (2, 40, 2560, 2560, 1),
                                                   (32, 80, 640, 77, 16), (16, 160, 40, 77, 16), (32, 40, 2560, 77, 16),
                                                   (6, 40, 100, 77, 3), (2, 160, 160, 64, 1)])
def test_spatial_attention(ops, cuda_device, images, d, nq, nk, kv_div):
    """softmax(q k^T d^-1/2) v per head == attention_processor.py:148-154 (head_to_batch_dim, baddbmm, softmax, bmm)."""
    heads, hs = 8, (d + 15) // 16 * 16
    C = heads * d
    groups = images // kv_div
    kv_stride = (nk + 7) // 8 * 8
    q = randn(images * nq, C, seed=1)
    k = randn(groups * kv_stride, C, seed=2)
    v = randn(groups * kv_stride, C, seed=3)
    qp, kp = _pad_heads(q, heads, d, hs), _pad_heads(k, heads, d, hs)
    kv = torch.cat([kp, v], dim=1)
    out = torch.empty(images * nq, C, dtype=torch.bfloat16, device=cuda_device)
    qd, kvd = bf(qp).to(cuda_device), bf(kv).to(cuda_device)
    ops.spatial_attn(qd, 0, kvd, 0, kvd, heads * hs, hs, out, images, heads, d, nq, nk, kv_div, kv_stride, d ** -0.5)
    qh = q.view(images, nq, heads, d).transpose(1, 2)
    kh = k.view(groups, kv_stride, heads, d)[:, :nk].transpose(1, 2).repeat_interleave(kv_div, 0)
    vh = v.view(groups, kv_stride, heads, d)[:, :nk].transpose(1, 2).repeat_interleave(kv_div, 0)
    want = Fn.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(images * nq, C)
    assert rel(out, want) < BF16_TOL


@pytest.mark.parametrize("images,nq", [(2, 256), (2, 300), (1, 2560)])
def test_spatial_attention_fp16_v(ops, cuda_device, images, nq):
    """head_dim 40 self-attention with V projected to fp16 (FMC_GEMM_F16_TAIL) and fp16x2 exponentials
    (fmc_spatial_attn_vf16): the fused q|k|v GEMM writes the fp16 tail, the attention consumes it."""
    heads, d, hs = 8, 40, 48
    C = heads * d
    x = randn(images * nq, C, seed=1)
    wq, wk, wv = (randn(C, C, seed=s, scale=C ** -0.5) for s in (2, 3, 4))
    w = torch.cat([_pad_heads(wq.t(), heads, d, hs).t(), _pad_heads(wk.t(), heads, d, hs).t(), wv], dim=0).contiguous()
    v_col0 = 2 * heads * hs
    qkv = ops.gemm(bf(x).to(cuda_device), bf(w).to(cuda_device), f16_from_col=v_col0)
    # the tail really is fp16
    v_got = qkv[:, v_col0:].contiguous().view(torch.float16).float().cpu()
    assert rel(v_got, x @ wv.t()) < 2e-3
    out = torch.empty(images * nq, C, dtype=torch.bfloat16, device=cuda_device)
    ops.spatial_attn(qkv, 0, qkv, heads * hs, qkv, v_col0, hs, out, images, heads, d, nq, nq, 1, nq, d ** -0.5, v_f16=True)
    q, k, v = (bf(x @ m.t()).float() for m in (wq, wk, wv))

    def hd(t):
        return t.view(images, nq, heads, d).transpose(1, 2)
    want = Fn.scaled_dot_product_attention(hd(q), hd(k), hd(v)).transpose(1, 2).reshape(images * nq, C)
    assert rel(out, want) < BF16_TOL


@pytest.mark.parametrize("B,F,HW,d", [(1, 16, 64, 40), (2, 16, 100, 40), (2, 16, 50, 80), (2, 16, 21, 160),
                                      (1, 4, 70, 40), (1, 8, 33, 80), (1, 32, 10, 40), (2, 16, 2560, 40)])
def test_temporal_attention(ops, cuda_device, B, F, HW, d):
    """Attention over the f frames of each latent position on channels-last rows: motion_module.py:349-389 with the
    '(b h w) f c' rearranges of :218/:230 folded into addressing."""
    heads, hs = 8, (d + 15) // 16 * 16
    C = heads * d
    rows = B * F * HW
    q, k, v = randn(rows, C, seed=1), randn(rows, C, seed=2), randn(rows, C, seed=3)
    qkv = torch.cat([_pad_heads(q, heads, d, hs), _pad_heads(k, heads, d, hs), v], dim=1)
    out = torch.empty(rows, C, dtype=torch.bfloat16, device=cuda_device)
    ops.temporal_attn(bf(qkv).to(cuda_device), 0, heads * hs, 2 * heads * hs, hs, out, B, F, HW, heads, d, d ** -0.5)

    def seq(x):  # [(B F HW), C] -> [(B HW), heads, F, d]
        return x.view(B, F, HW, heads, d).permute(0, 2, 3, 1, 4).reshape(B * HW, heads, F, d)
    o = Fn.scaled_dot_product_attention(seq(q), seq(k), seq(v))
    want = o.view(B, HW, heads, F, d).permute(0, 3, 1, 2, 4).reshape(rows, C)
    assert rel(out, want) < BF16_TOL


@pytest.mark.parametrize("B,F,HW", [(1, 16, 8), (1, 16, 64), (2, 16, 100), (1, 4, 70), (1, 8, 33), (1, 32, 10),
                                    (2, 16, 2560)])
def test_temporal_qkv_attention_fused(ops, cuda_device, B, F, HW):
    """Fused q|k|v projection + attention over frames (C = 320, 8 x 40) against fp32 torch on the same bf16 inputs:
    to_q/to_k/to_v + attention core of attention_processor.py:46-67 / :259-281."""
    heads, d, hs, C = 8, 40, 48, 320
    rows = B * F * HW
    x = randn(rows, C, seed=1)
    wq, wk, wv = (randn(C, C, seed=s, scale=C ** -0.5) for s in (2, 3, 4))
    blocks = [w.view(heads, d, C) for w in (wq, wk, wv)] + [torch.zeros(heads, 128 - 3 * d, C)]
    w_head_major = torch.cat(blocks, dim=1).reshape(heads * 128, C)
    out = torch.empty(rows, C, dtype=torch.bfloat16, device=cuda_device)
    ops.temporal_qkv_attn(bf(x).to(cuda_device), bf(w_head_major).to(cuda_device), out, B, F, HW, heads, d ** -0.5)

    def seq(t):  # [(B F HW), C] -> [(B HW), heads, F, d]
        return t.view(B, F, HW, heads, d).permute(0, 2, 3, 1, 4).reshape(B * HW, heads, F, d)
    # fp32 statement of the reference lines, NO rounding mirrored from the kernel (which feeds q, k, v and the
    # probabilities to the tensor cores as bf16 MMA operands: three roundings more than a single bf16-output op)
    q, k, v = (x @ w.t() for w in (wq, wk, wv))
    o = Fn.scaled_dot_product_attention(seq(q), seq(k), seq(v))
    want = o.view(B, HW, heads, F, d).permute(0, 3, 1, 2, 4).reshape(rows, C)
    err = rel(out, want)
    print(f"[parity] fused temporal q|k|v + attention B={B} F={F} HW={HW}: rel-L2 vs fp32 = {err:.3e}")
    assert err < FUSED_TOL
    torch.cuda.synchronize()


@pytest.mark.parametrize("N,H,W,cin,cout,stride,extras", [
    (2, 16, 32, 64, 64, 1, False), (3, 10, 16, 128, 160, 1, True), (4, 5, 8, 64, 32, 1, True), (3, 8, 16, 64, 96, 1, False),
    (2, 16, 32, 64, 96, 2, True), (2, 20, 32, 192, 320, 2, False), (1, 40, 64, 320, 320, 1, True), (16, 4, 4, 64, 64, 1, False)])
def test_conv3x3_implicit_gemm(ops, cuda_device, N, H, W, cin, cout, stride, extras):
    """fmc_conv3x3_bf16 (implicit GEMM, 4-D TMA windows with zero-filled borders) against torch conv2d in fp32 on the
    same bf16 inputs: tiles that span several images (H = 10, 5, 4), odd tile counts, stride 2, fused bias + residual."""
    assert ops.conv3x3_supported(H, W, cin, cout, stride)
    x = randn(N, H, W, cin, seed=1)
    w = randn(cout, cin, 3, 3, seed=2, scale=(9 * cin) ** -0.5)
    bias = randn(cout, seed=3) if extras else None
    res = randn(N, H // stride, W // stride, cout, seed=4) if extras else None
    want = Fn.conv2d(x.permute(0, 3, 1, 2), w, bias, stride=stride, padding=1).permute(0, 2, 3, 1)
    if res is not None:
        want = want + res
    w2d = bf(w.permute(0, 2, 3, 1).reshape(cout, 9 * cin)).to(cuda_device).contiguous()
    got = ops.conv3x3(bf(x).to(cuda_device), w2d, bias=None if bias is None else bias.to(cuda_device),
                      residual=None if res is None else bf(res).to(cuda_device), stride=stride)
    assert got.shape == want.shape
    assert rel(got, want) < BF16_TOL


# ---------------------------------------------------------------- norms / elementwise
@pytest.mark.parametrize("rows,C", [(1000, 320), (77, 640), (4096, 1280)])
def test_layernorm_plain(ops, cuda_device, rows, C):
    x, g, b = randn(rows, C, seed=1), 1 + 0.1 * randn(C, seed=2), 0.1 * randn(C, seed=3)
    got = ops.layernorm(bf(x).to(cuda_device), g.to(cuda_device), b.to(cuda_device), 1e-5)
    assert rel(got, Fn.layer_norm(x, (C,), g, b, 1e-5)) < BF16_TOL


@pytest.mark.parametrize("rows,C", [(2000, 320), (500, 640), (300, 1280)])
def test_layernorm_is_exact_up_to_output_rounding(ops, cuda_device, rows, C):
    """LayerNorm (+PE, + pose add) is fp32 inside: every bf16 output is within one bf16 ulp of the fp64 result."""
    F, HW = 4, 5
    x, g, b = randn(rows, C, seed=1, scale=2.0), 1 + 0.1 * randn(C, seed=2), 0.1 * randn(C, seed=3)
    pe, add = randn(F, C, seed=4), randn(rows, C, seed=5)
    out, out2 = ops.layernorm(bf(x).to(cuda_device), g.to(cuda_device), b.to(cuda_device), 1e-5, pe=pe.to(cuda_device),
                              F=F, HW=HW, add=bf(add).to(cuda_device))
    frame = (torch.arange(rows) // HW) % F
    want = Fn.layer_norm(x.double(), (C,), g.double(), b.double(), 1e-5) + pe.double()[frame]
    assert within_one_bf16_ulp(out, want) > 0.9999
    # out2 = (LN + PE) + pose in fp32 from the un-rounded LN value, then one rounding
    assert within_one_bf16_ulp(out2, want + add.double()) > 0.9999


@pytest.mark.parametrize("images,HW,C,silu", [(4, 2560, 320, True), (4, 640, 640, False), (2, 160, 1280, True),
                                              (2, 40, 2560, True), (2, 2560, 960, True)])
def test_groupnorm_is_exact_up_to_output_rounding(ops, cuda_device, images, HW, C, silu):
    """GroupNorm (+SiLU) at the step's shapes: within one bf16 ulp of the fp64 result (fp32 statistics and apply)."""
    x = bf(randn(images * HW, C, seed=1, scale=2.0) + 0.25).float()  # bf16-exact: both sides see the same numbers
    g, b = 1 + 0.1 * randn(C, seed=2), 0.1 * randn(C, seed=3)
    got = ops.groupnorm(bf(x).to(cuda_device), g.to(cuda_device), b.to(cuda_device), 1e-5, images, HW, groups=32, silu=silu)
    want = Fn.group_norm(x.view(images, HW, C).double().permute(0, 2, 1), 32, g.double(), b.double(), 1e-5)
    if silu:
        want = Fn.silu(want)
    frac = within_one_bf16_ulp(got, want.permute(0, 2, 1).reshape(images * HW, C))
    print(f"[parity] groupnorm images={images} HW={HW} C={C} silu={silu}: {frac:.6f} of outputs within 1 bf16 ulp of fp64")
    assert frac > 0.999


@pytest.mark.parametrize("images,HW,C,silu,bias", [(32, 2560, 320, True, False), (32, 640, 640, True, True),
                                                   (32, 2560, 640, False, False), (16, 1000, 320, True, True),
                                                   (32, 640, 1920, True, False), (64, 640, 1280, False, False)])
def test_groupnorm_persistent_is_bit_identical(ops, cuda_device, monkeypatch, images, HW, C, silu, bias):
    """The persistent, double-buffered GroupNorm (FMC_GN_FUSED=5: resident clusters walking (image, chunk) items with the
    next slab always in flight) keeps the arithmetic and reduction order of the one-item-per-cluster kernel (mode 4):
    identical bits, at shapes large enough to take the persistent path (several items per resident cluster)."""
    x = bf(randn(images * HW, C, seed=1, scale=1.5) + 0.2).to(cuda_device)
    g, b = (1 + 0.1 * randn(C, seed=2)).to(cuda_device), (0.1 * randn(C, seed=3)).to(cuda_device)
    rb = randn(images // 2, C, seed=4).to(cuda_device) if bias else None
    outs = {}
    for mode in ("4", "5"):
        monkeypatch.setenv("FMC_GN_FUSED", mode)
        outs[mode] = ops.groupnorm(x, g, b, 1e-5, images, HW, groups=32, silu=silu, rowbias=rb, rowbias_div=2 if bias else 1)
    assert torch.equal(outs["4"], outs["5"])
    xi = (x.float().view(images, HW, C) + (rb.repeat_interleave(2, 0)[:, None, :] if bias else 0)).permute(0, 2, 1)
    want = Fn.group_norm(xi, 32, g, b, 1e-5)
    want = Fn.silu(want) if silu else want
    assert rel(outs["5"], want.permute(0, 2, 1).reshape(images * HW, C)) < BF16_TOL


def test_layernorm_pe_pose(ops, cuda_device):
    """LN -> + pe[frame] (motion_module.py:319-321) and the x + pose_feature operand of qkv_merge
    (attention_processor.py:257)."""
    B, F, HW, C = 2, 16, 37, 320
    rows = B * F * HW
    x, g, b = randn(rows, C, seed=1), 1 + 0.1 * randn(C, seed=2), 0.1 * randn(C, seed=3)
    pe, pose = randn(32, C, seed=4), randn(rows, C, seed=5)
    got, got2 = ops.layernorm(bf(x).to(cuda_device), g.to(cuda_device), b.to(cuda_device), 1e-5, pe=pe.to(cuda_device),
                              F=F, HW=HW, add=bf(pose).to(cuda_device))
    frame = (torch.arange(rows) // HW) % F
    want = Fn.layer_norm(x, (C,), g, b, 1e-5) + pe[frame]
    assert rel(got, want) < BF16_TOL
    assert rel(got2, want + pose) < BF16_TOL


@pytest.mark.parametrize("images,HW,C,silu", [(4, 256, 320, False), (3, 100, 640, True), (2, 40, 1280, True),
                                              (2, 2560, 320, True), (2, 64, 960, True)])
def test_groupnorm(ops, cuda_device, images, HW, C, silu):
    """InflatedGroupNorm resnet.py:27-37 / ResnetBlock2D norm+SiLU on channels-last rows, 32 groups."""
    x, g, b = randn(images * HW, C, seed=1), 1 + 0.1 * randn(C, seed=2), 0.1 * randn(C, seed=3)
    got = ops.groupnorm(bf(x).to(cuda_device), g.to(cuda_device), b.to(cuda_device), 1e-6, images, HW, silu=silu)
    xi = x.view(images, HW, C).permute(0, 2, 1)
    want = Fn.group_norm(xi, 32, g, b, 1e-6)
    want = Fn.silu(want) if silu else want
    assert rel(got, want.permute(0, 2, 1).reshape(images * HW, C)) < BF16_TOL


@pytest.mark.parametrize("images,HW,C,silu", [
    (3, 2560, 320, True),    # cluster of 8, 4 groups per 40-channel chunk
    (2, 1000, 320, False),   # cluster of 4, ragged rows (last CTA short)
    (3, 640, 640, True),     # cluster of 2, 2 groups per chunk
    (2, 160, 1280, True),    # one CTA per (image, group)
    (2, 40, 2560, True),     # 80-channel chunks
    (2, 640, 1920, True),    # 120-channel chunks (60 channels per group), cluster of 8
    (2, 37, 960, False),     # 30 channels per group: vectors straddle groups
    (1, 2560, 960, True),    # does not fit the single-pass kernel: three-kernel form behind the same call
])
@pytest.mark.parametrize("mode", ["1", "2", "3", "4"])
def test_groupnorm_single_pass(ops, cuda_device, monkeypatch, images, HW, C, silu, mode):
    """The single-pass cluster GroupNorm (FMC_GN_FUSED: 1 / 2 slab in registers, 3 / 4 slab staged by TMA) against torch
    and against the three-kernel form; two runs are bit-identical (fixed reduction order through distributed shared
    memory)."""
    x, g, b = randn(images * HW, C, seed=1) + 0.3, 1 + 0.1 * randn(C, seed=2), 0.1 * randn(C, seed=3)
    xd, gd, bd = bf(x).to(cuda_device), g.to(cuda_device), b.to(cuda_device)
    monkeypatch.setenv("FMC_GN_FUSED", mode)
    got = ops.groupnorm(xd, gd, bd, 1e-6, images, HW, silu=silu)
    again = ops.groupnorm(xd, gd, bd, 1e-6, images, HW, silu=silu)
    monkeypatch.setenv("FMC_GN_FUSED", "0")
    three = ops.groupnorm(xd, gd, bd, 1e-6, images, HW, silu=silu)
    assert torch.equal(got, again)
    want = Fn.group_norm(bf(x).float().view(images, HW, C).permute(0, 2, 1), 32, g, b, 1e-6)
    want = (Fn.silu(want) if silu else want).permute(0, 2, 1).reshape(images * HW, C)
    assert rel(got, want) < BF16_TOL
    assert rel(got, three.float().cpu()) < BF16_TOL
    # strided input / output rows (a channel slice of a wider tensor)
    wide = torch.zeros(images * HW, C + 64, device=cuda_device, dtype=torch.bfloat16)
    wide[:, 32:32 + C] = xd
    out_wide = torch.zeros_like(wide)
    monkeypatch.setenv("FMC_GN_FUSED", mode)
    ops.groupnorm(wide[:, 32:32 + C], gd, bd, 1e-6, images, HW, silu=silu, out=out_wide[:, 32:32 + C])
    assert torch.equal(out_wide[:, 32:32 + C], got)
    assert float(out_wide[:, :32].abs().max()) == 0.0 and float(out_wide[:, 32 + C:].abs().max()) == 0.0


@pytest.mark.parametrize("mode", ["1", "3"])
def test_groupnorm_single_pass_time_embedding_bias(ops, cuda_device, monkeypatch, mode):
    B, F, HW, C = 2, 3, 640, 640
    images = B * F
    x, g, b = randn(images * HW, C, seed=1), 1 + 0.1 * randn(C, seed=2), 0.1 * randn(C, seed=3)
    tb = randn(B, C, seed=4)
    monkeypatch.setenv("FMC_GN_FUSED", mode)
    got = ops.groupnorm(bf(x).to(cuda_device), g.to(cuda_device), b.to(cuda_device), 1e-5, images, HW, silu=True,
                        rowbias=tb.to(cuda_device), rowbias_div=F)
    xi = (bf(x).float().view(B, F, HW, C) + tb[:, None, None, :]).view(images, HW, C).permute(0, 2, 1)
    want = Fn.silu(Fn.group_norm(xi, 32, g, b, 1e-5)).permute(0, 2, 1).reshape(images * HW, C)
    assert rel(got, want) < BF16_TOL


def test_groupnorm_time_embedding_bias(ops, cuda_device):
    """h + time_emb_proj(silu(temb))[:, :, None, None] folded in front of norm2 of diffusers ResnetBlock2D."""
    B, F, HW, C = 2, 3, 64, 320
    images = B * F
    x, g, b = randn(images * HW, C, seed=1), 1 + 0.1 * randn(C, seed=2), 0.1 * randn(C, seed=3)
    tb = randn(B, C, seed=4)
    got = ops.groupnorm(bf(x).to(cuda_device), g.to(cuda_device), b.to(cuda_device), 1e-5, images, HW, silu=True,
                        rowbias=tb.to(cuda_device), rowbias_div=F)
    xi = (x.view(B, F, HW, C) + tb[:, None, None, :]).view(images, HW, C).permute(0, 2, 1)
    want = Fn.silu(Fn.group_norm(xi, 32, g, b, 1e-5)).permute(0, 2, 1).reshape(images * HW, C)
    assert rel(got, want) < BF16_TOL


def test_add_relu_rowbias(ops, cuda_device):
    a, b = randn(500, 640, seed=1), randn(500, 640, seed=2)
    got = ops.add(bf(a).to(cuda_device), bf(b).to(cuda_device))
    assert rel(got, a + b) < BF16_TOL
    got = ops.add(bf(a).to(cuda_device), relu=True)
    assert torch.equal(got.cpu().float(), torch.relu(a))


def test_resize_avgpool_copy(ops, cuda_device):
    x = randn(3, 5, 8, 64, seed=1)
    xd = bf(x).to(cuda_device)
    got = ops.resize_nearest(xd, 10, 16)
    want = Fn.interpolate(x.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(got.cpu().float(), want)
    got = ops.resize_nearest(xd, 9, 15)  # explicit size (Upsample2D output_size path)
    want = Fn.interpolate(x.permute(0, 3, 1, 2), size=(9, 15), mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(got.cpu().float(), want)
    x = randn(3, 6, 8, 64, seed=2)
    got = ops.avgpool2(bf(x).to(cuda_device))
    want = Fn.avg_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert rel(got, want) < BF16_TOL
    dst = torch.zeros(40, 96, dtype=torch.bfloat16, device=cuda_device)
    src = bf(randn(40, 32, seed=3)).to(cuda_device)
    ops.copy2d(src, dst[:, 64:])
    assert torch.equal(dst[:, 64:], src) and float(dst[:, :64].abs().sum()) == 0.0


def test_layout_round_trip(ops, cuda_device):
    x = randn(2, 4, 3, 5, 7, seed=1)
    cl = ops.to_channels_last(x.to(cuda_device), c_pad=8)
    assert cl.shape == (2, 3, 5, 7, 8)
    assert torch.equal(cl[..., :4].cpu().float(), x.permute(0, 2, 3, 4, 1))
    assert float(cl[..., 4:].abs().sum()) == 0.0
    back = ops.from_channels_last(cl, C=4)
    assert torch.equal(back.cpu(), x)


def test_timestep_embedding(ops, cuda_device):
    from oracle.diffusers_restated import Timesteps
    t = torch.tensor([961.0, 1.0, 500.0])
    got = ops.timestep_embedding(t.to(cuda_device), 320)
    want = Timesteps(320, True, 0)(t)
    assert rel(got, want) < BF16_TOL


# ---------------------------------------------------------------- rays, object scatter, mask modulate, DDIM
@pytest.mark.parametrize("b,f,H,W", [(1, 4, 64, 96), (2, 16, 256, 256), (1, 16, 320, 512)])
def test_plucker_rays(ops, cuda_device, b, f, H, W):
    """ray_condition fmc/data/dataset.py:930-972 (fp32) and the fused PixelUnshuffle(8) bf16 layout."""
    from oracle.rays import to_plucker_embedding
    from synfmc_b200 import synth
    K, c2w = synth.synth_camera(b, f, H, W, seed=5)
    want = to_plucker_embedding(c2w, K, (H, W))  # [b, f, 6, H, W]
    got = ops.plucker(K.view(b * f, 4).to(cuda_device), c2w.view(b * f, 3, 4).to(cuda_device), H, W)
    got = got.view(b, f, H, W, 6).permute(0, 1, 4, 2, 3)
    assert rel(got, want) < F32_TOL
    assert float((got.cpu() - want).abs().max()) < 2e-6
    un = ops.plucker_unshuffle(K.view(b * f, 4).to(cuda_device), c2w.view(b * f, 3, 4).to(cuda_device), H, W)
    want_un = Fn.pixel_unshuffle(want.reshape(b * f, 6, H, W), 8).permute(0, 2, 3, 1)
    assert un.shape == (b * f, H // 8, W // 8, 384)
    assert rel(un, want_un) < BF16_TOL


def test_ray_condition_as_the_trainers_call_it(cuda_device):
    """train_cam_ctrl.py:77-90: host tensors in, a 4x4 c2w with the bottom row, `device='cpu'`, a zero flip flag -- the
    result comes back on the CPU (built by the kernel on the current CUDA device), equal to the device='cuda' result."""
    from oracle.rays import to_plucker_embedding
    from synfmc_b200 import synth
    from synfmc_b200.fmc.data.dataset import ray_condition
    b, f, H, W = 1, 16, 64, 96
    K, c2w = synth.synth_camera(b, f, H, W, seed=8)
    bottom = torch.tensor([0, 0, 0, 1.0]).view(1, 1, 1, 4).expand(b, f, 1, 4)
    c2w44 = torch.cat([c2w, bottom], dim=2)
    flip = torch.zeros(16, dtype=torch.bool)
    on_cpu = ray_condition(K, c2w44, H, W, device="cpu", flip_flag=flip)
    on_gpu = ray_condition(K, c2w44, H, W, device=cuda_device, flip_flag=flip)
    assert on_cpu.device.type == "cpu" and on_gpu.is_cuda and on_cpu.shape == (b, f, H, W, 6)
    assert torch.equal(on_cpu, on_gpu.cpu())
    want = to_plucker_embedding(c2w, K, (H, W)).permute(0, 1, 3, 4, 2)
    assert rel(on_cpu, want) < F32_TOL


@pytest.mark.parametrize("n_obj,gaussian", [(1, True), (3, True), (3, False)])
def test_traj_scatter_bit_exact(ops, cuda_device, n_obj, gaussian):
    """get_traj_features_v2 fmc/util.py:161-200: last object with mask > 0 wins, channels (info*m)*m and m*m --
    bit-exact in fp32, and bit-exact after bf16 rounding in the fused unshuffled layout."""
    from oracle.util import build_traj_inputs
    from synfmc_b200 import synth
    from synfmc_b200.fmc.util import pack_objects
    b, f, H, W = 2, 3, 64, 96
    infos, masks = synth.synth_objects(b, f, H, W, n_obj, seed=7, gaussian=gaussian)
    traj, maskf = build_traj_inputs(infos, masks, "cpu", torch.float32)
    feats = torch.cat([traj, maskf], dim=-1) * maskf
    want = feats.permute(0, 1, 4, 2, 3).reshape(b * f, 13, H, W)
    info_d, masks_d = pack_objects(infos, masks, cuda_device)
    got, got_mask = ops.traj_scatter(info_d.view(b * f, n_obj, 12), masks_d.view(b * f, n_obj, H, W))
    assert torch.equal(got.cpu(), want)
    assert torch.equal(got_mask.cpu(), maskf.reshape(b * f, H, W))
    assert float(maskf.sum()) > 0
    un, un_mask = ops.traj_scatter_unshuffle(info_d.view(b * f, n_obj, 12), masks_d.view(b * f, n_obj, H, W))
    want_un = Fn.pixel_unshuffle(want, 8).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(un.cpu(), want_un)
    assert torch.equal(un_mask.cpu(), maskf.reshape(b * f, H, W))


@pytest.mark.parametrize("n_obj,H,W", [(1, 64, 96), (3, 128, 192), (3, 320, 512)])
def test_sphere_masks_and_fused_scatter(ops, cuda_device, n_obj, H, W):
    """On-device `use_sphere_mask` preprocessing (fmc/data/dataset.py:5350-5403): Gaussian disc masks from the minimum
    enclosing circles vs the numpy restatement (fp64) -- values to fp32 accuracy, IDENTICAL support -- and the scatter
    with the masks generated on the fly vs scatter(masks): bit-identical features and combined mask."""
    from oracle.sphere_mask import sphere_masks as o_masks
    from synfmc_b200 import synth
    b, f = 1, 4
    info, circles = synth.synth_circles(b, f, H, W, n_obj, seed=n_obj)
    want = torch.from_numpy(o_masks(circles.numpy(), H, W)).view(b * f, n_obj, H, W)
    got = ops.sphere_masks(circles.view(b * f, n_obj, 3).to(cuda_device), H, W)
    assert torch.equal(got.cpu() > 0, want > 0)
    assert float((got.cpu().double() - want).abs().max()) < 2e-6
    assert float(got.max()) <= 1.0 and float(want.max()) == 1.0
    if n_obj > 1:
        assert float(got.view(b, f, n_obj, H, W)[0, f // 2, n_obj - 1].abs().max()) == 0.0  # absent object
    d_info = info.view(b * f, n_obj, 12).to(cuda_device)
    feat2, mask2 = ops.traj_scatter_unshuffle(d_info, got)
    feat1, mask1 = ops.traj_scatter_circles_unshuffle(d_info, circles.view(b * f, n_obj, 3).to(cuda_device), H, W)
    assert torch.equal(mask1, mask2) and torch.equal(feat1, feat2)


def test_mask_modulate_iterated_nearest(ops, cuda_device):
    """fmc/adapter.py:175-177: the mask is resized level after level with F.interpolate(mode='nearest')."""
    from synfmc_b200.engine import nearest_index_chain
    N, H, W = 2, 320, 512
    mask = torch.rand(N, 1, H, W, generator=torch.Generator().manual_seed(3))
    sizes = [(40, 64), (20, 32), (10, 16), (5, 8)]
    hs, ws = [H], [W]
    m = mask
    for (h, w), C in zip(sizes, (64, 128, 256, 256)):
        m = Fn.interpolate(m, size=(h, w), mode="nearest")
        hs.append(h)
        ws.append(w)
        x = randn(N, h, w, C, seed=h)
        ry = nearest_index_chain(hs)[-1].to(cuda_device)
        rx = nearest_index_chain(ws)[-1].to(cuda_device)
        got = ops.mask_modulate(bf(x).to(cuda_device), mask[:, 0].contiguous().to(cuda_device), ry, rx)
        want = x * m.permute(0, 2, 3, 1)
        assert rel(got, want) < BF16_TOL


def test_cfg_ddim_step(ops, cuda_device):
    """CFG combine (pipeline_animation_cm_om.py:711-713) + diffusers DDIMScheduler.step(eta=0)."""
    from oracle.diffusers_restated import DDIMScheduler
    sch = DDIMScheduler()
    sch.set_timesteps(25)
    assert sch.timesteps.tolist()[:3] == [961, 921, 881] and sch.timesteps.tolist()[-1] == 1
    g = torch.Generator().manual_seed(0)
    x, eu, ec = (torch.randn(1, 4, 16, 8, 8, generator=g) for _ in range(3))
    from synfmc_b200.fmc._blocks import DDIMScheduler as PS
    ps = PS()
    ps.set_timesteps(25)
    assert ps.timesteps.tolist() == sch.timesteps.tolist()
    for t in (961, 41, 1):
        eps = eu + 8.0 * (ec - eu)
        want = sch.step(eps, t, x).prev_sample
        a_t, a_prev = ps.alphas_for(t)
        got = ops.cfg_ddim_step(eu.to(cuda_device), ec.to(cuda_device), 8.0, x.to(cuda_device), a_t, a_prev)
        assert rel(got, want) < F32_TOL


def test_cpu_tensors_are_rejected(ops):
    from synfmc_b200._cabi import FmcError
    with pytest.raises(FmcError):
        ops.add(torch.zeros(4, 8, dtype=torch.bfloat16))
    assert math.isfinite(1.0)
