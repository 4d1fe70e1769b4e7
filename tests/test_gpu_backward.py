"""Backward kernels (csrc/backward.cu) against torch autograd on the same bf16-exact inputs (fp32 / fp64 reference
arithmetic).  Tolerances: bf16 outputs 6e-3 rel-L2 (one output rounding on top of recomputed statistics / probabilities),
fp32 outputs (parameter gradients, column sums) 1e-4."""
import pytest
import torch
import torch.nn.functional as Fn

pytestmark = pytest.mark.gpu

BF16_TOL = 6e-3
F32_TOL = 2e-4


def bf(x):
    return x.to(torch.bfloat16)


def rel(got, want):
    got, want = got.double().cpu(), want.double().cpu()
    return float((got - want).norm() / want.norm().clamp_min(1e-30))


def randn(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).float()


@pytest.fixture(scope="module")
def B(cuda_device):
    from synfmc_b200 import bwd_ops
    return bwd_ops


def test_transpose_and_colsum(B, cuda_device):
    x = randn(1000, 328, seed=1)
    assert torch.equal(B.transpose(bf(x).to(cuda_device)).float().cpu(), x.t())
    view = bf(randn(777, 640, seed=2)).to(cuda_device)[:, 64:384]
    assert torch.equal(B.transpose(view).cpu(), view.cpu().t())
    padded = B.transpose(bf(x[:301]).to(cuda_device), 8).float().cpu()
    assert padded.shape == (328, 304) and torch.equal(padded[:, :301], x[:301].t()) and float(padded[:, 301:].abs().max()) == 0
    got = B.colsum(bf(x).to(cuda_device))
    assert rel(got, x.double().sum(0)) < 1e-6
    acc = torch.ones(328, device=cuda_device)
    B.colsum(x.to(cuda_device), out=acc, accumulate=True)
    assert rel(acc, 1 + x.double().sum(0)) < 1e-6


@pytest.mark.parametrize("M,N,K", [(1000, 320, 320), (4096, 640, 1280), (300, 1280, 320)])
def test_linear_dgrad_wgrad(B, cuda_device, M, N, K):
    """y = x W^T: dX = dY W (GEMM against the transposed weight copy), dW = dY^T X (fp32 out), db = colsum(dY)."""
    x, w, dy = randn(M, K, seed=1), randn(N, K, seed=2, scale=K ** -0.5), randn(M, N, seed=3)
    w_t = bf(w.t().contiguous()).to(cuda_device)
    dx = B.linear_dgrad(bf(dy).to(cuda_device), w_t)
    assert rel(dx, dy.double() @ w.double()) < BF16_TOL
    dw = B.linear_wgrad(bf(dy).to(cuda_device), bf(x).to(cuda_device))
    assert dw.dtype == torch.float32 and rel(dw, dy.double().t() @ x.double()) < F32_TOL


@pytest.mark.parametrize("T,M,N", [(40960, 320, 320), (10240, 640, 2560), (300, 1280, 320), (1000, 328, 72), (64, 128, 128),
                                   (2560, 1280, 1280)])
def test_wgrad_tn(B, cuda_device, T, M, N):
    """fmc_wgrad_bf16: dW = dY^T X from the row-major operands (MN-major MMA operands, split token axis), plain and
    accumulating, on row-strided inputs; twice the same result (deterministic fold)."""
    dy_full, x_full = randn(T, M + 8, seed=1), randn(T, N + 16, seed=2)
    dy, x = bf(dy_full).to(cuda_device)[:, :M], bf(x_full).to(cuda_device)[:, 8:8 + N]
    want = bf(dy_full)[:, :M].double().t() @ bf(x_full)[:, 8:8 + N].double()
    dw = B.linear_wgrad(dy, x)
    assert dw.shape == (M, N) and dw.dtype == torch.float32 and rel(dw, want) < F32_TOL
    assert torch.equal(dw, B.linear_wgrad(dy, x))
    acc = torch.full((M, N), 2.0, device=cuda_device)
    B.linear_wgrad(dy, x, out=acc, accumulate=True)
    assert rel(acc, want + 2.0) < F32_TOL


@pytest.mark.parametrize("rows,C", [(1000, 320), (333, 640), (200, 1280)])
def test_layernorm_bwd(B, cuda_device, rows, C):
    x = randn(rows, C, seed=1, scale=2.0).requires_grad_(True)
    g = (1 + 0.1 * randn(C, seed=2)).requires_grad_(True)
    b = (0.1 * randn(C, seed=3)).requires_grad_(True)
    dy = randn(rows, C, seed=4)
    Fn.layer_norm(x.double(), (C,), g.double(), b.double(), 1e-5).backward(dy.double())
    dx, dg, db = B.layernorm_bwd(bf(x.detach()).to(cuda_device), bf(dy).to(cuda_device), g.detach().to(cuda_device), 1e-5,
                                 want_params=True)
    assert rel(dx, x.grad) < BF16_TOL
    assert rel(dg, g.grad) < F32_TOL and rel(db, b.grad) < F32_TOL
    dx2 = B.layernorm_bwd(bf(x.detach()).to(cuda_device), bf(dy).to(cuda_device), g.detach().to(cuda_device), 1e-5)
    assert torch.equal(dx, dx2)


@pytest.mark.parametrize("images,HW,C,silu,bias", [(4, 600, 320, True, True), (3, 160, 960, True, False),
                                                   (2, 40, 2560, False, False), (2, 100, 1280, True, True)])
def test_groupnorm_bwd(B, cuda_device, images, HW, C, silu, bias):
    x = (randn(images * HW, C, seed=1, scale=2.0) + 0.25).to(torch.bfloat16).float().requires_grad_(True)
    g, b = 1 + 0.1 * randn(C, seed=2), 0.1 * randn(C, seed=3)
    rb = randn(images // 2 if images % 2 == 0 else images, C, seed=4) if bias else None
    div = 2 if images % 2 == 0 else 1
    dy = randn(images * HW, C, seed=5)
    xin = x.double().view(images, HW, C)
    if bias:
        xin = xin + rb.double().repeat_interleave(div, 0)[:, None, :]
    y = Fn.group_norm(xin.permute(0, 2, 1), 32, g.double(), b.double(), 1e-5)
    if silu:
        y = Fn.silu(y)
    y.permute(0, 2, 1).reshape(images * HW, C).backward(dy.double())
    dx = B.groupnorm_bwd(bf(x.detach()).to(cuda_device), bf(dy).to(cuda_device), g.to(cuda_device), b.to(cuda_device), 1e-5,
                         images, HW, groups=32, silu=silu, rowbias=rb.to(cuda_device) if bias else None, rowbias_div=div)
    assert rel(dx, x.grad) < BF16_TOL


def test_geglu_fwd_bwd(B, cuda_device):
    from synfmc_b200.engine import _interleave_geglu
    M, H = 500, 1280
    proj = randn(M, 2 * H, seed=1).requires_grad_(True)   # reference layout: [value | gate]
    dy = randn(M, H, seed=2)
    a, g = proj.double().chunk(2, dim=-1)
    (a * Fn.gelu(g)).backward(dy.double())
    order = _interleave_geglu(torch.arange(2 * H).view(-1, 1).float(), None)[0].view(-1).long()
    inter = bf(proj.detach()[:, order]).to(cuda_device)  # the GEMM's interleaved column order
    y = B.geglu_fwd(inter)
    assert rel(y, a.detach() * Fn.gelu(g.detach())) < 4e-3
    dproj = B.geglu_bwd(inter, bf(dy).to(cuda_device))
    assert rel(dproj, proj.grad[:, order]) < BF16_TOL


def test_glue_bwd(B, cuda_device):
    y, dy = randn(4, 6, 10, 64, seed=1), randn(4, 6, 10, 64, seed=2)
    got = B.relu_bwd(bf(y).to(cuda_device), bf(dy).to(cuda_device))
    assert torch.equal(got.float().cpu(), dy * (y > 0))
    x = randn(2, 6, 10, 64, seed=3).requires_grad_(True)
    up = Fn.interpolate(x.permute(0, 3, 1, 2), size=(12, 20), mode="nearest")
    dyu = randn(2, 12, 20, 64, seed=4)
    up.backward(dyu.permute(0, 3, 1, 2))
    assert rel(B.resize_nearest_bwd(bf(dyu).to(cuda_device), 6, 10), x.grad) < 4e-3
    x2 = randn(2, 6, 10, 64, seed=5).requires_grad_(True)
    dyp = randn(2, 3, 5, 64, seed=6)
    Fn.avg_pool2d(x2.permute(0, 3, 1, 2), 2).backward(dyp.permute(0, 3, 1, 2))
    assert rel(B.avgpool2_bwd(bf(dyp).to(cuda_device), 6, 10), x2.grad) < 4e-3


def _pad_heads(x, heads, d, hs):
    if hs == d:
        return x
    out = torch.zeros(x.shape[0], heads, hs)
    out[:, :, :d] = x.view(x.shape[0], heads, d)
    return out.view(x.shape[0], heads * hs)


@pytest.mark.parametrize("d,images,n", [(40, 2, 300), (80, 2, 160), (160, 2, 40), (40, 1, 1000)])
def test_attention_bwd_spatial_self(B, cuda_device, d, images, n):
    """Spatial self-attention (attention_processor.py:148-154) on the fused [token, q|k|v] buffer with padded q / k heads."""
    from synfmc_b200 import ops
    heads, hs, C = 8, (d + 15) // 16 * 16, 8 * d
    q, k, v = (randn(images * n, C, seed=s).requires_grad_(True) for s in (1, 2, 3))
    do = randn(images * n, C, seed=4)

    def heads_view(t):
        return t.view(images, n, heads, d).transpose(1, 2)
    o = Fn.scaled_dot_product_attention(heads_view(q.double()), heads_view(k.double()), heads_view(v.double()))
    o = o.transpose(1, 2).reshape(images * n, C)
    o.backward(do.double())
    qkv = torch.cat([_pad_heads(q.detach(), heads, d, hs), _pad_heads(k.detach(), heads, d, hs), v.detach()], dim=1)
    dqkv = torch.zeros_like(qkv, dtype=torch.bfloat16, device=cuda_device)
    dev_qkv, dev_o, dev_do = bf(qkv).to(cuda_device), bf(o.detach().float()).to(cuda_device), bf(do).to(cuda_device)
    k0, v0 = heads * hs, 2 * heads * hs
    B.attention_bwd(dev_qkv, 0, dev_qkv, k0, dev_qkv, v0, hs, dev_o, dev_do, dqkv, 0, dqkv, k0, dqkv, v0, images, heads, d, n,
                    n, 1, n, 1, d ** -0.5)
    got = dqkv.float().cpu()
    dq = got[:, :k0].view(-1, heads, hs)[:, :, :d].reshape(-1, C)
    dk = got[:, k0:v0].view(-1, heads, hs)[:, :, :d].reshape(-1, C)
    dv = got[:, v0:]
    assert rel(dq, q.grad) < BF16_TOL and rel(dk, k.grad) < BF16_TOL and rel(dv, v.grad) < BF16_TOL
    if hs != d:  # the padding columns of the gradient buffer stay zero
        assert float(got[:, :k0].view(-1, heads, hs)[:, :, d:].abs().max()) == 0.0


@pytest.mark.parametrize("d,images,n", [(40, 3, 2560), (40, 2, 129), (40, 1, 128), (80, 3, 640), (80, 2, 200), (80, 1, 65), (160, 3, 160), (160, 2, 40), (160, 1, 300)])
def test_attention_bwd_tensor_core_matches_simt(B, cuda_device, monkeypatch, d, images, n):
    """head_dim 40 / 80 self-attention: the tcgen05 kernels (csrc/attn_bwd_tc.cu) against the SIMT kernels on the same
    buffers, including the level-0 / level-1 sequence lengths (2560, 640) and ragged last tiles."""
    from synfmc_b200 import ops
    heads, hs = 8, (d + 15) // 16 * 16
    qkv = torch.zeros(images * n, 2 * heads * hs + heads * d)
    qkv[:, :heads * hs] = _pad_heads(randn(images * n, heads * d, seed=1), heads, d, hs)
    qkv[:, heads * hs:2 * heads * hs] = _pad_heads(randn(images * n, heads * d, seed=2), heads, d, hs)
    qkv[:, 2 * heads * hs:] = randn(images * n, heads * d, seed=3)
    dev_qkv = bf(qkv).to(cuda_device)
    k0, v0 = heads * hs, 2 * heads * hs
    dev_o = torch.empty(images * n, heads * d, dtype=torch.bfloat16, device=cuda_device)
    ops.spatial_attn(dev_qkv, 0, dev_qkv, k0, dev_qkv, v0, hs, dev_o, images, heads, d, n, n, 1, n, d ** -0.5)
    dev_do = bf(randn(images * n, heads * d, seed=4)).to(cuda_device)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("FMC_ATTN_BWD_SIMT", mode)
        dqkv = torch.zeros_like(dev_qkv)
        B.attention_bwd(dev_qkv, 0, dev_qkv, k0, dev_qkv, v0, hs, dev_o, dev_do, dqkv, 0, dqkv, k0, dqkv, v0, images, heads, d,
                        n, n, 1, n, 1, d ** -0.5)
        torch.cuda.synchronize()
        out[mode] = dqkv.float().cpu()
    for lo, hi in ((0, k0), (k0, v0), (v0, qkv.shape[1])):
        assert rel(out["0"][:, lo:hi], out["1"][:, lo:hi]) < BF16_TOL
    if hs != d:
        assert float(out["0"][:, :v0].view(-1, 2 * heads, hs)[:, :, d:].abs().max()) == 0.0


@pytest.mark.parametrize("d,images,n", [(40, 2, 2560), (40, 3, 200), (80, 2, 640), (160, 2, 160)])
def test_attention_bwd_with_the_forward_log_sum_exp(B, cuda_device, d, images, n):
    """ops.spatial_attn(..., lse=) writes the row log-sum-exp (log2 units) the backward would recompute; given it, the
    tcgen05 dQ kernel skips its first sweep -- same gradients; the lse itself against torch."""
    from synfmc_b200 import ops
    heads, hs = 8, (d + 15) // 16 * 16
    rows = images * n
    qkv = torch.zeros(rows, 2 * heads * hs + heads * d)
    qkv[:, :heads * hs] = _pad_heads(randn(rows, heads * d, seed=1), heads, d, hs)
    qkv[:, heads * hs:2 * heads * hs] = _pad_heads(randn(rows, heads * d, seed=2), heads, d, hs)
    qkv[:, 2 * heads * hs:] = randn(rows, heads * d, seed=3)
    dev_qkv = bf(qkv).to(cuda_device)
    k0, v0 = heads * hs, 2 * heads * hs
    o_plain = torch.empty(rows, heads * d, dtype=torch.bfloat16, device=cuda_device)
    o_lse = torch.empty_like(o_plain)
    lse = torch.empty(rows, heads, device=cuda_device)
    ops.spatial_attn(dev_qkv, 0, dev_qkv, k0, dev_qkv, v0, hs, o_plain, images, heads, d, n, n, 1, n, d ** -0.5)
    ops.spatial_attn(dev_qkv, 0, dev_qkv, k0, dev_qkv, v0, hs, o_lse, images, heads, d, n, n, 1, n, d ** -0.5, lse=lse)
    assert torch.equal(o_plain, o_lse)
    qf = bf(qkv)[:, :k0].view(images, n, heads, hs)[..., :d].transpose(1, 2).double()
    kf = bf(qkv)[:, k0:v0].view(images, n, heads, hs)[..., :d].transpose(1, 2).double()
    want_lse = torch.logsumexp(qf @ kf.transpose(-1, -2) * d ** -0.5, dim=-1) / torch.log(torch.tensor(2.0, dtype=torch.float64))
    got_lse = lse.cpu().double().view(images, n, heads).transpose(1, 2)
    assert float((got_lse - want_lse).abs().max()) < 2e-3
    dev_do = bf(randn(rows, heads * d, seed=4)).to(cuda_device)
    out = {}
    for given in (False, True):
        dqkv = torch.zeros_like(dev_qkv)
        B.attention_bwd(dev_qkv, 0, dev_qkv, k0, dev_qkv, v0, hs, o_plain, dev_do, dqkv, 0, dqkv, k0, dqkv, v0, images, heads, d,
                        n, n, 1, n, 1, d ** -0.5, lse=lse.clone() if given else None)
        out[given] = dqkv.float().cpu()
    assert rel(out[True], out[False]) < 2e-3


@pytest.mark.parametrize("d,Bc,HW", [(40, 1, 300), (80, 2, 33), (160, 1, 9)])
def test_attention_bwd_temporal16_matches_simt(B, cuda_device, monkeypatch, d, Bc, HW):
    """16-frame temporal self-attention: the one-warp-per-sequence kernel against the generic SIMT kernels."""
    heads, hs, F = 8, (d + 15) // 16 * 16, 16
    rows = Bc * F * HW
    qkv = torch.zeros(rows, 2 * heads * hs + heads * d)
    qkv[:, :heads * hs] = _pad_heads(randn(rows, heads * d, seed=1), heads, d, hs)
    qkv[:, heads * hs:2 * heads * hs] = _pad_heads(randn(rows, heads * d, seed=2), heads, d, hs)
    qkv[:, 2 * heads * hs:] = randn(rows, heads * d, seed=3)
    dev_qkv = bf(qkv).to(cuda_device)
    k0, v0 = heads * hs, 2 * heads * hs
    dev_o = bf(randn(rows, heads * d, seed=5)).to(cuda_device)  # only the SIMT path reads O (D = sum dO O); see below
    from synfmc_b200 import ops
    ops.temporal_attn(dev_qkv, 0, k0, v0, hs, dev_o, Bc, F, HW, heads, d, d ** -0.5)
    dev_do = bf(randn(rows, heads * d, seed=4)).to(cuda_device)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("FMC_ATTN_BWD_SIMT", mode)
        dqkv = torch.zeros_like(dev_qkv)
        B.attention_bwd(dev_qkv, 0, dev_qkv, k0, dev_qkv, v0, hs, dev_o, dev_do, dqkv, 0, dqkv, k0, dqkv, v0, Bc * HW, heads, d,
                        F, F, 1, F, HW, d ** -0.5)
        torch.cuda.synchronize()
        out[mode] = dqkv.float().cpu()
    for lo, hi in ((0, k0), (k0, v0), (v0, qkv.shape[1])):
        assert rel(out["0"][:, lo:hi], out["1"][:, lo:hi]) < BF16_TOL


@pytest.mark.parametrize("d,images,nq,kv_div", [(40, 4, 200, 2), (160, 4, 64, 4), (80, 4, 150, 2), (40, 6, 2560, 3)])
def test_attention_bwd_text_cross(B, cuda_device, d, images, nq, kv_div):
    """Text cross-attention: dQ only (the text embeddings are frozen), 77 keys inside 80-row kv groups."""
    heads, hs, C, nk, stride = 8, (d + 15) // 16 * 16, 8 * d, 77, 80
    groups = images // kv_div
    q = randn(images * nq, C, seed=1).requires_grad_(True)
    k, v = randn(groups * stride, C, seed=2), randn(groups * stride, C, seed=3)
    do = randn(images * nq, C, seed=4)
    kh = k.view(groups, stride, heads, d)[:, :nk].transpose(1, 2).repeat_interleave(kv_div, 0).double()
    vh = v.view(groups, stride, heads, d)[:, :nk].transpose(1, 2).repeat_interleave(kv_div, 0).double()
    o = Fn.scaled_dot_product_attention(q.double().view(images, nq, heads, d).transpose(1, 2), kh, vh)
    o = o.transpose(1, 2).reshape(images * nq, C)
    o.backward(do.double())
    kv = torch.cat([_pad_heads(k, heads, d, hs), v], dim=1)
    dq = torch.zeros(images * nq, heads * hs, dtype=torch.bfloat16, device=cuda_device)
    dkv = bf(kv).to(cuda_device)
    B.attention_bwd(bf(_pad_heads(q.detach(), heads, d, hs)).to(cuda_device), 0, dkv, 0, dkv, heads * hs, hs,
                    bf(o.detach().float()).to(cuda_device), bf(do).to(cuda_device), dq, 0, None, 0, None, 0, images, heads, d,
                    nq, nk, kv_div, stride, 1, d ** -0.5)
    got = dq.float().cpu().view(-1, heads, hs)[:, :, :d].reshape(-1, C)
    assert rel(got, q.grad) < BF16_TOL


@pytest.mark.parametrize("d,Bc,F,HW", [(40, 2, 16, 60), (80, 1, 16, 20), (160, 1, 8, 9), (160, 1, 16, 9)])
def test_attention_bwd_temporal(B, cuda_device, d, Bc, F, HW):
    """Temporal attention over the frame axis of channels-last rows (inner = HW)."""
    heads, hs, C = 8, (d + 15) // 16 * 16, 8 * d
    rows = Bc * F * HW
    q, k, v = (randn(rows, C, seed=s).requires_grad_(True) for s in (1, 2, 3))
    do = randn(rows, C, seed=4)

    def seq(t):
        return t.view(Bc, F, HW, heads, d).permute(0, 2, 3, 1, 4)
    o = Fn.scaled_dot_product_attention(seq(q.double()), seq(k.double()), seq(v.double()))
    o = o.permute(0, 3, 1, 2, 4).reshape(rows, C)
    o.backward(do.double())
    qkv = torch.cat([_pad_heads(q.detach(), heads, d, hs), _pad_heads(k.detach(), heads, d, hs), v.detach()], dim=1)
    dqkv = torch.zeros_like(qkv, dtype=torch.bfloat16, device=cuda_device)
    dev_qkv = bf(qkv).to(cuda_device)
    k0, v0 = heads * hs, 2 * heads * hs
    B.attention_bwd(dev_qkv, 0, dev_qkv, k0, dev_qkv, v0, hs, bf(o.detach().float()).to(cuda_device), bf(do).to(cuda_device),
                    dqkv, 0, dqkv, k0, dqkv, v0, Bc * HW, heads, d, F, F, 1, F, HW, d ** -0.5)
    got = dqkv.float().cpu()
    dq = got[:, :k0].view(-1, heads, hs)[:, :, :d].reshape(-1, C)
    dk = got[:, k0:v0].view(-1, heads, hs)[:, :, :d].reshape(-1, C)
    assert rel(dq, q.grad) < BF16_TOL and rel(dk, k.grad) < BF16_TOL and rel(got[:, v0:], v.grad) < BF16_TOL
