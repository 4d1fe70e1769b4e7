#!/usr/bin/env python
"""Launch ONE hot-path kernel at its BASELINE config-2 shape a few times (for `ncu --set full -k regex:... -s 2 -c 1`).
usage: python profiles/kernel_probe.py <case> [reps]      cases: see CASES"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synfmc_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
BF = torch.bfloat16


def rnd(*shape, scale=1.0, dtype=BF):
    return (torch.randn(*shape, device=dev) * scale).to(dtype)


def gemm_case(M, N, K, res=False, geglu=False):
    a, w = rnd(M, K), rnd(N, K, scale=K ** -0.5)
    bias = rnd(N, dtype=torch.float32)
    r = rnd(M, N // 2 if geglu else N) if res else None
    tile_n = int(os.environ.get("FMC_PROBE_TILE_N", "0"))
    return lambda: ops.gemm(a, w, bias=bias, residual=r, geglu=geglu, tile_n=tile_n)


def spatial_case(images, d, n):
    heads, hs = 8, (d + 15) // 16 * 16
    qkv = rnd(images * n, 2 * heads * hs + heads * d)
    out = torch.empty(images * n, heads * d, device=dev, dtype=BF)
    return lambda: ops.spatial_attn(qkv, 0, qkv, heads * hs, qkv, 2 * heads * hs, hs, out, images, heads, d, n, n, 1, n,
                                    d ** -0.5)


def cross_case(images, d, n, frames=16, nk=77):
    heads, hs = 8, (d + 15) // 16 * 16
    q = rnd(images * n, heads * hs)
    kv = rnd((images // frames) * 80, heads * hs + heads * d)
    out = torch.empty(images * n, heads * d, device=dev, dtype=BF)
    return lambda: ops.spatial_attn(q, 0, kv, 0, kv, heads * hs, hs, out, images, heads, d, n, nk, frames, 80, d ** -0.5)


def spatial_f16_case(images, n):
    heads, d, hs = 8, 40, 48
    qkv = rnd(images * n, 2 * heads * hs + heads * d)
    qkv[:, 2 * heads * hs:] = torch.randn(images * n, heads * d, device=dev).half().view(BF)
    out = torch.empty(images * n, heads * d, device=dev, dtype=BF)
    return lambda: ops.spatial_attn(qkv, 0, qkv, heads * hs, qkv, 2 * heads * hs, hs, out, images, heads, d, n, n, 1, n,
                                    d ** -0.5, v_f16=True)


def temporal_case(B, F, HW, d):
    heads, hs = 8, (d + 15) // 16 * 16
    qkv = rnd(B * F * HW, 2 * heads * hs + heads * d)
    out = torch.empty(B * F * HW, heads * d, device=dev, dtype=BF)
    return lambda: ops.temporal_attn(qkv, 0, heads * hs, 2 * heads * hs, hs, out, B, F, HW, heads, d, d ** -0.5)


def temporal_fused_case(B, F, HW):
    x = rnd(B * F * HW, 320)
    w = rnd(8 * 128, 320, scale=320 ** -0.5)
    out = torch.empty(B * F * HW, 320, device=dev, dtype=BF)
    return lambda: ops.temporal_qkv_attn(x, w, out, B, F, HW, 8, 40 ** -0.5)


def conv_case(N, H, W, cin, cout, stride=1, cudnn=False):
    x = rnd(N, H, W, cin)
    w = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5)
    if cudnn:
        wc = w.contiguous(memory_format=torch.channels_last)
        return lambda: ops.conv2d_cl(x, wc, None, stride=stride, padding=1)
    w2d = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    return lambda: ops.conv3x3(x, w2d, stride=stride)


def groupnorm_case(images, HW, C):
    x, g, b = rnd(images * HW, C), rnd(C, dtype=torch.float32), rnd(C, dtype=torch.float32)
    return lambda: ops.groupnorm(x, g, b, 1e-6, images, HW, silu=True)


def layernorm_case(rows, C):
    x, g, b = rnd(rows, C), rnd(C, dtype=torch.float32), rnd(C, dtype=torch.float32)
    return lambda: ops.layernorm(x, g, b, 1e-5)


CASES = {
    "gemm_l0_out": lambda: gemm_case(81920, 320, 320, res=True),        # to_out / proj_out + residual at level 0
    "gemm_l0_qkv": lambda: gemm_case(81920, 1088, 320),                 # fused q|k|v (heads padded 40 -> 48)
    "gemm_l0_geglu": lambda: gemm_case(81920, 2560, 320, geglu=True),   # FeedForward GEGLU at level 0
    "gemm_l0_ff2": lambda: gemm_case(81920, 320, 1280, res=True),
    "gemm_l1_geglu": lambda: gemm_case(20480, 5120, 640, geglu=True),
    "gemm_l2_qkv": lambda: gemm_case(5120, 3840, 1280),
    "gemm_l1_qkv": lambda: gemm_case(20480, 1920, 640),
    "gemm_l1_out": lambda: gemm_case(20480, 640, 640, res=True),
    "gemm_l2_out": lambda: gemm_case(5120, 1280, 1280, res=True),
    "gemm_l2_geglu": lambda: gemm_case(5120, 10240, 1280, geglu=True),
    "gemm_l2_ff2": lambda: gemm_case(5120, 1280, 5120, res=True),
    "gemm_l1_ff2": lambda: gemm_case(20480, 640, 2560, res=True),
    "conv_l0": lambda: conv_case(32, 40, 64, 320, 320),
    "conv_l0_cudnn": lambda: conv_case(32, 40, 64, 320, 320, cudnn=True),
    "conv_l0b": lambda: conv_case(32, 40, 64, 640, 320),
    "conv_l0b_cudnn": lambda: conv_case(32, 40, 64, 640, 320, cudnn=True),
    "conv_l1": lambda: conv_case(32, 20, 32, 640, 640),
    "conv_l1_cudnn": lambda: conv_case(32, 20, 32, 640, 640, cudnn=True),
    "conv_l2": lambda: conv_case(32, 10, 16, 1280, 1280),
    "conv_l2_cudnn": lambda: conv_case(32, 10, 16, 1280, 1280, cudnn=True),
    "conv_l3": lambda: conv_case(32, 5, 8, 1280, 1280),
    "conv_l3_cudnn": lambda: conv_case(32, 5, 8, 1280, 1280, cudnn=True),
    "conv_down0": lambda: conv_case(32, 40, 64, 320, 320, stride=2),
    "conv_down0_cudnn": lambda: conv_case(32, 40, 64, 320, 320, stride=2, cudnn=True),
    "cross_l0": lambda: cross_case(32, 40, 2560),
    "cross_l1": lambda: cross_case(32, 80, 640),
    "spatial_l0_f16": lambda: spatial_f16_case(32, 2560),
    "spatial_l0": lambda: spatial_case(32, 40, 2560),
    "spatial_l1": lambda: spatial_case(32, 80, 640),
    "spatial_l2": lambda: spatial_case(32, 160, 160),
    "cross_l2": lambda: cross_case(32, 160, 160),
    "temporal_fused_l0": lambda: temporal_fused_case(2, 16, 2560),
    "temporal_l1": lambda: temporal_case(2, 16, 640, 80),
    "temporal_l2": lambda: temporal_case(2, 16, 160, 160),
    "temporal_l0": lambda: temporal_case(2, 16, 2560, 40),
    "groupnorm_l0": lambda: groupnorm_case(32, 2560, 320),
    "layernorm_l0": lambda: layernorm_case(81920, 320),
}

if __name__ == "__main__":
    fn = CASES[sys.argv[1]]()
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e) / reps * 1e3)
    times.sort()
    print(f"{sys.argv[1]} [{os.environ.get('FMC_B200_LIB', 'default')}]: min {times[0]:.1f} median {times[2]:.1f} us per call")
