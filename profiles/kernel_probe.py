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


def plucker_case(BF, H, W):
    K = torch.tensor([0.9 * W, 0.9 * W, W / 2.0, H / 2.0], device=dev).repeat(BF, 1)
    c2w = torch.eye(4, device=dev)[:3].repeat(BF, 1, 1) + 0.01 * torch.randn(BF, 3, 4, device=dev)
    return lambda: ops.plucker_unshuffle(K, c2w, H, W)


def traj_case(BF, n_obj, H, W):
    info = torch.randn(BF, n_obj, 12, device=dev)
    masks = (torch.rand(BF, n_obj, H, W, device=dev) - 0.6).clamp_min(0)
    return lambda: ops.traj_scatter_unshuffle(info, masks)


def mask_modulate_case(N, h, w, C, H=320, W=512):
    x = rnd(N, h, w, C)
    mask = torch.rand(N, H, W, device=dev)
    ry = (torch.arange(h, device=dev) * (H // h)).int()
    rx = (torch.arange(w, device=dev) * (W // w)).int()
    return lambda: ops.mask_modulate(x, mask, ry, rx)


def layernorm_pose_case(B, F, HW, C):
    rows = B * F * HW
    x, g, b = rnd(rows, C), rnd(C, dtype=torch.float32), rnd(C, dtype=torch.float32)
    pe, add = rnd(32, C, dtype=torch.float32), rnd(rows, C)
    return lambda: ops.layernorm(x, g, b, 1e-5, pe=pe, F=F, HW=HW, add=add)


def rowstats_case(rows, C):
    x = rnd(rows, C)
    return lambda: ops.rowstats(x)


def add_case(rows, C):
    a, b = rnd(rows, C), rnd(rows, C)
    return lambda: ops.add(a, b)


# algorithmic bytes of one call (DESIGN.md section 3 / SURVEY 8d) for the HBM-bound cases: GB/s = BYTES / time
BYTES = {
    "plucker_cfg2": 16 * 320 * 512 * 6 * 2,                                   # write-only: 31.5 MB per clip
    "traj_cfg2_1obj": 16 * 320 * 512 * (1 * 4 + 4 + 13 * 2),                  # masks read + mask written + 13 ch bf16
    "traj_cfg2_3obj": 16 * 320 * 512 * (3 * 4 + 4 + 13 * 2),
    "mask_mod_l0": 2 * 16 * 40 * 64 * 320 * 2, "mask_mod_l1": 2 * 16 * 20 * 32 * 640 * 2,
    "mask_mod_l2": 2 * 16 * 10 * 16 * 1280 * 2, "mask_mod_l3": 2 * 16 * 5 * 8 * 1280 * 2,
    "groupnorm_l0": 2 * 81920 * 320 * 2, "groupnorm_l1": 2 * 20480 * 640 * 2, "groupnorm_l0_cat": 2 * 81920 * 640 * 2,
    "layernorm_l0": 2 * 81920 * 320 * 2, "layernorm_l0_pose": 4 * 81920 * 320 * 2, "layernorm_l1_pose": 4 * 20480 * 640 * 2,
    "rowstats_l0": 81920 * 320 * 2, "add_l0": 3 * 81920 * 320 * 2,
}

CASES = {
    "plucker_cfg2": lambda: plucker_case(16, 320, 512),                 # one clip of config 2: rays -> unshuffled bf16
    "traj_cfg2_1obj": lambda: traj_case(16, 1, 320, 512),
    "traj_cfg2_3obj": lambda: traj_case(16, 3, 320, 512),
    "mask_mod_l0": lambda: mask_modulate_case(16, 40, 64, 320),
    "mask_mod_l1": lambda: mask_modulate_case(16, 20, 32, 640),
    "mask_mod_l2": lambda: mask_modulate_case(16, 10, 16, 1280),
    "mask_mod_l3": lambda: mask_modulate_case(16, 5, 8, 1280),
    "layernorm_l0_pose": lambda: layernorm_pose_case(2, 16, 2560, 320),
    "layernorm_l1_pose": lambda: layernorm_pose_case(2, 16, 640, 640),
    "groupnorm_l1": lambda: groupnorm_case(32, 640, 640),
    "groupnorm_l0_cat": lambda: groupnorm_case(32, 2560, 640),
    "rowstats_l0": lambda: rowstats_case(81920, 320),
    "add_l0": lambda: add_case(81920, 320),
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

def flush_l2():
    """write a buffer larger than the 126 MB L2 so the next call starts cold"""
    global _flush
    try:
        _flush.zero_()
    except NameError:
        _flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)


if __name__ == "__main__" and sys.argv[1] == "--once":
    # every named case exactly once, in order (for ONE ncu invocation capturing all of them: -k regex:... -c N)
    fns = [(c, CASES[c]()) for c in sys.argv[2:]]
    torch.cuda.synchronize()
    for c, fn in fns:
        fn()
    torch.cuda.synchronize()
    sys.exit(0)

if __name__ == "__main__" and sys.argv[1] == "--gbs":
    # HBM-bound cases: cold-L2 timing with CUDA events, achieved GB/s on the ALGORITHMIC bytes vs the measured peak
    import json
    peak = 6451.5
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs", peak)
    print(f"# cold-L2 (256 MB flush before every call), median of 9, CUDA events; peak = {peak} GB/s (MEASURED_PEAKS.json)")
    print(f"{'case':22s} {'us':>9s} {'MB(alg)':>9s} {'GB/s':>9s} {'frac':>6s}")
    for c in sys.argv[2:]:
        fn = CASES[c]()
        for _ in range(3):
            fn()
        ts = []
        for _ in range(9):
            flush_l2()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        ts.sort()
        us = ts[4]
        gbs = BYTES[c] / us * 1e-3
        print(f"{c:22s} {us:9.1f} {BYTES[c] / 1e6:9.1f} {gbs:9.1f} {gbs / peak:6.3f}")
    sys.exit(0)

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
