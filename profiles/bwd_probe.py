#!/usr/bin/env python
"""Backward kernels of the training step at their config-3 shapes (one clip 320x512x16f): us per call (CUDA events, warm-up,
L2 flushed between calls) and the algorithmic rate.  `--once` runs every case exactly once (for ncu captures).

    python profiles/bwd_probe.py
    ncu --set full --clock-control none -k regex:'attn_bwd|temporal16|groupnorm_bwd|layernorm_bwd|wgrad' -o gpurun_out/bwd \
        python profiles/bwd_probe.py --once"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BF16 = torch.bfloat16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    from synfmc_b200 import bwd_ops as B
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def rnd(*shape):
        return (torch.randn(*shape, generator=g, device=dev) * 0.5).to(BF16)
    cases = []

    def attn_case(name, images, n, d, inner=1, nk=None, kv_div=1):
        heads, hs = 8, (d + 15) // 16 * 16
        rows = images * n if inner == 1 else images * n
        k0, v0 = heads * hs, 2 * heads * hs
        o, do = rnd(rows, heads * d), rnd(rows, heads * d)
        if nk is None:  # self-attention on the fused buffer
            qkv = rnd(rows, 2 * heads * hs + heads * d)
            dqkv = torch.zeros_like(qkv)
            fn = lambda: B.attention_bwd(qkv, 0, qkv, k0, qkv, v0, hs, o, do, dqkv, 0, dqkv, k0, dqkv, v0, images, heads, d, n, n,
                                         1, n, inner, d ** -0.5)
            flops = 2.0 * images * heads * n * n * d * 5
        else:
            q, kv = rnd(rows, heads * hs), rnd(images // kv_div * 80, heads * hs + heads * d)
            dq = torch.zeros_like(q)
            fn = lambda: B.attention_bwd(q, 0, kv, 0, kv, k0, hs, o, do, dq, 0, None, 0, None, 0, images, heads, d, n, nk, kv_div,
                                         80, 1, d ** -0.5)
            flops = 2.0 * images * heads * n * nk * d * 3
        cases.append((name, fn, flops, None))
    attn_case("attention_bwd self L0 16x2560 d40 (tcgen05)", 16, 2560, 40)
    attn_case("attention_bwd self L1 16x640 d80 (tcgen05)", 16, 640, 80)
    attn_case("attention_bwd self L2 16x160 d160 (tcgen05)", 16, 160, 160)
    attn_case("attention_bwd text cross L0 16x2560x77 d40 (tcgen05, dQ)", 16, 2560, 40, nk=77, kv_div=16)
    attn_case("attention_bwd temporal L0 2560 seq x 16 frames d40 (warp per sequence)", 2560, 16, 40, inner=2560)
    attn_case("attention_bwd temporal L1 640 seq x 16 frames d80", 640, 16, 80, inner=640)
    for name, images, HW, C in (("groupnorm_bwd L0 16x2560x320", 16, 2560, 320), ("groupnorm_bwd L1 16x640x640", 16, 640, 640)):
        x, dy = rnd(images * HW, C), rnd(images * HW, C)
        gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        cases.append((name, (lambda x=x, dy=dy, gam=gam, bet=bet, images=images, HW=HW:
                             B.groupnorm_bwd(x, dy, gam, bet, 1e-6, images, HW, silu=True)), None, 3 * x.numel() * 2))
    for name, rows, C in (("layernorm_bwd L0 40960x320", 40960, 320), ("layernorm_bwd L2 2560x1280", 2560, 1280)):
        x, dy = rnd(rows, C), rnd(rows, C)
        gam = torch.ones(C, device=dev)
        cases.append((name, (lambda x=x, dy=dy, gam=gam: B.layernorm_bwd(x, dy, gam, 1e-5)), None, 3 * x.numel() * 2))
    for name, T, M, N in (("wgrad 320x320 over 40960 tokens (split-token TN kernel)", 40960, 320, 320),
                          ("wgrad 2560x320 over 40960 tokens", 40960, 2560, 320), ("wgrad 1280x1280 over 2560 tokens", 2560, 1280, 1280)):
        dy, x = rnd(T, M), rnd(T, N)
        cases.append((name, (lambda dy=dy, x=x: B.linear_wgrad(dy, x)), 2.0 * T * M * N, None))
    out = []
    for name, fn, flops, nbytes in cases:
        if args.once:
            fn()
            torch.cuda.synchronize()
            continue
        for _ in range(2):
            fn()
        ts = []
        for _ in range(args.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = sorted(ts)[len(ts) // 2]
        row = {"case": name, "us": round(us, 1)}
        if flops:
            row["tflops"] = round(flops / us / 1e6, 1)
        if nbytes:
            row["algorithmic_gbs"] = round(nbytes / us / 1e3, 1)
        out.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
