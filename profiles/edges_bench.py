#!/usr/bin/env python
"""SURVEY 8 f3 timing: the pipeline edges around the denoising loop at BASELINE configs[1] (1 clip 320x512x16f) --
decode_latents (16 frames through the SD1.5 VAE decoder), vae.encode of the same clip (the trainers' first step) and the CLIP
text encoder on a CFG pair of prompts -- on the B200 kernels.  CUDA events, warm-up first; one JSON line.  (The CPU time of
the fp32 restatements on the same box is printed by tests/test_gpu_edges.py: oracle/ is test infrastructure and is not
imported from here.)"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--trace", action="store_true")
    args = ap.parse_args()
    from synfmc_b200 import _cabi
    from synfmc_b200.edge import AutoencoderKL, CLIPTextModel
    from synfmc_b200.synth import synth_init_
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    vae, clip = AutoencoderKL(), CLIPTextModel()
    synth_init_(vae, seed=1)
    synth_init_(clip, seed=2)
    vae.to(dev)
    clip.to(dev)
    g = torch.Generator().manual_seed(0)
    latents = (torch.randn(1, 4, 16, 40, 64, generator=g) * 0.18215).to(dev)
    video = (torch.rand(16, 3, 320, 512, generator=g) * 2 - 1).to(dev)
    ids = torch.randint(0, 49408, (2, 77), generator=g).to(dev)
    noise = torch.randn(16, 4, 40, 64, generator=g).to(dev)
    out = {"what": "pipeline edges at 320x512x16f on one B200 (bf16 activations)", "steps": args.steps, "warmup": args.warmup}
    out["decode_latents_ms"] = round(timed(lambda: vae.decode_video(latents), args.steps, args.warmup), 3)
    out["decode_ms_per_frame"] = round(out["decode_latents_ms"] / 16, 3)
    out["vae_encode_16_frames_ms"] = round(timed(lambda: vae.encode(video).latent_dist.sample(noise=noise, scale=0.18215),
                                                 args.steps, args.warmup), 3)
    out["clip_text_2x77_ms"] = round(timed(lambda: clip(ids), args.steps * 4, args.warmup), 3)
    out["peak_memory_gib"] = round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2)
    if args.trace:
        _cabi.trace = []
        vae.decode_video(latents)
        torch.cuda.synchronize()
        trace, _cabi.trace = _cabi.trace, None
        agg = {}
        for name, a, s0, s1 in trace:
            d = agg.setdefault(name, [0, 0.0])
            d[0] += 1
            d[1] += s0.elapsed_time(s1)
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"{ms:9.2f} ms {n:6d} calls  {name}", file=sys.stderr)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
