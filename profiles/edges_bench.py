#!/usr/bin/env python
"""SURVEY 8 f3 timing: the pipeline edges around the denoising loop at BASELINE configs[1] (1 clip 320x512x16f) --
decode_latents (16 frames through the SD1.5 VAE decoder), vae.encode of the same clip (the trainers' first step) and the CLIP
text encoder on a CFG pair of prompts -- on the B200 kernels, next to the CPU restatement on a bounded sample (1 frame /
1 prompt pair).  CUDA events, warm-up first; one JSON line."""
import argparse
import json
import os
import sys
import time

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
    ap.add_argument("--cpu", action="store_true", help="also time the CPU restatement (1 frame, 1 prompt pair)")
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
    if args.cpu:
        from oracle.clip_text import CLIPTextModel as OC
        from oracle.vae import AutoencoderKL as OV
        ov, oc = OV().eval().requires_grad_(False), OC().eval().requires_grad_(False)
        ov.load_state_dict(vae.state_dict())
        oc.load_state_dict(clip.state_dict())
        with torch.no_grad():
            z1 = (latents[:, :, 0] / 0.18215).cpu()
            ov.decode(z1)
            t0 = time.perf_counter()
            ov.decode(z1)
            out["cpu_decode_ms_per_frame"] = round((time.perf_counter() - t0) * 1e3, 1)
            oc(ids.cpu())
            t0 = time.perf_counter()
            oc(ids.cpu())
            out["cpu_clip_text_2x77_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
            out["cpu_threads"] = torch.get_num_threads()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
