#!/usr/bin/env python
"""BASELINE config 5: long-clip inference 320x512x64f, 50-step DDIM schedule, cam + 3 objects, batch-shard sweep (one clip
per GPU, no data-path collective; DESIGN.md "Config 5 -- definition used": 13 multidiff windows of 16 frames with overlap
12, per-window pose-embedding list, object features sliced per window).  One step = 13 CFG U-Net evaluations through the
captured graph + the window-average / DDIM kernel.  Not the bench.py metric (that is config 2); same timing rules.

    python profiles/cfg5_bench.py --steps 3 --warmup 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 \
        profiles/cfg5_bench.py --steps 3 --warmup 1"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

L, OVERLAP, N_WIN, N_OBJ, STEPS_SCHEDULE = 16, 12, 13, 3, 50
F_TOTAL = N_WIN * (L - OVERLAP) + OVERLAP  # 64


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    from synfmc_b200 import shard, synth
    from synfmc_b200.engine import CL
    from synfmc_b200.fmc.util import traj_features_from_circles
    rank, world, local = shard.world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    shard.init(backend="nccl", device=dev)
    pipe, omcm = bench.build_product(dev)
    pipe.scheduler.set_timesteps(STEPS_SCHEDULE)
    H, W = bench.H, bench.W
    K, c2w = synth.synth_camera(1, F_TOTAL, H, W, seed=500 + rank)
    info, circles = synth.synth_circles(1, F_TOTAL, H, W, N_OBJ, seed=500 + rank)
    latents, text = synth.synth_step_inputs(1, F_TOTAL, H // 8, W // 8, cfg=True, seed=500 + rank)
    latents, text = latents.to(dev), text.to(dev)

    # once per clip: CameraEncoder per window (its positional encoding ends at 16 frames), ObjectEncoder over all frames
    # from the objects' circles (Gaussian masks generated on the device), CFG duplication / zeroing
    def encoders():
        feats = []
        for k in range(N_WIN):
            s = k * (L - OVERLAP)
            f = pipe.pose_encoder.encode_cameras(K[:, s:s + L].to(dev), c2w[:, s:s + L].to(dev), H, W)
            feats.append([CL(torch.cat([x.t, x.t], dim=0)) for x in f])
        trajs = traj_features_from_circles(info.to(dev), circles.to(dev), omcm, H, W)
        trajs = [CL(torch.cat([torch.zeros_like(x.t), x.t], dim=0)) for x in trajs]
        return feats, trajs
    feats, trajs = encoders()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    feats, trajs = encoders()
    e1.record()
    torch.cuda.synchronize()
    encoders_ms = e0.elapsed_time(e1)
    win_trajs = pipe._window_features(trajs, L, N_WIN, OVERLAP)
    timesteps = pipe.scheduler.timesteps.tolist()

    def step(lat, i):
        t = timesteps[i % STEPS_SCHEDULE]
        return pipe.denoise_step(lat, t, text, feats, L, traj_features=win_trajs if t >= bench.OMCM_MIN_STEP else None,
                                 guidance_scale=bench.GUIDANCE, multidiff_total_steps=N_WIN, multidiff_overlaps=OVERLAP)
    lat = latents
    for i in range(args.warmup):
        lat = step(lat, i)
    step(lat, next(i for i, t in enumerate(timesteps) if t < bench.OMCM_MIN_STEP))  # capture the graph without objects
    lat = latents
    shard.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(args.steps):
        lat = step(lat, i)
    e.record()
    shard.barrier()
    ms = shard.max_over_ranks(s.elapsed_time(e), device=dev)
    assert bool(torch.isfinite(lat).all())
    if rank == 0:
        print(json.dumps({"metric": "denoise-steps/sec 320x512x64f (13 windows x CFG U-Net batch 2, cam list + 3 objects)",
                          "value": round(world * args.steps / (ms * 1e-3), 4), "unit": "steps/s", "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 2),
                          "unet_evaluations_per_step": N_WIN, "ms_per_unet_evaluation": round(ms / args.steps / N_WIN, 2),
                          "scaling": "weak", "dtype": "bf16", "data": "synthetic", "encoders_ms_per_clip": round(encoders_ms, 1),
                          "config": {"workload": "BASELINE configs[4]: 320x512x64f, 50-step DDIM schedule, cfg 8.0, cam + 3 "
                                                 "objects, 1 clip/GPU, multidiff windows 13 x 16 frames overlap 12"}}),
              flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
