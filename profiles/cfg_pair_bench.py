#!/usr/bin/env python
"""Single-clip LATENCY mode (optional): one clip on two GPUs, one CFG half each, one 0.66 MB all-gather per
step (synfmc_b200.shard.cfg_pair, CameraCtrlPipeline.denoise_step(cfg_pair=...)).  Not the bench.py metric (that one is
whole-job throughput with one clip per GPU); this prints the per-step latency of ONE clip next to it.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        profiles/cfg_pair_bench.py --steps 10 --warmup 3
Checks inside the run: both ranks of a pair end with bit-identical latents, and (unless --no-check) the pair-mode latents
equal those of the ordinary doubled-batch step on rank 0 within bf16 tolerance."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload constants, build_product, synth_clip)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    from synfmc_b200 import shard
    from synfmc_b200.engine import CL
    from synfmc_b200.fmc.util import pack_objects, traj_features_cl
    rank, world, local = shard.world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    shard.init(backend="nccl", device=dev)
    which, group, pair = shard.cfg_pair()
    pipe, omcm = bench.build_product(dev)
    K, c2w, infos, masks, latents_h, text_h = bench.synth_clip(pair)   # both ranks of a pair hold the same clip
    feats = pipe.pose_encoder.encode_cameras(K.to(dev), c2w.to(dev), bench.H, bench.W)
    info_d, masks_d = pack_objects(infos, masks, dev)
    trajs = traj_features_cl(info_d, masks_d, omcm)
    b = latents_h.shape[0]
    text_full = text_h.to(dev)
    text_half = text_full[which * b:(which + 1) * b].contiguous()
    trajs_half = None if which == 0 else trajs          # unconditional half: no object features
    timesteps = pipe.scheduler.timesteps.tolist()

    def step(lat, i):
        t = timesteps[i % bench.SCHEDULE_STEPS]
        return pipe.denoise_step(lat, t, text_half, feats, bench.FRAMES,
                                 traj_features=trajs_half if t >= bench.OMCM_MIN_STEP else None,
                                 guidance_scale=bench.GUIDANCE, cfg_pair=(which, group))

    lat = latents_h.to(dev)
    for i in range(args.warmup):
        lat = step(lat, i)
    first_without = next(i for i, t in enumerate(timesteps) if t < bench.OMCM_MIN_STEP)
    step(lat, first_without)
    lat = latents_h.to(dev)
    shard.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        lat = step(lat, i)
    e1.record()
    shard.barrier()
    ms = shard.max_over_ranks(e0.elapsed_time(e1), device=dev)
    both = [torch.empty_like(lat), torch.empty_like(lat)]
    dist.all_gather(both, lat, group=group)
    identical = bool(torch.equal(both[0], both[1]))
    rel = None
    if not args.no_check and rank == 0:
        full_feats = [CL(torch.cat([f.t, f.t], dim=0)) for f in feats]
        full_trajs = [CL(torch.cat([torch.zeros_like(f.t), f.t], dim=0)) for f in trajs]
        ref = latents_h.to(dev)
        for i in range(args.steps):
            t = timesteps[i % bench.SCHEDULE_STEPS]
            ref = pipe.denoise_step(ref, t, text_full, full_feats, bench.FRAMES,
                                    traj_features=full_trajs if t >= bench.OMCM_MIN_STEP else None,
                                    guidance_scale=bench.GUIDANCE)
        rel = float((lat - ref).norm() / ref.norm())
    if rank == 0:
        print(json.dumps({"mode": "cfg-pair latency", "n_gpus": world, "clips": world // 2, "steps": args.steps,
                          "ms_per_step_one_clip": round(ms / args.steps, 3),
                          "steps_per_s_one_clip": round(args.steps / (ms * 1e-3), 3),
                          "pair_latents_bit_identical": identical, "rel_l2_vs_doubled_batch_step": rel,
                          "exchange_bytes_per_step": lat.numel() * 4}), flush=True)
    shard.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
