#!/usr/bin/env python
"""BASELINE configs 3 and 4: one TRAINING step per GPU on one clip 320x512x16f (the reference trains batch 1 per GPU,
configs/cam.yaml:149), data-parallel over N GPUs with the gradient all-reduce as the only collective.

  --stage cmc   train_cam_ctrl.py:586-665: PoseAdaptor forward (CameraEncoder + frozen U-Net with the CameraAdapter),
                MSE loss, backward, all-reduce of the 218 M trainable parameters (CameraEncoder + 20 qkv_merge), AdamW
  --stage omc   train_cam_obj_ctrl.py:843-866: 3 objects per clip -> ObjectEncoder (trainable, 152.5 M) -> CamObjPoseAdaptor

One step = zero_grad, forward on the tape, loss, backward (bucket all-reduces launched from gradient hooks), fused
unscale + clip + AdamW.  Timed with CUDA events, barrier + synchronize on both sides, max over ranks.

    python profiles/train_step_bench.py --stage cmc --steps 3 --warmup 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        profiles/train_step_bench.py --stage cmc"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default="cmc", choices=["cmc", "omc"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--graph", action="store_true", help="capture the whole step into a CUDA graph (train.GraphedStep) and time "
                    "replays; the warm-up steps run eagerly before the capture")
    ap.add_argument("--profile", action="store_true", help="one extra step under torch.profiler (CUPTI): top kernels by "
                    "device time on stderr, cuDNN and torch kernels included")
    ap.add_argument("--trace", action="store_true", help="one extra step with CUDA events around every library call: "
                                                         "per-entry-point time table on stderr")
    args = ap.parse_args()
    from synfmc_b200 import shard, synth
    from synfmc_b200.fmc.models.pose_adaptor import PoseAdaptor
    from synfmc_b200.fmc.models.pose_obj_adaptor import CamObjPoseAdaptor
    from synfmc_b200.fmc.util import get_traj_features_v2, pack_objects, traj_features_packed
    from synfmc_b200.train import FlatParams, FusedAdamW, GradAllReduce, GraphedStep
    rank, world, local = shard.world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    shard.init(backend="nccl", device=dev)
    pipe, omcm = bench.build_product(dev)
    unet, enc = pipe.unet, pipe.pose_encoder
    if args.stage == "cmc":
        enc.requires_grad_(True)
        for n, p in unet.named_parameters():
            if "merge" in n and "lora" not in n:
                p.requires_grad_(True)
        trainable = [p for p in enc.parameters()] + [p for p in unet.parameters() if p.requires_grad]
        wrapper = PoseAdaptor(unet, enc)
    else:
        omcm.requires_grad_(True)
        trainable = list(omcm.parameters())
        wrapper = CamObjPoseAdaptor(unet, enc)
    flat = FlatParams(trainable)
    red = GradAllReduce(flat, bucket_bytes=256 << 20).install_hooks()
    opt = FusedAdamW(flat, lr=1e-4, max_grad_norm=1.0)
    H, W, F = bench.H, bench.W, bench.FRAMES
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_free_inputs import plucker_embedding  # the product's own ray kernel: oracle/ stays out of the bench
    K, c2w = synth.synth_camera(1, F, H, W, seed=300 + rank)
    pose = plucker_embedding(K.to(dev), c2w.to(dev), H, W)            # [1, 6, F, H, W] built by fmc_plucker_f32
    latents, text = synth.synth_step_inputs(1, F, H // 8, W // 8, cfg=False, seed=300 + rank)
    latents, text = latents.to(dev), text.to(dev)
    target = torch.randn(latents.shape, generator=torch.Generator().manual_seed(rank)).to(dev)
    infos, masks = synth.synth_objects(1, F, H, W, 3, seed=300 + rank, gaussian=True) if args.stage == "omc" else (None, None)
    packed = pack_objects(infos, masks, dev) if infos is not None else None  # data loading: outside the (captured) step
    t = torch.tensor([801], device=dev)
    losses = []

    def step():
        opt.zero_grad()
        red.reset()
        if args.stage == "omc":
            trajs = (traj_features_packed(*packed, omcm) if args.graph else
                     get_traj_features_v2(infos, masks, omcm, False, 0.0, None, dev, torch.float32))
            pred = wrapper(latents, t, text, pose, trajs)
        else:
            pred = wrapper(latents, t, text, pose)
        loss = torch.nn.functional.mse_loss(pred.float(), target)
        loss.backward()
        n = red.wait()
        opt.step(loss_scale=1.0, world=n, device_state=args.graph)
        if not args.graph:
            losses.append(loss.detach())
        return loss.detach()
    if args.graph:
        eager = step
        graphed = GraphedStep(eager, warmup=max(args.warmup, 2))

        def step():
            losses.append(graphed().clone())
    else:
        for _ in range(args.warmup):
            step()
    shard.barrier()
    mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    shard.barrier()
    ms = shard.max_over_ranks(e0.elapsed_time(e1), device=dev) / args.steps
    ls = [float(x) for x in losses]
    if args.trace and rank == 0:
        from synfmc_b200 import _cabi
        _cabi.trace = []
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        step()
        t1.record()
        torch.cuda.synchronize()
        trace, _cabi.trace = _cabi.trace, None
        agg, shapes = {}, {}
        for name, a, s0, s1 in trace:
            if name == "fmc_attention_bwd_bf16":  # split by shape: images x head_dim, nq x nk, inner, self / cross
                name += (f"[img={a[-10]} d={a[-8]} nq={a[-7]} nk={a[-6]} inner={a[-3]} "
                         f"{'self' if a[17] else 'dQ only'}]")
            elif name == "fmc_gemm_bf16":  # M x N x K, fp32 output (weight gradients) flagged
                shapes.setdefault(f"gemm M={a[6]} N={a[7]} K={a[8]} flags={a[15]}", [0, 0.0])
                sh = shapes[f"gemm M={a[6]} N={a[7]} K={a[8]} flags={a[15]}"]
                sh[0] += 1
                sh[1] += s0.elapsed_time(s1)
            elif name == "fmc_layernorm_bwd_bf16":
                sh = shapes.setdefault(f"layernorm_bwd rows={a[9]} C={a[10]} params={'yes' if a[8] else 'no'}", [0, 0.0])
                sh[0] += 1
                sh[1] += s0.elapsed_time(s1)
            elif name == "fmc_groupnorm_bwd_bf16":
                key = f"groupnorm_bwd images={a[10]} HW={a[11]} C={a[12]}"
                sh = shapes.setdefault(key, [0, 0.0])
                sh[0] += 1
                sh[1] += s0.elapsed_time(s1)
            d = agg.setdefault(name, [0, 0.0])
            d[0] += 1
            d[1] += s0.elapsed_time(s1)
        total = sum(v[1] for v in agg.values())
        print(f"# traced step: {t0.elapsed_time(t1):.1f} ms wall on the device, {total:.1f} ms inside library calls "
              f"(the rest: cuDNN convolutions forward / backward, torch glue)", file=sys.stderr)
        for name, (n, ms_) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"{ms_:9.2f} ms {n:6d} calls  {name}", file=sys.stderr)
        print("# by shape (top 25)", file=sys.stderr)
        for name, (n, ms_) in sorted(shapes.items(), key=lambda kv: -kv[1][1])[:25]:
            print(f"{ms_:9.2f} ms {n:6d} calls  {name}", file=sys.stderr)
    if args.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        rows_ = sorted(((e.key, e.count, e.device_time_total / 1e3) for e in prof.key_averages() if e.device_time_total > 0
                        and e.device_type == torch.autograd.DeviceType.CUDA), key=lambda r: -r[2])
        print(f"# torch.profiler: {sum(r[2] for r in rows_):.1f} ms of kernels in the step", file=sys.stderr)
        for name, n, ms_ in rows_[:45]:
            print(f"{ms_:9.2f} ms {n:6d} x  {name[:150]}", file=sys.stderr)
    if rank == 0:
        print(json.dumps({"metric": f"training steps/sec, {args.stage.upper()} stage, 1 clip 320x512x16f per GPU",
                          "value": round(world / (ms * 1e-3), 4), "unit": "steps/s (clips/s)", "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 2), "scaling": "weak",
                          "dtype": "bf16 activations / gradients, fp32 master weights", "data": "synthetic",
                          "cuda_graph": bool(args.graph), "optimizer_steps_on_device": opt.steps_taken(),
                          "trainable_params_m": round(flat.numel / 1e6, 1), "allreduce_buckets": len(red.buckets),
                          "peak_memory_gib": round(mem, 1), "losses": [round(x, 5) for x in ls],
                          "grad_norm_last": opt.last_norm(), "found_inf": opt.found_inf()}), flush=True)
    if world > 1:
        import torch.distributed as dist
        if args.graph:
            # a captured graph holds NCCL kernels of this communicator: tearing the group down under it was seen to hang
            # (round 2, call 15); the result line is out, leave without the collective teardown
            sys.stdout.flush()
            torch.cuda.synchronize()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
