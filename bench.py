#!/usr/bin/env python
"""Benchmark of the FMC denoising hot path on B200 (BASELINE.json: denoise-steps/sec at 320x512x16f).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one denoising iteration of CameraObjCtrlPipeline (pipeline_animation_cm_om.py:678-726) on ONE clip per
GPU: CFG-doubled U-Net forward (U-Net batch 2, cam + 1 object; object features only while t >= 700 as the reference
does), CFG combine and DDIM update.  BASELINE config 2 (configs[1]): 320x512x16f, 25-step schedule, bf16.  Weights and
inputs are synthetic (no checkpoints / datasets offline).  The CameraEncoder / ObjectEncoder run once per clip outside
the loop (as in the reference, :657) and are timed separately (`encoders_ms`).

Prints ONE JSON line (rank 0).  `value` = steps/s summed over all GPUs with inputs resident in HBM; `e2e` = the same
through the pipeline's public step call with pinned-host latents/text copied in and the new latents copied out every
step; `roofline` = the dominant kernel of this library, timed live with CUDA events on the launching stream over a
second pass of the same K steps; `cpu_baseline` = the fp32 CPU oracle on this host's cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, FRAMES, N_OBJ = 320, 512, 16, 1
SCHEDULE_STEPS, GUIDANCE, OMCM_MIN_STEP = 25, 8.0, 700
CHANNELS = (320, 640, 1280, 1280)
METRIC = "denoise-steps/sec 320x512x16f (CFG U-Net batch 2, cam + 1 object, DDIM)"
CONFIG = {"workload": "BASELINE configs[1]: 1 clip/GPU 320x512x16f, 25-step DDIM schedule, cfg 8.0, cam + 1 object "
                      "(configs/obj.yaml), bf16", "clips_per_gpu": 1, "frames": FRAMES, "height": H, "width": W,
          "unet_batch": 2, "objects": N_OBJ,
          "l2": "inputs larger than L2: each step streams 2.6 GB of bf16 weights + GBs of activations (L2 = 126 MB)"}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"],
                "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons of this rank's GPU, sampled every 200 ms during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id):
        self.gpu_id, self.proc, self.lines = str(gpu_id), None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.gpu_id, f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val == "Active":
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workload
def build_product(device):
    """Full SD1.5-shaped FMC U-Net (CamObj variant, LoRA rank C/2 + CameraAdapter processors), CameraEncoder and
    ObjectEncoder with synthetic name-seeded weights, on `device`."""
    from synfmc_b200.fmc._blocks import DDIMScheduler
    from synfmc_b200.fmc.adapter import Adapter
    from synfmc_b200.fmc.models.pose_adaptor import CameraPoseEncoder
    from synfmc_b200.fmc.models.unet_cam_obj import UNet3DConditionModelCamObjCond
    from synfmc_b200.fmc.modified_modules import bind_omcm_forwards
    from synfmc_b200.fmc.pipelines.pipeline_animation_cm_om import CameraObjCtrlPipeline
    from synfmc_b200.synth import synth_init_
    from synfmc_b200 import workload_config as wc
    cfg = wc.unet_config()
    unet = UNet3DConditionModelCamObjCond(**cfg)
    wc.set_processors(unet, CHANNELS)
    synth_init_(unet, seed=0)
    bind_omcm_forwards(unet)
    enc = synth_init_(CameraPoseEncoder(channels=list(CHANNELS), **wc.POSE_ENCODER_KWARGS), seed=1)
    omcm = synth_init_(Adapter(channels=list(CHANNELS), **wc.OMCM_KWARGS), seed=2)
    unet, enc, omcm = (m.to(device).eval().requires_grad_(False) for m in (unet, enc, omcm))  # inference: forward only
    sched = DDIMScheduler()
    sched.set_timesteps(SCHEDULE_STEPS)
    return CameraObjCtrlPipeline(None, None, None, unet, sched, enc), omcm


def synth_clip(rank):
    from synfmc_b200 import synth
    K, c2w = synth.synth_camera(1, FRAMES, H, W, seed=100 + rank)
    infos, masks = synth.synth_objects(1, FRAMES, H, W, N_OBJ, seed=100 + rank, gaussian=True)
    latents, text = synth.synth_step_inputs(1, FRAMES, H // 8, W // 8, cfg=True, seed=100 + rank)
    return K, c2w, infos, masks, latents, text


def flops_of(name, args):
    """Algorithmic FLOPs / bytes of one traced call (kernels table)."""
    if name in ("fmc_gemm_bf16", "fmc_gemm_ln_bf16"):
        M, N, K = args[6], args[7], args[8]
        return 2.0 * M * N * K, 0.0
    if name == "fmc_rowstats_bf16":
        rows, C = args[3], args[4]
        return 0.0, rows * C * 2.0
    if name == "fmc_conv3x3_bf16":
        images, Hh, Ww, cin, cout, stride = args[5], args[6], args[7], args[8], args[9], args[10]
        return 2.0 * images * (Hh // stride) * (Ww // stride) * cout * 9 * cin, 0.0
    if name == "cudnn_conv2d":
        M, N, K = args
        return 2.0 * M * N * K, 0.0
    if name == "fmc_spatial_attn_bf16":
        images, heads, d, nq, nk = args[14], args[15], args[16], args[17], args[18]
        return 4.0 * images * heads * nq * nk * d, 0.0
    if name == "fmc_temporal_attn_bf16":
        B, F, HW, heads, d = args[8], args[9], args[10], args[11], args[12]
        rows = B * F * HW
        return 4.0 * rows * F * heads * d, rows * heads * d * 2.0 * 4  # q,k,v read + o write (bf16)
    if name == "fmc_temporal_qkv_attn_bf16":
        B, F, HW, C = args[6], args[7], args[8], args[9]
        rows = B * F * HW
        return 2.0 * rows * C * 3 * C + 4.0 * rows * F * C, rows * C * 2.0 * 2
    if name == "fmc_layernorm_bf16":
        rows, C = args[14], args[15]
        n_tensors = 2 + (2 if args[10] else 0)
        return 0.0, rows * C * 2.0 * n_tensors
    if name == "fmc_groupnorm_bf16":
        images, HW, C = args[8], args[9], args[10]
        return 0.0, images * HW * C * 2.0 * 2  # algorithmic minimum: x read once, y written once (the single-pass kernel)
    if name == "fmc_add_bf16":
        rows, C = args[9], args[10]
        return 0.0, rows * C * 2.0 * (3 if args[2] else 2)
    return 0.0, 0.0


def summarise_trace(trace, steps, peaks):
    """Per-kernel time from the CUDA-event pairs of the traced pass.  Calls are grouped by (entry point, shape); a
    group contributes `median duration x number of calls`, so a stray host stall between a start event and its kernel
    (seen once in a while on the kernel-by-kernel launch path) does not leak into the kernel's figure."""
    import torch
    torch.cuda.synchronize()
    groups = {}
    for name, args, e0, e1 in trace:
        key = (name,) + tuple(a for a in args if isinstance(a, int) and abs(a) < (1 << 31))
        g = groups.setdefault(key, {"name": name, "each": [], "flops": 0.0, "bytes": 0.0})
        g["each"].append(e0.elapsed_time(e1))
        fl, by = flops_of(name, args)
        g["flops"] += fl
        g["bytes"] += by
    agg = {}
    for key, g in groups.items():
        g["each"].sort()
        n = len(g["each"])
        g["median"] = g["each"][n // 2]
        g["ms"] = g["median"] * n
        a = agg.setdefault(g["name"], {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        a["launches"] += n
        a["ms"] += g["ms"]
        a["flops"] += g["flops"]
        a["bytes"] += g["bytes"]
    total = sum(a["ms"] for a in agg.values()) or 1.0
    dump = os.environ.get("FMC_BENCH_TRACE")
    if dump:  # per-shape breakdown for kernel work (not part of the JSON line)
        with open(dump, "w") as fh:
            for key, g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"]):
                n, ms, each = len(g["each"]), g["ms"], g["each"]
                fh.write(f"{ms / steps:9.3f} ms/step  {n / steps:6.1f} calls/step  "
                         f"{g['flops'] / (ms * 1e-3) / 1e12 if ms else 0:8.1f} TF/s  "
                         f"[us min {each[0] * 1e3:.1f} med {g['median'] * 1e3:.1f} max {each[-1] * 1e3:.1f}]  {key}\n")
    table = {}
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        row = {"launches_per_step": round(a["launches"] / steps, 1), "ms_per_step": round(a["ms"] / steps, 3),
               "share_of_kernel_time": round(a["ms"] / total, 4)}
        if a["flops"]:
            row["tflops"] = round(a["flops"] / (a["ms"] * 1e-3) / 1e12, 1)
        elif a["bytes"]:
            row["gbs"] = round(a["bytes"] / (a["ms"] * 1e-3) / 1e9, 1)
        table[name] = row
    # the dominant KERNEL: fmc_gemm_bf16 and fmc_gemm_ln_bf16 are two entry points of one kernel (gemm_bf16_tma_kernel,
    # the second with the LayerNorm correction in its epilogue), so they are judged together
    same_kernel = {"fmc_gemm_ln_bf16": "fmc_gemm_bf16"}
    mine = {}
    for k, v in agg.items():
        if k.startswith("fmc_"):
            m = mine.setdefault(same_kernel.get(k, k), {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            for f in m:
                m[f] += v[f]
    top = max(mine, key=lambda k: mine[k]["ms"])
    a = mine[top]
    if top == "fmc_gemm_bf16" and "fmc_gemm_ln_bf16" in agg:
        top = "fmc_gemm_bf16 + fmc_gemm_ln_bf16 (one kernel: gemm_bf16_tma_kernel)"
    if a["flops"] and a["flops"] / (a["ms"] * 1e-3) / 1e12 > 1.0:
        achieved, peak, bound, unit = a["flops"] / (a["ms"] * 1e-3) / 1e12, peaks["tflops_sustained"], "tensor", "TFLOP/s"
    else:
        achieved, peak, bound, unit = a["bytes"] / (a["ms"] * 1e-3) / 1e9, peaks["hbm_gbs"], "hbm", "GB/s"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(top.split(" ")[0])  # dram bytes per launch (ncu), see the file's _source
    roofline = {"kernel": top, "bound": bound, "achieved": round(achieved, 1), "peak": peak, "unit": unit,
                "frac": round(achieved / peak, 4), "traffic": traffic,
                "peak_source": f"{peaks['source']} MEASURED_PEAKS.json, sustained figure (kernel timed inside a long step)",
                "avg_launch_us": round(a["ms"] * 1e3 / a["launches"], 2), "launches_per_step": a["launches"] / steps,
                "share_of_kernel_time": round(a["ms"] / total, 4)}
    return roofline, table


# ------------------------------------------------------------------------------------------------ CPU oracle legs
def oracle_steps():
    """Generator: each next() runs and times ONE WHOLE denoise step of the fp32 CPU oracle (oracle/, a restatement of the
    reference PyTorch path; the reference itself needs diffusers==0.24.0 which is not installable offline): the two CFG
    halves of the step -- unconditional text + zeroed object features, conditional text + object features
    (pipeline_animation_cm_om.py:671-676) -- as two U-Net forwards of batch 1 at 320x512x16f (one batch-2 forward
    materialises a 6.7 GB score tensor per level-0 attention; two batch-1 forwards are the same arithmetic), plus the CFG
    combine and DDIM update."""
    import torch
    from oracle import harness
    from oracle.diffusers_restated import DDIMScheduler
    torch.set_num_threads(os.cpu_count() or 1)
    _, _, _, _, latents, text = synth_clip(0)
    unet = harness.build_oracle_unet(tiny=False, obj=True)
    sched = DDIMScheduler()
    sched.set_timesteps(SCHEDULE_STEPS)
    g = torch.Generator().manual_seed(0)
    hh, ww = H // 8, W // 8
    # the encoders run once per clip outside the step; random features of the right shape stand in for them here
    feats = [torch.randn(1, c, FRAMES, hh >> l, ww >> l, generator=g) for l, c in enumerate(CHANNELS)]
    trajs = [0.5 * torch.randn(1, c, FRAMES, hh >> l, ww >> l, generator=g) for l, c in enumerate(CHANNELS)]
    zeros = [torch.zeros_like(t) for t in trajs]
    with torch.no_grad():
        while True:
            t0 = time.perf_counter()
            e_u = unet(latents, 961, text[:1], pose_embedding_features=feats, traj_features=zeros).sample
            e_c = unet(latents, 961, text[1:], pose_embedding_features=feats, traj_features=trajs).sample
            eps = e_u + GUIDANCE * (e_c - e_u)
            sched.step(eps, 961, latents)
            yield time.perf_counter() - t0


def run_reference(args):
    """`--impl reference`: the reference's CPU path (fp32 oracle port) on this host's cores, rank 0 only.  A step is one
    WHOLE denoise step (both CFG halves); the arm is bounded by a wall-clock budget: warm-ups are cut first, then the
    number of timed steps -- `steps` / `warmup` in the line are what was actually run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    budget_s = float(os.environ.get("FMC_REFERENCE_BUDGET_S", "240"))
    t_start = time.perf_counter()
    times, warm_done = [], 0
    for dt in oracle_steps():
        remaining = budget_s - (time.perf_counter() - t_start)
        if not times and warm_done < min(args.warmup, 1) and remaining > 2 * dt:
            warm_done += 1  # one untimed pass (thread pool / allocator warm-up) whenever a timed one still fits after it
            continue
        times.append(dt)
        if len(times) >= args.steps or remaining < dt:
            break
    steps_done = len(times)
    sec_per_step = sum(times) / steps_done
    value = 1.0 / sec_per_step
    sample = (f"whole denoise steps of the fp32 oracle port (2 U-Net forwards of batch 1 = the CFG pair, 320x512x16f, "
              f"{cores} threads): {steps_done} timed step(s) and {warm_done} warm-up(s) of the requested {args.steps} + "
              f"{args.warmup} fit the {int(budget_s)} s budget")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": steps_done, "warmup": warm_done, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "torch_threads": torch.get_num_threads()}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ main arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from synfmc_b200 import _cabi, ops, shard
    from synfmc_b200.engine import CL
    from synfmc_b200.fmc.util import pack_objects, traj_features_cl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the "
                         "CPU oracle)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    shard.init(backend="nccl", device=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    peaks = load_peaks()
    _cabi.lib()

    pipe, omcm = build_product(dev)
    pipe.use_cuda_graph = not os.environ.get("FMC_NO_GRAPH")  # default: each step replays a captured CUDA graph
    K, c2w, infos, masks, latents_h, text_h = synth_clip(rank)

    barrier = shard.barrier

    # --- once per clip: CameraEncoder + ObjectEncoder (outside the step, like pipeline_animation_cm_om.py:657-676)
    def encoders():
        feats = pipe.pose_encoder.encode_cameras(K.to(dev), c2w.to(dev), H, W)
        feats = [CL(torch.cat([f.t, f.t], dim=0)) for f in feats]
        info_d, masks_d = pack_objects(infos, masks, dev)
        trajs = traj_features_cl(info_d, masks_d, omcm)
        trajs = [CL(torch.cat([torch.zeros_like(f.t), f.t], dim=0)) for f in trajs]
        return feats, trajs
    feats, trajs = encoders()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    feats, trajs = encoders()
    e1.record()
    torch.cuda.synchronize()
    encoders_ms = e0.elapsed_time(e1)

    timesteps = pipe.scheduler.timesteps.tolist()
    text_d = text_h.to(dev)

    def step(latents, i, text):
        t = timesteps[i % SCHEDULE_STEPS]
        return pipe.denoise_step(latents, t, text, feats, FRAMES, traj_features=trajs if t >= OMCM_MIN_STEP else None,
                                 guidance_scale=GUIDANCE)

    def timed(fn, n):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn(n)
        e.record()
        barrier()
        return shard.max_over_ranks(s.elapsed_time(e), device=dev)

    # --- device-resident loop
    state = {"lat": latents_h.to(dev)}

    def resident(n, offset=0, head_start=False):
        lat = state["lat"]
        for i in range(n):
            if head_start:
                # per-kernel event timing: park the GPU for ~10 ms so that the (slower) kernel-by-kernel Python launch
                # path stays ahead of it; a starved GPU would add the CPU launch latency to every event interval
                torch.cuda._sleep(20_000_000)
            lat = step(lat, offset + i, text_d)
        state["lat"] = lat

    resident(args.warmup)
    # both step variants (with / without object features, t >= / < omcm_min_step) are warmed -- and their CUDA graphs
    # captured -- outside the timed region
    first_without = next(i for i, t in enumerate(timesteps) if t < OMCM_MIN_STEP)
    step(state["lat"], first_without, text_d)
    state["lat"] = latents_h.to(dev)
    sampler = ClockSampler(getattr(torch.cuda.get_device_properties(dev), "uuid", None) and
                           f"GPU-{torch.cuda.get_device_properties(dev).uuid}" or local)
    launches0 = _cabi.launch_count
    sampler.start()
    ms_total = timed(resident, args.steps)
    clocks = sampler.stop()
    gpu_launches = _cabi.launch_count - launches0
    assert bool(torch.isfinite(state["lat"]).all()), "non-finite latents"

    # --- end to end: pinned host latents + text in, latents out, every step, through the same public call
    lat_pin, text_pin = latents_h.clone().pin_memory(), text_h.clone().pin_memory()
    out_pin = torch.empty_like(lat_pin).pin_memory()
    lat_dev, text_dev = torch.empty_like(lat_pin, device=dev), torch.empty_like(text_pin, device=dev)

    def e2e(n):
        for i in range(n):
            lat_dev.copy_(lat_pin, non_blocking=True)
            text_dev.copy_(text_pin, non_blocking=True)
            new = step(lat_dev, i, text_dev)
            out_pin.copy_(new, non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the caller reads the result of every step
            lat_pin.copy_(out_pin)
    e2e(1)
    lat_pin.copy_(latents_h)
    ms_e2e = timed(e2e, args.steps)

    # --- per-kernel timing (second pass over the same K steps, CUDA events around every launch)
    # (kernel-by-kernel launches; one untimed pass first so that allocator growth / lazy cuDNN setup of the eager path
    # do not land between a start event and its kernel)
    graph_mode, pipe.use_cuda_graph = pipe.use_cuda_graph, False
    resident(2)
    step(state["lat"], first_without, text_d)
    state["lat"] = latents_h.to(dev)
    _cabi.trace = []
    resident(args.steps, head_start=True)
    trace, _cabi.trace = _cabi.trace, None
    pipe.use_cuda_graph = graph_mode
    roofline, table = summarise_trace(trace, args.steps, peaks)

    value = world * args.steps / (ms_total * 1e-3)
    line = {"metric": METRIC, "value": round(value, 4), "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": CONFIG,
            "e2e": {"value": round(world * args.steps / (ms_e2e * 1e-3), 4), "unit": "steps/s",
                    "h2d_bytes_per_step": lat_pin.numel() * 4 + text_pin.numel() * 4,
                    "d2h_bytes_per_step": out_pin.numel() * 4},
            "gpu_launches": gpu_launches, "cuda_graph": bool(pipe.use_cuda_graph), "clocks": clocks,
            "roofline": roofline, "kernels": table,
            "encoders_ms": round(encoders_ms, 2)}
    # once per clip, outside the step: CameraEncoder 1.52 TFLOP + ObjectEncoder 1.34 TFLOP
    line["kernels"]["encoders (once per clip: CameraEncoder + ObjectEncoder, all their kernels)"] = {
        "ms_per_clip": round(encoders_ms, 2), "tflops": round(2.86 / (encoders_ms * 1e-3), 1)}
    if world == 1 and not args.no_cpu_baseline:
        gen = oracle_steps()
        next(gen)  # untimed warm-up step (thread pool, allocator)
        sec = next(gen)
        line["cpu_baseline"] = {"value": round(1.0 / sec, 6), "unit": "steps/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "one whole denoise step of the fp32 oracle port (the CFG pair as 2 U-Net forwards "
                                          "of batch 1, 320x512x16f, all host threads) after one untimed warm-up step"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    del ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
