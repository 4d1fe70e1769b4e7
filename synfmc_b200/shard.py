"""Clip sharding across the GPUs of one box: the reference is data-parallel only (DistributedSampler +
DDP, train_cam_ctrl.py:337-344,445); at inference the denoise loop of a clip never talks to another clip, so rank r
takes clips r, r + world, ... with NO data-path collective, and results are gathered once at the end.  One process per
GPU under torchrun; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is used only for the final gather, the
barrier around timed regions and the max-over-ranks reduction of device times."""
import os

import torch
import torch.distributed as dist


def world():
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when not launched distributed."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend=None, device=None):
    """Initialise the default process group when WORLD_SIZE > 1 (rendezvous from MASTER_ADDR / MASTER_PORT)."""
    rank, size, _ = world()
    if size > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=size, **kw)
    return rank, size


def clip_indices(n_clips, rank, size):
    """Clips owned by `rank`: r, r + size, ... (the DistributedSampler order without shuffling or padding)."""
    return list(range(rank, n_clips, size))


def gather_clips(local, n_clips, rank, size, dst=0):
    """local: list of per-clip tensors (same shape/dtype) for clip_indices(n_clips, rank, size).  Returns the list of
    all n_clips results in clip order on rank `dst` (None elsewhere).  One collective per call, after the loop."""
    if size == 1:
        return list(local)
    device = local[0].device if local else torch.device("cpu")
    counts = [len(clip_indices(n_clips, r, size)) for r in range(size)]
    template = None
    # every rank needs the clip shape to size the receive buffers; ranks with no clip learn it from rank 0
    meta = [None]
    if rank == 0:
        meta = [(tuple(local[0].shape), local[0].dtype)]
    dist.broadcast_object_list(meta, src=0)
    shape, dtype = meta[0]
    template = torch.zeros((max(counts),) + shape, dtype=dtype, device=device)
    mine = template.clone()
    for i, t in enumerate(local):
        mine[i].copy_(t)
    bufs = [torch.zeros_like(template) for _ in range(size)] if rank == dst else None
    dist.gather(mine, bufs, dst=dst)
    if rank != dst:
        return None
    out = [None] * n_clips
    for r in range(size):
        for i, clip in enumerate(clip_indices(n_clips, r, size)):
            out[clip] = bufs[r][i]
    return out


def max_over_ranks(value, device=None):
    """Max of a Python float over ranks (device times of a timed region are reported as the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    if t.device.type == "cpu" and dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------
# Single-clip latency mode ("optional"): the two classifier-free-guidance halves of ONE clip on two GPUs.
# Rank `which` = 0 evaluates the unconditional half of the U-Net batch, 1 the conditional half (pose features are the
# same for both, object features are zero -- i.e. absent -- for the unconditional half,
# pipeline_animation_cm_om.py:671-676); per step the halves exchange their noise prediction ([b, 4, f, h, w] fp32 =
# 0.66 MB at 320x512x16f) with ONE all-gather and both ranks apply the CFG combine + DDIM update redundantly, so the
# latents stay identical on both without a second message.
# ------------------------------------------------------------------------------------------------
def cfg_pair(rank=None, size=None):
    """(which, group, pair_index): consecutive ranks (2p, 2p + 1) form pair p.  Creates one process group per pair (every
    rank must call this, as torch.distributed.new_group is collective); size 2 uses the default group."""
    if rank is None or size is None:
        rank, size, _ = world()
    if size < 2 or size % 2:
        raise ValueError(f"the CFG-pair mode needs an even number of ranks, got {size}")
    if size == 2:
        return rank, None, 0
    mine = None
    for p in range(size // 2):
        g = dist.new_group(ranks=[2 * p, 2 * p + 1])
        if rank // 2 == p:
            mine = g
    return rank % 2, mine, rank // 2


def cfg_half(which, b, text_embeddings, features):
    """This rank's half of CFG-doubled inputs: text [2b, 77, 768] -> [b, 77, 768]; every feature [2b, ...] -> [b, ...]."""
    sl = slice(which * b, (which + 1) * b)
    return text_embeddings[sl], [f[sl] for f in features]


def cfg_exchange(eps_half, group=None):
    """all-gather of the two halves' noise predictions -> (eps_uncond, eps_cond), identical on both ranks"""
    eps_half = eps_half.contiguous()
    both = [torch.empty_like(eps_half), torch.empty_like(eps_half)]
    dist.all_gather(both, eps_half, group=group)
    return both[0], both[1]
