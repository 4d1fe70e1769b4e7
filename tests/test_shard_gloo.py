"""World-size-2 (and 3) gloo runs of the clip-sharding plumbing on CPU: partition, end-of-loop gather in clip order,
max-over-ranks timing reduction (inference shards by clip with no data-path collective)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, size, port, n_clips, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(size), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    from synfmc_b200 import shard
    r, s = shard.init(backend="gloo")
    assert (r, s) == (rank, size) == shard.world()[:2]
    mine = shard.clip_indices(n_clips, rank, size)
    # stand-in for the per-clip denoise loop: a deterministic function of the clip index, no communication
    local = [torch.full((4, 2, 3), float(c)) + torch.arange(3.0) for c in mine]
    shard.barrier()
    out = shard.gather_clips(local, n_clips, rank, size)
    slow = shard.max_over_ranks(10.0 + rank)
    if rank == 0:
        ok = len(out) == n_clips and all(torch.equal(t, torch.full((4, 2, 3), float(c)) + torch.arange(3.0))
                                         for c, t in enumerate(out))
        q.put((ok, slow))
    else:
        assert out is None
    dist.destroy_process_group()


@pytest.mark.parametrize("size,n_clips", [(2, 2), (2, 5), (3, 4)])
def test_clip_sharding_gloo(size, n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, size, port, n_clips, q)) for r in range(size)]
    for p in procs:
        p.start()
    ok, slow = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok
    assert slow == 10.0 + size - 1


def test_partition_covers_every_clip_once():
    from synfmc_b200.shard import clip_indices
    for size in (1, 2, 4, 8):
        for n in (0, 1, 7, 8, 9):
            seen = sorted(c for r in range(size) for c in clip_indices(n, r, size))
            assert seen == list(range(n))


# ---------------------------------------------------------------------------------------------------------------
# single-clip latency mode: the two CFG halves on two ranks (synfmc_b200.shard.cfg_pair / cfg_half / cfg_exchange)
# ---------------------------------------------------------------------------------------------------------------
def _stub_unet(latents, text, feats):
    """stands in for the U-Net: a per-sample function of (latents, text, features), like the real one"""
    out = latents * text.mean(dim=(1, 2)).view(-1, 1, 1, 1, 1)
    for f in feats:
        out = out + f.mean(dim=(1, 2, 3, 4)).view(-1, 1, 1, 1, 1)
    return out


def _cfg_inputs(pair):
    g = torch.Generator().manual_seed(100 + pair)
    b = 1
    latents = torch.randn(b, 4, 4, 6, 8, generator=g)
    text = torch.randn(2 * b, 77, 16, generator=g)
    feats = [torch.randn(2 * b, 8 << l, 4, 6 >> l, 8 >> l, generator=g) for l in range(2)]
    return b, latents, text, feats


def _cfg_worker(rank, size, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(size), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    from synfmc_b200 import shard
    shard.init(backend="gloo")
    which, group, pair = shard.cfg_pair()
    assert (which, pair) == (rank % 2, rank // 2)
    b, latents, text, feats = _cfg_inputs(pair)
    text_h, feats_h = shard.cfg_half(which, b, text, feats)
    guidance = 8.0
    for _ in range(3):  # three "steps": the latents must stay identical on both ranks of a pair
        eps_half = _stub_unet(latents, text_h, feats_h)
        e_u, e_c = shard.cfg_exchange(eps_half, group)
        latents = latents - 0.1 * (e_u + guidance * (e_c - e_u))
    q.put((rank, latents))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("size", [2, 4])
def test_cfg_pair_mode_matches_the_doubled_batch(size):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cfg_worker, args=(r, size, port, q)) for r in range(size)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(size))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for pair in range(size // 2):
        b, latents, text, feats = _cfg_inputs(pair)
        for _ in range(3):  # the single-process reference: U-Net on the CFG-doubled batch, then the combine
            eps = _stub_unet(torch.cat([latents] * 2), text, feats)
            latents = latents - 0.1 * (eps[:b] + 8.0 * (eps[b:] - eps[:b]))
        assert torch.equal(got[2 * pair], got[2 * pair + 1])
        assert torch.allclose(got[2 * pair], latents, rtol=0, atol=1e-6)


def test_cfg_pair_needs_an_even_world():
    from synfmc_b200 import shard
    for size in (1, 3):
        with pytest.raises(ValueError):
            shard.cfg_pair(0, size)
