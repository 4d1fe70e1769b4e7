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
