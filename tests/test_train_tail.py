"""Collective + optimizer half of the training step (synfmc_b200/train.py): host logic on CPU (flat views, bucket
construction, backward-overlapped gloo all-reduce == mean of per-rank gradients) and, on the GPU, the fused
unscale * clip * AdamW kernels against GradScaler.unscale_ + clip_grad_norm_ + torch.optim.AdamW."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def _net(seed=0):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Linear(37, 64), nn.SiLU(), nn.Linear(64, 64), nn.SiLU(), nn.Linear(64, 5))


def test_flat_params_are_views():
    from synfmc_b200.train import FlatParams, GradAllReduce
    net = _net()
    before = [p.detach().clone() for p in net.parameters()]
    flat = FlatParams(net.parameters())
    assert all(torch.equal(a, b) for a, b in zip(before, net.parameters()))
    assert all(off % 4 == 0 for off in flat.offsets) and flat.numel >= sum(p.numel() for p in net.parameters())
    x = torch.randn(8, 37)
    net(x).square().mean().backward()
    # autograd accumulated straight into the flat buffer
    for p, off in zip(flat.params, flat.offsets):
        assert p.grad.data_ptr() == flat.grads[off:].data_ptr()
        assert torch.equal(flat.grads[off:off + p.numel()].view(p.shape), p.grad) and float(p.grad.abs().sum()) > 0
    flat.values.mul_(2.0)   # the module's parameters ARE the flat buffer
    assert all(torch.equal(2 * a, b) for a, b in zip(before, net.parameters()))
    red = GradAllReduce(flat, bucket_bytes=4096 * 4)
    covered = sorted((lo, hi) for lo, hi, _ in red.buckets)
    assert covered[0][0] == 0 and covered[-1][1] == flat.numel
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))           # a partition of the flat buffer
    assert red.buckets[0][1] == flat.numel                                    # first bucket = the LAST layers
    flat.zero_grad()
    assert float(flat.grads.abs().sum()) == 0.0 and flat.params[0].grad is not None


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from synfmc_b200.train import FlatParams, GradAllReduce
    net = _net()                                     # same weights on every rank
    flat = FlatParams(net.parameters())
    red = GradAllReduce(flat, bucket_bytes=2048 * 4).install_hooks()
    torch.manual_seed(100 + rank)
    x = torch.randn(8, 37)                           # a different "clip" per rank
    net(x).square().mean().backward()                # the hooks launch the bucket all-reduces inside backward
    n = red.wait()
    if rank == 0:
        out.put((flat.grads / n).clone())
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_mean_of_rank_gradients_gloo():
    """World size 2 (gloo): hook-launched bucket all-reduces give exactly the mean of the two ranks' gradients, i.e. the
    gradient of the concatenated batch (DDP semantics, train_cam_ctrl.py:445)."""
    world, port = 2, 29531
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from synfmc_b200.train import FlatParams
    want = None
    for r in range(world):
        net = _net()
        flat = FlatParams(net.parameters())
        torch.manual_seed(100 + r)
        net(torch.randn(8, 37)).square().mean().backward()
        want = flat.grads.clone() if want is None else want + flat.grads
    assert torch.allclose(got, want / world, rtol=0, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("device_state", [False, True])
@pytest.mark.parametrize("loss_scale,max_norm", [(1.0, 1.0), (65536.0, 1.0), (1024.0, 0.05), (1.0, 0.0)])
def test_fused_adamw_matches_torch(cuda_device, loss_scale, max_norm, device_state):
    """5 steps of unscale -> clip_grad_norm_ -> AdamW (train_cam_ctrl.py:647-655) on real autograd gradients; with
    `device_state` the bias-correction step count and the learning rate are read on the device (graph-capturable form)."""
    from synfmc_b200.train import FlatParams, FusedAdamW
    ref, net = _net(1).to(cuda_device), _net(1).to(cuda_device)
    opt_ref = torch.optim.AdamW(ref.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    flat = FlatParams(net.parameters())
    opt = FusedAdamW(flat, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=max_norm)
    g = torch.Generator().manual_seed(5)
    for step in range(5):
        x = torch.randn(16, 37, generator=g).to(cuda_device)
        (ref(x).square().mean() * loss_scale).backward()
        for p in ref.parameters():
            p.grad.div_(loss_scale)                                   # scaler.unscale_
        norm_ref = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm) if max_norm > 0 else None
        opt_ref.step()
        opt_ref.zero_grad(set_to_none=True)
        (net(x).square().mean() * loss_scale).backward()
        opt.step(loss_scale=loss_scale, device_state=device_state)
        if norm_ref is not None:
            assert abs(opt.last_norm() - float(norm_ref)) <= 1e-5 * float(norm_ref)
        opt.zero_grad()
        for a, b in zip(net.parameters(), ref.parameters()):
            assert float((a - b).abs().max()) <= (4e-6 if device_state else 2e-6) * float(b.abs().max()), step
    assert not opt.found_inf() and opt.steps_taken() == 5


@pytest.mark.gpu
def test_fused_adamw_skips_on_nonfinite_gradients(cuda_device):
    """GradScaler.step skips the optimizer when a gradient overflowed: nothing may change, the flag is raised."""
    from synfmc_b200.train import FlatParams, FusedAdamW
    net = _net(2).to(cuda_device)
    flat = FlatParams(net.parameters())
    opt = FusedAdamW(flat, lr=1e-3)
    before = flat.values.clone()
    flat.grads.normal_()
    flat.grads[12345 % flat.numel] = float("inf")
    opt.step(loss_scale=1024.0)
    assert opt.found_inf() and torch.equal(flat.values, before) and opt.steps_taken() == 0
    assert float(opt.exp_avg.abs().sum()) == 0.0 and float(opt.exp_avg_sq.abs().sum()) == 0.0
    flat.grads[12345 % flat.numel] = 0.5
    opt.step(loss_scale=1024.0)
    assert not opt.found_inf() and not torch.equal(flat.values, before)


def test_early_gradient_delivery_protocol_cpu():
    """Host logic of the tape -> reducer hand-over (train_engine.EARLY_GRAD_SINK): contribution counts are learned in the
    first backward, from the second one a gradient is added to `.grad` and its bucket released the moment its last
    contribution arrives; a structure change invalidates the learned counts; a second reducer cannot steal the sink."""
    from synfmc_b200 import engine, train_engine
    from synfmc_b200.train import FlatParams, GradAllReduce
    a, b = torch.nn.Parameter(torch.zeros(4, 8)), torch.nn.Parameter(torch.zeros(8))
    flat = FlatParams([a, b])
    red = GradAllReduce(flat, bucket_bytes=16).install_hooks()      # one bucket per parameter
    launched = []
    red._launch = lambda bkt: launched.append(bkt)
    try:
        with pytest.raises(RuntimeError):
            GradAllReduce(flat).install_hooks()
        for step in range(3):
            if step == 2:
                engine.structure_changed()                          # e.g. set_processor: counts must be re-learned
            flat.zero_grad()
            red.reset()
            del launched[:]
            tape = train_engine.Tape()
            tape.add_param_grad(a, torch.ones(4, 8))                # `a` receives two contributions per backward
            assert float(a.grad.sum()) == 0.0
            tape.add_param_grad(b, torch.full((8,), 2.0))
            tape.add_param_grad(a, torch.ones(32))
            early = step == 1
            assert (tape.param_grads[a] is None) == early and (tape.param_grads[b] is None) == early
            assert float(a.grad.sum()) == (64.0 if early else 0.0) and float(b.grad.sum()) == (16.0 if early else 0.0)
            assert sorted(launched) == ([0, 1] if early else [])
            tape.finish_params()
            assert a._fmc_grad_parts[1] == 2 and b._fmc_grad_parts[1] == 1
        # gradient accumulation: inside no_sync() gradients are delivered (early) but no bucket goes out
        flat.zero_grad()
        red.reset()
        del launched[:]
        a._fmc_grad_parts = b._fmc_grad_parts = (engine.STRUCTURE_EPOCH, 1)
        with red.no_sync():
            tape = train_engine.Tape()
            tape.add_param_grad(a, torch.ones(4, 8))
            tape.add_param_grad(b, torch.ones(8))
        assert launched == [] and float(a.grad.sum()) == 32.0
        tape = train_engine.Tape()
        tape.add_param_grad(a, torch.ones(4, 8))
        tape.add_param_grad(b, torch.ones(8))
        assert sorted(launched) == [0, 1] and float(a.grad.sum()) == 64.0 and float(b.grad.sum()) == 16.0
    finally:
        red.remove_hooks()
    assert train_engine.EARLY_GRAD_SINK is None
