"""Training-step tail for the FMC trainers: what follows `loss.backward()` in
train_cam_ctrl.py:647-665 / train_cam_obj_ctrl.py:843-862 -- DDP's gradient all-reduce, `scaler.unscale_`,
`clip_grad_norm_`, `AdamW.step`, `zero_grad` -- on ONE flat fp32 buffer per role (values, gradients, two Adam moments):

  * `FlatParams`      re-homes the trainable parameters (CMC: CameraEncoder + 20 qkv_merge = 218 M; OMC: the ObjectEncoder,
                      152.5 M) into a flat buffer and makes every `p.grad` a view of a flat gradient buffer, so autograd
                      accumulates straight into it;
  * `GradAllReduce`   buckets the flat gradient (default 64 MiB: sized for launch latency and overlap, not link count --
                      NVSwitch gives every GPU full bandwidth to every peer) and all-reduces each bucket with NCCL as
                      soon as the last gradient of the bucket has been produced (post-accumulate-grad hooks), i.e.
                      overlapped with the rest of backward; the mean's 1 / world is folded into the unscale factor;
  * `FusedAdamW`      fmc_grad_norm_f32 + fmc_adamw_step_f32: one read of the gradients for the global norm, one fused
                      unscale * clip * AdamW pass; the clip coefficient and the found-inf flag never leave the device.

  * `GraphedStep`     captures one WHOLE training step (zero_grad, forward on the tape of train_engine.py, loss, backward,
                      bucket all-reduces, optimizer) into a CUDA graph: the step is ~4000 kernels of 5-50 us, more than
                      Python can launch in the time the GPU needs for them.

The gradients come from the backward kernels behind train_engine.py or from any other source (the
tests also feed these classes torch-autograd gradients of small modules and compare with DDP-style mean + AdamW)."""
import contextlib
import os

import torch
import torch.distributed as dist

from . import _cabi, ops


class FlatParams:
    def __init__(self, params, device=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        device = torch.device(device) if device is not None else self.params[0].device
        self.offsets, total = [], 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise TypeError("FlatParams keeps fp32 master parameters (the reference trains fp32 weights under autocast)")
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4  # every parameter starts 16-byte aligned (float4 access in the kernels)
        self.numel = total
        self.values = torch.zeros(total, device=device, dtype=torch.float32)
        self.grads = torch.zeros(total, device=device, dtype=torch.float32)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = self.values[off:off + p.numel()].view(p.shape)
                view.copy_(p.detach())
                p.data = view
        self.attach_grads()

    def attach_grads(self):
        """(Re-)point every `p.grad` at its slice of the flat gradient buffer (after a `zero_grad(set_to_none=True)`)."""
        for p, off in zip(self.params, self.offsets):
            p.grad = self.grads[off:off + p.numel()].view(p.shape)

    def zero_grad(self):
        self.grads.zero_()
        self.attach_grads()


class GradAllReduce:
    """Bucketed, backward-overlapped all-reduce (SUM; the division by world size happens in FusedAdamW's unscale factor)."""

    def __init__(self, flat, bucket_bytes=64 << 20, group=None):
        self.flat, self.group = flat, group
        per = max(4, bucket_bytes // 4 // 4 * 4)
        # buckets are ranges of the flat buffer cut at parameter boundaries, filled from the END of the parameter list:
        # backward produces the gradients of the last layers first
        self.buckets, start = [], flat.numel
        idx_end = len(flat.params)
        i = len(flat.params) - 1
        while i >= 0:
            if start - flat.offsets[i] >= per or i == 0:
                self.buckets.append((flat.offsets[i], start, range(i, idx_end)))
                start, idx_end = flat.offsets[i], i
            i -= 1
        self.bucket_of = {}
        for b, (_, _, idxs) in enumerate(self.buckets):
            for j in idxs:
                self.bucket_of[j] = b
        self.pending, self.handles, self.hooks = [], [], []
        self.sync = True  # False inside no_sync(): gradients accumulate locally, no bucket is released
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1

    def install_hooks(self):
        """Launch a bucket's all-reduce from inside backward, when its last gradient has been accumulated: through torch's
        post-accumulate-grad hooks for gradients autograd delivers, and through the tape's early delivery
        (train_engine.EARLY_GRAD_SINK) for the parameters of the B200 training path -- their gradients are complete long
        before the tape's single autograd node returns, and only this way does the all-reduce overlap the backward."""
        from . import train_engine
        if train_engine.EARLY_GRAD_SINK is not None and getattr(train_engine.EARLY_GRAD_SINK, "owner", None) is not self:
            raise RuntimeError("another GradAllReduce has its hooks installed: call its remove_hooks() first (the tape "
                               "hands gradients to one reducer)")
        self.reset()
        index = {id(p): j for j, p in enumerate(self.flat.params)}
        for j, p in enumerate(self.flat.params):
            self.hooks.append(p.register_post_accumulate_grad_hook(lambda _p, j=j: self._ready(j)))

        def early(param, grad):
            j = index.get(id(param))
            if j is None or param.grad is None:
                return False
            param.grad.add_(grad.to(param.grad.dtype))
            self._ready(j)
            return True
        early.owner = self
        # FMC_NO_EARLY_GRADS=1: A/B switch -- every gradient reaches the reducer through autograd at the end of the tape
        train_engine.EARLY_GRAD_SINK = None if os.environ.get("FMC_NO_EARLY_GRADS") == "1" else early
        return self

    def remove_hooks(self):
        from . import train_engine
        for h in self.hooks:
            h.remove()
        self.hooks = []
        train_engine.EARLY_GRAD_SINK = None

    def reset(self):
        self.pending = [len(idxs) for _, _, idxs in self.buckets]
        self.handles = []

    @contextlib.contextmanager
    def no_sync(self):
        """DDP.no_sync for gradient accumulation: backward passes inside the block only accumulate into `.grad`; the
        all-reduce happens in the first backward after it (call `reset()` before that one as usual)."""
        self.sync = False
        try:
            yield self
        finally:
            self.sync = True

    def _ready(self, j):
        if not self.sync:
            return
        b = self.bucket_of[j]
        self.pending[b] -= 1
        if self.pending[b] == 0:
            self._launch(b)

    def _launch(self, b):
        lo, hi, _ = self.buckets[b]
        if self.world > 1:
            self.handles.append(dist.all_reduce(self.flat.grads[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def reduce_all(self):
        """Without hooks: all buckets now (last layers first)."""
        self.reset()
        for b in range(len(self.buckets)):
            self._launch(b)
        return self.wait()

    def wait(self):
        for h in self.handles:
            h.wait()
        self.reset()
        return self.world


class FusedAdamW:
    """torch.optim.AdamW(lr, betas, eps, weight_decay) + GradScaler.unscale_ + clip_grad_norm_ on a FlatParams, as two
    kernels (train_cam_ctrl.py:321-327, :647-655).  `step()` never synchronises; `last_norm()` / `found_inf()` read the
    device-side state when the caller wants them (GradScaler.update needs found_inf once per step)."""

    def __init__(self, flat, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=1.0):
        self.flat, self.lr, self.betas, self.eps, self.weight_decay = flat, lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        dev = flat.values.device
        self.exp_avg = torch.zeros_like(flat.values)
        self.exp_avg_sq = torch.zeros_like(flat.values)
        # norm, coefficient, found_inf, steps taken (counted on the device), learning rate (device-side copy), 3 spare
        self.state = torch.zeros(8, device=dev, dtype=torch.float32)
        self.state[4] = lr
        self.workspace = torch.zeros(_cabi.lib().fmc_grad_norm_workspace_floats(), device=dev, dtype=torch.float32)
        self.steps = 0

    def set_lr(self, lr):
        """Learning rate for the following steps; also updates the device-side copy a captured step reads."""
        self.lr = lr
        self.state[4:5].fill_(lr)

    def step(self, loss_scale=1.0, world=1, lr=None, device_state=False):
        """One optimizer step on the gradients in `flat.grads` = SUM over `world` ranks of `loss_scale` * dL/dp.
        `device_state`: bias-correction step count and learning rate are read from the device (state[3], state[4]) instead
        of being passed as launch arguments -- what a captured step (GraphedStep) needs, since a graph freezes arguments."""
        ops._check_cuda(self.flat.values, self.flat.grads)
        f = self.flat
        self.steps += 1
        from . import engine
        engine.OPTIMIZER_EPOCH += 1  # inference plans derived from these parameters are stale from here on
        stream = ops._stream()
        _cabi.call("fmc_grad_norm_f32", f.grads.data_ptr(), f.numel, 1.0 / (float(loss_scale) * world),
                   float(self.max_grad_norm or 0.0), self.workspace.data_ptr(), self.state.data_ptr(), stream)
        _cabi.call("fmc_adamw_step_f32", f.values.data_ptr(), f.grads.data_ptr(), self.exp_avg.data_ptr(),
                   self.exp_avg_sq.data_ptr(), f.numel, -1.0 if device_state else float(self.lr if lr is None else lr),
                   float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                   0 if device_state else self.steps, self.state.data_ptr(), stream)

    def zero_grad(self, set_to_none=False):
        self.flat.zero_grad()

    def steps_taken(self):
        """Optimizer steps that were not skipped for inf / nan gradients, as counted on the device."""
        return int(self.state[3])

    def last_norm(self):
        return float(self.state[0])

    def found_inf(self):
        return bool(self.state[2] != 0)


class GraphedStep:
    """One whole training step as a CUDA graph.

    `fn()` runs the step -- `opt.zero_grad(); red.reset(); pred = wrapper(...); loss = ...; loss.backward(); n = red.wait();
    opt.step(world=n, device_state=True)` -- and returns the tensors to keep (the loss).  It must read its inputs from
    tensors whose storage does not change (refill them with `copy_` between replays), must not read device values on
    the host, and must call the optimizer with `device_state=True` (step count and learning rate live on the device;
    `FusedAdamW.set_lr` changes the rate without re-capturing).  `warmup` eager steps run first on a side stream -- they
    build every plan and workspace and ARE training steps -- then the step is captured; every call replays it.  The NCCL
    all-reduces launched by the gradient hooks during capture are part of the graph (torch's process group captures)."""

    def __init__(self, fn, warmup=3):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedStep captures a CUDA graph: no CUDA device (there is no CPU path)")
        cur = torch.cuda.current_stream()
        # warm-up and capture run on ONE side stream: autograd remembers the stream each AccumulateGrad node was created on
        # (the first warm-up step) and synchronises with it in every later backward -- a capture on a different stream would
        # then depend on uncaptured work (cudaErrorStreamCaptureIsolation)
        self.stream = torch.cuda.Stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                fn()
        cur.wait_stream(self.stream)
        torch.cuda.synchronize()
        self.warmup_steps = warmup
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.out = fn()

    def __call__(self):
        from . import engine
        self.graph.replay()
        engine.OPTIMIZER_EPOCH += 1  # the captured optimizer moved the parameters: inference plans derived from them are stale
        return self.out
