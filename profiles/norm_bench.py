#!/usr/bin/env python
"""GroupNorm / LayerNorm / row-statistics kernels at the config-2 shapes.  GroupNorm modes (FMC_GN_FUSED, read per
call): 0 = partial + finalize + apply, 1 = single-pass cluster kernel with the slab in registers (40 / 80-channel
chunks), 3 = 1 with the TMA-staged shared-memory slab where it is 8 - 32 KB, 4 = 3 including 120-channel chunks.
`cold` rotates over enough buffers to exceed the 126 MB L2 (the producer did not just write x), `hot` reuses one buffer
(x fully L2-resident when it fits) -- inside a denoising step the truth lies in between.
usage: python profiles/norm_bench.py [reps] [gn]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synfmc_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
BF = torch.bfloat16
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 20


def time_us(fns, reps=REPS):
    """mean time of one launch: the closures of `fns` are captured round-robin into ONE CUDA graph (so the Python /
    ctypes launch path, ~12 us per call, is out of the picture -- as in the graph-replayed denoising step) and the
    graph is replayed `reps` times between two events"""
    for f in fns:
        f()
    torch.cuda.synchronize()
    inner = max(1, 16 // len(fns))
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(inner):
                for f in fns:
                    f()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * inner * len(fns))


def buffers(rows, C, n):
    return [(torch.randn(rows, C, device=dev) * 0.7 + 0.1).to(BF) for _ in range(n)]


def gn_row(images, HW, C, silu=True):
    rows = images * HW
    mb = rows * C * 2 / 1e6
    ncold = max(2, int(300 / mb) + 1)
    xs = buffers(rows, C, ncold)
    outs = [torch.empty_like(x) for x in xs]
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    cols = []
    best = 1e30
    for mode in ("0", "1", "3", "4", "5"):
        os.environ["FMC_GN_FUSED"] = mode
        cold = time_us([lambda x=x, o=o: ops.groupnorm(x, g, b, 1e-6, images, HW, silu=silu, out=o) for x, o in zip(xs, outs)])
        hot = time_us([lambda: ops.groupnorm(xs[0], g, b, 1e-6, images, HW, silu=silu, out=outs[0])])
        best = min(best, cold)
        cols.append(f"mode {mode} cold {cold:6.1f} hot {hot:6.1f}")
    print(f"groupnorm  images {images:3d} HW {HW:5d} C {C:5d}  {2 * mb:6.1f} MB r+w | " + " | ".join(cols) +
          f" us | best cold {2 * mb / best:5.2f} TB/s", flush=True)


def ln_row(rows, C, F=16, HW=0, add=False):
    mb = rows * C * 2 / 1e6
    ncold = max(2, int(300 / mb) + 1)
    xs = buffers(rows, C, ncold)
    adds = buffers(rows, C, ncold) if add else [None] * ncold
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    pe = torch.randn(32, C, device=dev) if HW else None
    traffic = mb * (4 if add else 2)
    for r in (("2", "1") if add else ("2",)):
        os.environ["FMC_LN_ADD_R"] = r
        cold = time_us([lambda x=x, a=a: ops.layernorm(x, g, b, 1e-5, pe=pe, F=F if HW else 0, HW=HW, add=a) for x, a in zip(xs, adds)])
        hot = time_us([lambda: ops.layernorm(xs[0], g, b, 1e-5, pe=pe, F=F if HW else 0, HW=HW, add=adds[0])])
        print(f"layernorm  rows {rows:6d} C {C:5d} pe {int(bool(HW))} add {int(add)} R {r}  {traffic:7.1f} MB | cold {cold:7.1f} hot {hot:7.1f} us"
              f" | cold {traffic / cold:5.2f} TB/s", flush=True)
    os.environ.pop("FMC_LN_ADD_R", None)
    if add:
        return
    cold = time_us([lambda x=x: ops.rowstats(x) for x in xs])
    hot = time_us([lambda: ops.rowstats(xs[0])])
    print(f"rowstats   rows {rows:6d} C {C:5d}               {mb:7.1f} MB | cold {cold:7.1f} hot {hot:7.1f} us"
          f" | cold {mb / cold:5.2f} TB/s", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "reps", REPS)
    for images, HW, C in [(32, 2560, 320), (32, 2560, 640), (32, 2560, 960), (32, 640, 640), (32, 640, 1280),
                          (32, 640, 1920), (32, 640, 960), (32, 640, 320), (32, 160, 1280), (32, 160, 2560),
                          (32, 160, 1920), (32, 160, 640), (32, 40, 1280), (32, 40, 2560)]:
        try:
            gn_row(images, HW, C)
        except Exception as e:  # keep the table going
            print("groupnorm", images, HW, C, "failed:", repr(e)[:200], flush=True)
    os.environ.pop("FMC_GN_FUSED", None)
    if len(sys.argv) > 2 and sys.argv[2] == "gn":
        sys.exit(0)
    for rows, C, HW in [(81920, 320, 2560), (20480, 640, 640), (5120, 1280, 160), (1280, 1280, 40)]:
        ln_row(rows, C)
        ln_row(rows, C, HW=HW, add=True)
