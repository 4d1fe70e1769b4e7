"""Diagnostic: python call sites of the torch copy / fill / add kernels inside one training step (torch.profiler with stacks)."""
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'profiles'))
sys.argv = ['x', '--stage', 'cmc', '--steps', '1', '--warmup', '2']
from torch.profiler import ProfilerActivity, profile
# monkeypatch: capture the step function by running main with a hook
orig_profile = None
import json
src = open('profiles/train_step_bench.py').read()
src = src.replace("    if args.profile and rank == 0:", "    if rank == 0:\n        globals()['_STEP'] = step\n    if args.profile and rank == 0:")
ns = {'__name__': 'probe', '__file__': os.path.join(os.getcwd(), 'profiles', 'train_step_bench.py')}
exec(compile(src, 'tsb', 'exec'), ns)
ns['main']()
step = ns['_STEP']
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
agg = {}
for e in prof.events():
    if e.name in ('aten::copy_', 'aten::contiguous', 'aten::clone', 'aten::fill_', 'aten::zero_', 'aten::add', 'aten::add_') and e.device_time_total > 0:
        st = [s for s in (e.stack or []) if 'synfmc_b200' in s or 'profiles' in s]
        key = (e.name, (st[0] if st else '') + ' shapes=' + str(e.input_shapes)[:90])
        d = agg.setdefault(key, [0, 0.0]); d[0] += 1; d[1] += e.device_time_total / 1e3
for (name, where), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{ms:8.2f} ms {n:5d} x {name:18s} {where[:140]}")
