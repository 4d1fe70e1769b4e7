#!/usr/bin/env python
"""Pipeline timeline of CTA 0 of the fused temporal kernel (fmc_debug_set_timeline): cycles relative to the first event.
usage: python profiles/timeline_probe.py > gpurun_out/timeline.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synfmc_b200 import _cabi, ops  # noqa: E402

dev = torch.device("cuda", 0)
B, F, HW = 2, 16, 2560
x = torch.randn(B * F * HW, 320, device=dev).bfloat16()
w = (torch.randn(8 * 128, 320, device=dev) * 320 ** -0.5).bfloat16()
out = torch.empty(B * F * HW, 320, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.temporal_qkv_attn(x, w, out, B, F, HW, 8, 40 ** -0.5)
buf = torch.zeros(4, 64, 8, device=dev, dtype=torch.int64)
_cabi.call("fmc_debug_set_timeline", buf.data_ptr())
ops.temporal_qkv_attn(x, w, out, B, F, HW, 8, 40 ** -0.5)
torch.cuda.synchronize()
_cabi.call("fmc_debug_set_timeline", 0)
t = buf.cpu()
t0 = int(t[t > 0].min())
names = {0: "MMA  g_begin g_issued s_begin s_issued pv_begin pv_issued - -",
         1: "WG-A begin g_full_ok qk_arrived s_full_ok p_arrived - - -",
         2: "WG-B begin g_full_ok o_full_ok v_arrived epi_done - - -",
         3: "TMA  a_issue w_issue[kb 0..4] - -"}
for role in range(4):
    print(names[role])
    for m in range(24):
        row = [int(v) - t0 if v > 0 else -1 for v in t[role, m]]
        print(f"  m={m:2d} " + " ".join(f"{v:7d}" for v in row))
