#!/bin/bash
# Round 2, GPU call 19: backward-kernel probe (us per call at config-3 shapes) + one ncu --set full capture of them
TAG=r02s
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 300 python profiles/bwd_probe.py > gpurun_out/${TAG}_bwd_probe.txt 2> gpurun_out/${TAG}_bwd_probe.err; echo "probe rc=$?"; cat gpurun_out/${TAG}_bwd_probe.txt; tail -3 gpurun_out/${TAG}_bwd_probe.err
timeout 900 ncu --set full --clock-control none -k regex:'attn_bwd|temporal16|groupnorm_bwd|layernorm_bwd' -o gpurun_out/${TAG}_bwd python profiles/bwd_probe.py --once > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/${TAG}_bwd.ncu-rep
