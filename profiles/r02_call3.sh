#!/bin/bash
# Round 2, GPU call 3: persistent GroupNorm (bit-identity + timing), fixed / new parity tests (config 4, config 5), smoke.
TAG=r02c
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -s -k "groupnorm" > gpurun_out/${TAG}_gn.log 2>&1; echo "gn tests rc=$?"; grep -E "parity|passed|failed|Error" gpurun_out/${TAG}_gn.log | tail -12
timeout 300 python profiles/norm_bench.py 20 gn > gpurun_out/${TAG}_norm_bench.txt 2>&1; tail -16 gpurun_out/${TAG}_norm_bench.txt
timeout 300 python -m pytest tests/test_gpu_precise.py -q -k "conv3x3" > gpurun_out/${TAG}_precise.log 2>&1; echo "precise conv rc=$?"; tail -2 gpurun_out/${TAG}_precise.log
timeout 2400 python -m pytest tests/test_gpu_models.py -q -s -k "multidiff or config4 or config5 or autocast or window" > gpurun_out/${TAG}_models.log 2>&1; echo "models rc=$?"; grep -E "parity|passed|failed|Error" gpurun_out/${TAG}_models.log | tail -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
for m in 4 5; do FMC_GN_FUSED=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_gn$m.json 2>/dev/null; echo "bench gn mode $m:"; cut -c1-200 gpurun_out/${TAG}_bench_gn$m.json | grep -o '"value": [0-9.]*, "unit": "steps/s", "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*'; grep -o '"fmc_groupnorm_bf16": {[^}]*}' gpurun_out/${TAG}_bench_gn$m.json; done
