#!/bin/bash
# Round 2, GPU call 2: reference-precision parity (whole file), op tests with the per-kernel 1-ulp cases, model tests with
# measured errors printed, bench (e2e after the fingerprint / text-KV changes), memory-bound probes of the new kernels.
TAG=r02b
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_precise.py -q -s > gpurun_out/${TAG}_precise.log 2>&1; echo "precise rc=$?"; grep -E "rel-L2|precision|passed|failed|Error|error" gpurun_out/${TAG}_precise.log | tail -60
timeout 600 python -m pytest tests/test_gpu_ops.py -q -s > gpurun_out/${TAG}_ops.log 2>&1; echo "ops rc=$?"; grep -E "parity|passed|failed" gpurun_out/${TAG}_ops.log | tail -30
timeout 1500 python -m pytest tests/test_gpu_models.py -q -s > gpurun_out/${TAG}_models.log 2>&1; echo "models rc=$?"; grep -E "parity|passed|failed|Error" gpurun_out/${TAG}_models.log | tail -30
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/${TAG}_bench_n1.json; grep -o '"e2e": {[^}]*}' gpurun_out/${TAG}_bench_n1.json; grep -o '"encoders_ms": [0-9.]*' gpurun_out/${TAG}_bench_n1.json
timeout 200 python profiles/kernel_probe.py --gbs plucker_cfg2 traj_cfg2_1obj traj_cfg2_3obj mask_mod_l0 > gpurun_out/${TAG}_membound_gbs.txt 2>&1; cat gpurun_out/${TAG}_membound_gbs.txt
