#!/bin/bash
# Round 2, GPU call 10 (2 GPUs): train-step bench with the all-reduce, per-call trace of the step, full-depth gradient test
TAG=r02g
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python profiles/train_step_bench.py --stage cmc --steps 2 --warmup 1 --trace > gpurun_out/${TAG}_train_cmc_trace.json 2> gpurun_out/${TAG}_train_cmc_trace.txt; echo "trace rc=$?"; head -30 gpurun_out/${TAG}_train_cmc_trace.txt
timeout 600 $TR --master-port 29517 profiles/train_step_bench.py --stage cmc --steps 3 --warmup 1 > gpurun_out/${TAG}_train_cmc_n2.json 2> gpurun_out/${TAG}_train_cmc_n2.err; echo "cmc n2 rc=$?"; tail -1 gpurun_out/${TAG}_train_cmc_n2.json | cut -c1-700
timeout 600 $TR --master-port 29518 profiles/train_step_bench.py --stage omc --steps 3 --warmup 1 > gpurun_out/${TAG}_train_omc_n2.json 2> gpurun_out/${TAG}_train_omc_n2.err; echo "omc n2 rc=$?"; tail -1 gpurun_out/${TAG}_train_omc_n2.json | cut -c1-700
timeout 1700 python -m pytest tests/test_gpu_training.py -q -s -k full_depth 2>&1 | grep -E "parity|passed|failed|Error|assert" | cut -c1-300 | tail -8
