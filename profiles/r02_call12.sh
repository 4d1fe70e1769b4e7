#!/bin/bash
# Round 2, GPU call 12: attention backward on tensor cores at head_dim 80 + 16-frame temporal backward kernel
TAG=r02i
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_backward.py -q -x -k "attention_bwd" 2>&1 | tail -15 > gpurun_out/${TAG}_attn_bwd_tests.log; echo "attn tests rc=$?"; tail -12 gpurun_out/${TAG}_attn_bwd_tests.log
timeout 600 python -m pytest tests/test_gpu_training.py -q -x -k "not full_depth" 2>&1 | tail -8 > gpurun_out/${TAG}_training_tests.log; tail -5 gpurun_out/${TAG}_training_tests.log
timeout 600 python profiles/train_step_bench.py --stage cmc --steps 3 --warmup 1 --trace > gpurun_out/${TAG}_train_cmc_trace.json 2> gpurun_out/${TAG}_train_cmc_trace.txt; echo "trace rc=$?"; head -70 gpurun_out/${TAG}_train_cmc_trace.txt; tail -1 gpurun_out/${TAG}_train_cmc_trace.json | cut -c1-600
