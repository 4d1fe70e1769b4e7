#!/bin/bash
# Round 2, GPU call 13: vectorised GroupNorm backward; torch.profiler kernel list of the training step
TAG=r02j
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_backward.py -q -x 2>&1 | tail -5 > gpurun_out/${TAG}_bwd_tests.log; tail -3 gpurun_out/${TAG}_bwd_tests.log
timeout 600 python -m pytest tests/test_gpu_training.py -q -x -k "not full_depth" 2>&1 | tail -8 > gpurun_out/${TAG}_training_tests.log; tail -2 gpurun_out/${TAG}_training_tests.log
timeout 600 python profiles/train_step_bench.py --stage cmc --steps 3 --warmup 1 --trace --profile > gpurun_out/${TAG}_train_cmc_trace.json 2> gpurun_out/${TAG}_train_cmc_trace.txt; echo "trace rc=$?"; grep -v "^ *0\.[0-4]" gpurun_out/${TAG}_train_cmc_trace.txt | head -120; tail -1 gpurun_out/${TAG}_train_cmc_trace.json | cut -c1-400
