#!/bin/bash
# Round 2, GPU call 14: whole training step as a CUDA graph; norm backward tuning
TAG=r02k
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_backward.py tests/test_train_tail.py -q -x 2>&1 | tail -5 > gpurun_out/${TAG}_bwd_tests.log; tail -3 gpurun_out/${TAG}_bwd_tests.log
timeout 900 python -m pytest tests/test_gpu_training.py -q -x -k "not full_depth" 2>&1 | tail -15 > gpurun_out/${TAG}_training_tests.log; tail -12 gpurun_out/${TAG}_training_tests.log
timeout 600 python profiles/train_step_bench.py --stage cmc --steps 5 --warmup 2 > gpurun_out/${TAG}_train_cmc_eager.json 2> gpurun_out/${TAG}_train_cmc_eager.err; echo "eager rc=$?"; tail -1 gpurun_out/${TAG}_train_cmc_eager.json | cut -c1-400
timeout 600 python profiles/train_step_bench.py --stage cmc --steps 5 --warmup 2 --graph > gpurun_out/${TAG}_train_cmc_graph.json 2> gpurun_out/${TAG}_train_cmc_graph.err; echo "graph rc=$?"; tail -5 gpurun_out/${TAG}_train_cmc_graph.err; tail -1 gpurun_out/${TAG}_train_cmc_graph.json | cut -c1-600
timeout 600 python profiles/train_step_bench.py --stage omc --steps 5 --warmup 2 --graph > gpurun_out/${TAG}_train_omc_graph.json 2> gpurun_out/${TAG}_train_omc_graph.err; echo "omc graph rc=$?"; tail -5 gpurun_out/${TAG}_train_omc_graph.err; tail -1 gpurun_out/${TAG}_train_omc_graph.json | cut -c1-600
