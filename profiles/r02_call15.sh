#!/bin/bash
# Round 2, GPU call 15 (2 GPUs): graphed training step test, OMC graphed, 2-GPU eager / graphed with NCCL all-reduce in the graph
TAG=r02l
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_training.py -q -x -s -k "graphed" 2>&1 | grep -E "parity|passed|failed|Error|assert" | cut -c1-400 | tail -8
timeout 600 python profiles/train_step_bench.py --stage omc --steps 5 --warmup 2 > gpurun_out/${TAG}_train_omc_eager.json 2> gpurun_out/${TAG}_train_omc_eager.err; echo "omc eager rc=$?"; tail -1 gpurun_out/${TAG}_train_omc_eager.json | cut -c1-300
timeout 600 python profiles/train_step_bench.py --stage omc --steps 5 --warmup 2 --graph > gpurun_out/${TAG}_train_omc_graph.json 2> gpurun_out/${TAG}_train_omc_graph.err; echo "omc graph rc=$?"; tail -3 gpurun_out/${TAG}_train_omc_graph.err; tail -1 gpurun_out/${TAG}_train_omc_graph.json | cut -c1-300
timeout 600 $TR --master-port 29517 profiles/train_step_bench.py --stage cmc --steps 5 --warmup 2 > gpurun_out/${TAG}_train_cmc_n2.json 2> gpurun_out/${TAG}_train_cmc_n2.err; echo "cmc n2 rc=$?"; tail -1 gpurun_out/${TAG}_train_cmc_n2.json | cut -c1-300
timeout 600 $TR --master-port 29518 profiles/train_step_bench.py --stage cmc --steps 5 --warmup 2 --graph > gpurun_out/${TAG}_train_cmc_n2_graph.json 2> gpurun_out/${TAG}_train_cmc_n2_graph.err; echo "cmc n2 graph rc=$?"; tail -4 gpurun_out/${TAG}_train_cmc_n2_graph.err | cut -c1-300; tail -1 gpurun_out/${TAG}_train_cmc_n2_graph.json | cut -c1-300
