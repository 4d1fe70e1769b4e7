#!/bin/bash
# Round 2, GPU call 18 (8 GPUs): BASELINE configs 3 / 4 as defined -- one clip per GPU x 8, the whole training step as a CUDA
# graph with the NCCL gradient all-reduce inside it
TAG=r02w
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29521 profiles/train_step_bench.py --stage cmc --steps 5 --warmup 2 --graph > gpurun_out/${TAG}_train_cmc_n8_graph.json 2> gpurun_out/${TAG}_train_cmc_n8_graph.err; echo "cmc n8 rc=$?"; tail -1 gpurun_out/${TAG}_train_cmc_n8_graph.json | cut -c1-400
timeout 240 $TR --master-port 29522 profiles/train_step_bench.py --stage omc --steps 5 --warmup 2 --graph > gpurun_out/${TAG}_train_omc_n8_graph.json 2> gpurun_out/${TAG}_train_omc_n8_graph.err; echo "omc n8 rc=$?"; tail -1 gpurun_out/${TAG}_train_omc_n8_graph.json | cut -c1-400
