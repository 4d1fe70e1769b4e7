#!/bin/bash
# Round 2, GPU call 5 (2 GPUs): single-clip CFG-pair latency mode, config 5 at N = 2, gradient all-reduce + fused AdamW,
# training-tail GPU tests, bench.py --gpus 2.
TAG=r02e
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_train_tail.py -q > gpurun_out/${TAG}_train_tests.log 2>&1; echo "train tests rc=$?"; tail -3 gpurun_out/${TAG}_train_tests.log
timeout 600 $TR --master-port 29511 profiles/cfg_pair_bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_cfg_pair.json 2> gpurun_out/${TAG}_cfg_pair.err; echo "cfg pair rc=$?"; tail -1 gpurun_out/${TAG}_cfg_pair.json; tail -2 gpurun_out/${TAG}_cfg_pair.err
timeout 600 $TR --master-port 29513 profiles/allreduce_bench.py --params-m 218 > gpurun_out/${TAG}_allreduce_218.json 2> gpurun_out/${TAG}_allreduce.err; echo "allreduce rc=$?"; tail -1 gpurun_out/${TAG}_allreduce_218.json; tail -2 gpurun_out/${TAG}_allreduce.err
timeout 300 $TR --master-port 29514 profiles/allreduce_bench.py --params-m 152.5 > gpurun_out/${TAG}_allreduce_152.json 2>> gpurun_out/${TAG}_allreduce.err; tail -1 gpurun_out/${TAG}_allreduce_152.json
timeout 900 $TR --master-port 29512 profiles/cfg5_bench.py --steps 3 --warmup 1 > gpurun_out/${TAG}_cfg5_n2.json 2> gpurun_out/${TAG}_cfg5_n2.err; echo "cfg5 n2 rc=$?"; tail -1 gpurun_out/${TAG}_cfg5_n2.json; tail -2 gpurun_out/${TAG}_cfg5_n2.err
timeout 600 $TR --master-port 29515 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-260 gpurun_out/${TAG}_bench_n2.json
