#!/bin/bash
# Round 2, GPU call 6 (8 GPUs): gradient all-reduce + fused AdamW at config-3 / config-4 sizes, config 5 batch-shard, CFG pairs.
TAG=r02f
export PYTHONUNBUFFERED=1
export NCCL_DEBUG=WARN
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NCCL_DEBUG=INFO timeout 300 $TR --master-port 29513 profiles/allreduce_bench.py --params-m 218 > gpurun_out/${TAG}_allreduce_218.json 2> gpurun_out/${TAG}_allreduce.err; echo "allreduce rc=$?"; tail -1 gpurun_out/${TAG}_allreduce_218.json; grep -m3 -i "nvls\|NVLS" gpurun_out/${TAG}_allreduce.err | cut -c1-200
timeout 300 $TR --master-port 29514 profiles/allreduce_bench.py --params-m 152.5 > gpurun_out/${TAG}_allreduce_152.json 2>> gpurun_out/${TAG}_allreduce.err; tail -1 gpurun_out/${TAG}_allreduce_152.json
timeout 300 $TR --master-port 29516 profiles/allreduce_bench.py --params-m 218 --bucket-mb 256 > gpurun_out/${TAG}_allreduce_218_b256.json 2>> gpurun_out/${TAG}_allreduce.err; tail -1 gpurun_out/${TAG}_allreduce_218_b256.json
timeout 900 $TR --master-port 29512 profiles/cfg5_bench.py --steps 3 --warmup 1 > gpurun_out/${TAG}_cfg5.json 2> gpurun_out/${TAG}_cfg5.err; echo "cfg5 rc=$?"; tail -1 gpurun_out/${TAG}_cfg5.json; tail -2 gpurun_out/${TAG}_cfg5.err
timeout 600 $TR --master-port 29511 profiles/cfg_pair_bench.py --steps 10 --warmup 3 --no-check > gpurun_out/${TAG}_cfg_pair.json 2> gpurun_out/${TAG}_cfg_pair.err; echo "cfg pair rc=$?"; tail -1 gpurun_out/${TAG}_cfg_pair.json
