#!/bin/bash
# Round 2, GPU call 16: pipeline edges (VAE, CLIP text) parity + timing; OMC graphed training step
TAG=r02m
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edges.py -q -s 2>&1 | grep -E "parity|passed|failed|Error|error|assert|FAILED" | cut -c1-300 | tail -30 | tee gpurun_out/${TAG}_edges_tests.log
timeout 600 python profiles/edges_bench.py --trace --cpu > gpurun_out/${TAG}_edges_bench.json 2> gpurun_out/${TAG}_edges_bench.txt; echo "edges bench rc=$?"; head -20 gpurun_out/${TAG}_edges_bench.txt; tail -1 gpurun_out/${TAG}_edges_bench.json
timeout 300 python profiles/train_step_bench.py --stage omc --steps 5 --warmup 2 --graph > gpurun_out/${TAG}_train_omc_graph.json 2> gpurun_out/${TAG}_train_omc_graph.err; echo "omc graph rc=$?"; tail -3 gpurun_out/${TAG}_train_omc_graph.err | cut -c1-300; tail -1 gpurun_out/${TAG}_train_omc_graph.json | cut -c1-300
