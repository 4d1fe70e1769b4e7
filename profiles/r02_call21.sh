#!/bin/bash
# Round 2, GPU call 21: full GPU suite + smoke + bench on the final tree
TAG=r02y
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_gpu_suite.log; tail -3 gpurun_out/${TAG}_gpu_suite.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/${TAG}_bench.json | cut -c1-330
timeout 200 python profiles/train_step_bench.py --stage cmc --steps 5 --warmup 2 --graph > gpurun_out/${TAG}_train_cmc_graph.json 2>/dev/null; tail -1 gpurun_out/${TAG}_train_cmc_graph.json | cut -c1-300
