#!/bin/bash
# Round 2, GPU call 20: full GPU suite on the final tree, smoke, bench, training step (eager / graph, CMC / OMC), edges
TAG=r02u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_gpu_suite.log; tail -4 gpurun_out/${TAG}_gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/${TAG}_bench.json | cut -c1-700
for st in cmc omc; do
  timeout 300 python profiles/train_step_bench.py --stage $st --steps 5 --warmup 2 --graph > gpurun_out/${TAG}_train_${st}_graph.json 2>/dev/null; tail -1 gpurun_out/${TAG}_train_${st}_graph.json | cut -c1-330
done
timeout 300 python profiles/train_step_bench.py --stage cmc --steps 5 --warmup 2 > gpurun_out/${TAG}_train_cmc_eager.json 2>/dev/null; tail -1 gpurun_out/${TAG}_train_cmc_eager.json | cut -c1-330
timeout 300 python profiles/edges_bench.py > gpurun_out/${TAG}_edges_bench.json 2>/dev/null; tail -1 gpurun_out/${TAG}_edges_bench.json
timeout 200 python profiles/bwd_probe.py > gpurun_out/${TAG}_bwd_probe.txt 2>/dev/null; cat gpurun_out/${TAG}_bwd_probe.txt
