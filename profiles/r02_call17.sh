#!/bin/bash
# Round 2, GPU call 17: full GPU suite on the current tree + smoke + bench (both arms short)
TAG=r02n
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_gpu_suite.log; tail -6 gpurun_out/${TAG}_gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/${TAG}_bench.json | cut -c1-900
