#!/bin/bash
# Round 2, GPU call 4: sphere-mask kernels, config-5 bench at N = 1, full GPU suite (state of the tree), bench.
TAG=r02d
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -k "sphere or traj or plucker" > gpurun_out/${TAG}_ops.log 2>&1; echo "ops rc=$?"; tail -3 gpurun_out/${TAG}_ops.log
timeout 900 python profiles/cfg5_bench.py --steps 3 --warmup 1 > gpurun_out/${TAG}_cfg5_n1.json 2> gpurun_out/${TAG}_cfg5_n1.err; echo "cfg5 rc=$?"; cat gpurun_out/${TAG}_cfg5_n1.json; tail -3 gpurun_out/${TAG}_cfg5_n1.err
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
timeout 200 python profiles/kernel_probe.py --gbs traj_cfg2_3obj > gpurun_out/${TAG}_gbs.txt 2>&1; tail -2 gpurun_out/${TAG}_gbs.txt
