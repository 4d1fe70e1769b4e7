#!/bin/bash
# Round 2, GPU call 1: reference-precision parity first, then the whole GPU suite, bench, memory-bound GB/s + ncu.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/r02_call1.sh'
TAG=r02a
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_precise.py -x -q -s > gpurun_out/${TAG}_precise.log 2>&1; echo "precise rc=$?"; tail -25 gpurun_out/${TAG}_precise.log
FMC_TEST_UNVERIFIED=1 timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_precise.py > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/${TAG}_tests.log
FMC_BENCH_TRACE=gpurun_out/${TAG}_trace_shapes.txt timeout 400 python bench.py --steps 10 --warmup 3 \
    > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
MEM="plucker_cfg2 traj_cfg2_1obj traj_cfg2_3obj mask_mod_l0 mask_mod_l1 mask_mod_l2 mask_mod_l3 groupnorm_l0 groupnorm_l1 groupnorm_l0_cat layernorm_l0 layernorm_l0_pose layernorm_l1_pose rowstats_l0 add_l0"
timeout 200 python profiles/kernel_probe.py --gbs $MEM > gpurun_out/${TAG}_membound_gbs.txt 2>&1; cat gpurun_out/${TAG}_membound_gbs.txt
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:plucker|traj_|mask_modulate|groupnorm|layernorm|rowstats|add_kernel" -c 15 -f \
    -o gpurun_out/${TAG}_prof_membound python profiles/kernel_probe.py --once $MEM > gpurun_out/${TAG}_prof_membound.log 2>&1; echo "ncu membound rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "launch list rc=$?"
