#!/bin/bash
# One gpurun call that refreshes every piece of evidence of the current build (about 8 GPU-minutes on one B200):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/first_call.sh r02'
# writes gpurun_out/<tag>_*; copy what should be judged into profiles/ (python profiles/ncu_summary.py <rep> > profiles/<tag>_ncu_<case>.txt,
# python profiles/summarise_launches.py gpurun_out/<tag>_launches.csv > profiles/<tag>_launches_bench.txt).
TAG=${1:-r02}
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
# 1. parity (FMC_TEST_UNVERIFIED=1 also runs the tests written after round 1's GPU minutes were spent)
FMC_TEST_UNVERIFIED=1 timeout 400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
# 2. the bench line (with the CPU leg), the per-shape trace behind its kernel table
FMC_BENCH_TRACE=gpurun_out/${TAG}_trace_shapes.txt timeout 240 python bench.py --steps 10 --warmup 3 \
    > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_n1.json
# 3. norm kernels per shape (GroupNorm modes, LayerNorm, row statistics)
timeout 150 python profiles/norm_bench.py 20 > gpurun_out/${TAG}_norm_bench.txt 2>&1; tail -32 gpurun_out/${TAG}_norm_bench.txt
# 4. launch list of two steps under ncu (serialised, cold cache: shares, not absolute times)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "launch list rc=$?"
# 5. full-set captures of the kernels that carry the step (one launch each)
cap() {  # case, kernel regex
  timeout 120 ncu --set full --clock-control none --import-source on -k "regex:$2" -s 2 -c 1 -f \
      -o gpurun_out/${TAG}_prof_$1 python profiles/kernel_probe.py $1 4 > gpurun_out/${TAG}_prof_$1.log 2>&1; echo "ncu $1 rc=$?"
}
cap spatial_l0 spatial_attn
cap gemm_l0_out gemm_bf16_tma
cap gemm_l2_geglu gemm_bf16_tma
cap temporal_fused_l0 temporal_qkv_attn
cap groupnorm_l0 groupnorm
cap layernorm_l0 layernorm
cap cross_l0 cross_attn
