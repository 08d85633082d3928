#!/bin/bash
# ncu passes for one round (run on the GPU box: gpurun --timeout 2400 -- 'bash tools/profile_round.sh r02').
# 1. launch list of a short bench run (per-launch durations, cold cache, serialised): the kernels' SHARES of a step;
# 2. one full capture of the persistent decode kernel and of each tcgen05 kernel (source page needs -lineinfo: on).
# Numbers printed by bench.py under ncu are NOT bench values.  Summaries: python tools/summarize_profiles.py r02.
set -u
R=${1:-r02}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --faces 16 --steps 1 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/launches_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_mega_kernel -c 1 -o gpurun_out/prof_mega_$R -f \
    python bench.py --faces 16 --steps 1 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_mega_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -c 3 -o gpurun_out/prof_attn_tc_$R -f \
    python tools/bench_encoder.py --batch 8 --faces 800 > gpurun_out/ncu_attn_tc_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 3 -o gpurun_out/prof_gemm_tc_$R -f \
    python tools/bench_encoder.py --batch 8 --faces 800 > gpurun_out/ncu_gemm_tc_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_ws_kernel -s 30 -c 5 -o gpurun_out/prof_gemm_ws_$R -f \
    python tools/bench_batched.py --skip-attention > gpurun_out/ncu_gemm_ws_$R.log 2>&1
ls -la gpurun_out/*$R*.ncu-rep
