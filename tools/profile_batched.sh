#!/bin/bash
# ncu captures of the two kernels of the batched decode step (run on the GPU box after profile_round.sh, same tag):
# the persistent attention and the cluster split-K weight-streaming GEMM.  python tools/summarize_profiles.py <tag> picks them up.
set -u
R=${1:-r02}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:attention_stream_kernel -s 5 -c 2 -o gpurun_out/prof_attn_stream_$R -f \
    python tools/bench_batched.py --skip-linear --only-keys 7459 > gpurun_out/ncu_attn_stream_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_ws_kernel -s 30 -c 5 -o gpurun_out/prof_gemm_ws_$R -f \
    python tools/bench_batched.py --skip-attention > gpurun_out/ncu_gemm_ws_$R.log 2>&1
ls -la gpurun_out/prof_attn_stream_$R.ncu-rep gpurun_out/prof_gemm_ws_$R.ncu-rep
