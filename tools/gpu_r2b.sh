#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for cfg in "3 600" "24 60" "3 40"; do echo "== $cfg"; timeout 120 python tools/debug_mega.py $cfg 4 | grep -v "last phase"; done
echo "== decoder"; timeout 1500 python -m pytest tests/test_gpu_decoder.py -x -q -m gpu 2>&1 | tail -15
echo "== trace"; timeout 300 python tools/trace_mega.py 6 > gpurun_out/mega_trace_r02b.txt 2>&1; tail -30 gpurun_out/mega_trace_r02b.txt
echo "== bench"; timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; tail -3 gpurun_out/bench_r02b.err; cat gpurun_out/bench_r02b.json
