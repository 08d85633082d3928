#!/bin/bash
# round-2 first GPU pass: unit parity, decoder parity, trace, quick bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== ops"; timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "linear_bit or segmented or layernorm or attention_bit" 2>&1 | tail -15
echo "== decoder"; timeout 1200 python -m pytest tests/test_gpu_decoder.py -x -q -m gpu 2>&1 | tail -40
echo "== trace"; timeout 300 python tools/trace_mega.py 6 > gpurun_out/mega_trace_r02a.txt 2>&1; tail -40 gpurun_out/mega_trace_r02a.txt
echo "== bench"; timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; tail -3 gpurun_out/bench_r02a.err; cat gpurun_out/bench_r02a.json
