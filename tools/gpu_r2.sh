#!/bin/bash
# GPU pass of round 2: parity suites, phase trace, short bench.   usage: gpurun -- bash tools/gpu_r2.sh [tag]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-x}
mkdir -p gpurun_out
echo "== ops"; timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "linear_bit or segmented or layernorm or attention_bit" 2>&1 | tail -4
echo "== decoder"; timeout 1500 python -m pytest tests/test_gpu_decoder.py -x -q -m gpu 2>&1 | tail -15
echo "== trace short"; timeout 300 python tools/trace_mega.py 30 > gpurun_out/mega_trace_${TAG}.txt 2>&1; tail -24 gpurun_out/mega_trace_${TAG}.txt
echo "== bench"; timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err; python - <<P
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
r=d["roofline"]
print("value", d["value"], "frac", r["frac"], "us/step", r["us_per_step_avg"], "short", r["short_context"]["us_per_step"], r["short_context"]["frac"], "err", d["check"])
P
