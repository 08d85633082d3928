#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== ops"; timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention_decode_stream or weight_streaming" -x 2>&1 | tail -3
echo "== decoder"; timeout 1200 python -m pytest tests/test_gpu_decoder.py -q -m gpu -x 2>&1 | tail -3
for B in 64 8 32; do
echo "== batched attention $B"; timeout 900 python tools/bench_batched.py --batch $B --skip-linear 2>> gpurun_out/batched_r02b.err | python -c "
import json,sys; d=json.load(sys.stdin)
for a in d['attention']: print(a)
"
done
echo "== config 3 full"; timeout 1500 python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --lean > gpurun_out/bench_r02b_cfg3.json 2> gpurun_out/bench_r02b_cfg3.err; tail -2 gpurun_out/bench_r02b_cfg3.err
python - <<P
import json
d=json.load(open("gpurun_out/bench_r02b_cfg3.json"))
print("cfg3 value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["us_per_step_avg"], d["gpu_launches"])
P
