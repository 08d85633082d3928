#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== decoder"; timeout 900 python -m pytest tests/test_gpu_decoder.py -q -m gpu 2>&1 | tail -3
echo "== default bench"; timeout 900 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -2 gpurun_out/bench_final2.err
python - <<P
import json
d=json.loads(open("gpurun_out/bench_final2.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "us/step", d["roofline"]["us_per_step_avg"], "cpu", d.get("cpu_baseline",{}).get("value"), d["clocks"])
P
echo "== launch list config 3"; timeout 600 bash tools/profile_cfg3.sh r02 2>&1 | tail -2
