#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== config 5 (batch 32, F=1600, sampling)"; timeout 1800 python bench.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline --lean > gpurun_out/bench_r02_cfg5.json 2> gpurun_out/bench_r02_cfg5.err; tail -2 gpurun_out/bench_r02_cfg5.err; python - <<P
import json
d=json.load(open("gpurun_out/bench_r02_cfg5.json"))
print("cfg5 value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["us_per_step_avg"])
P
echo "== profiles"; timeout 1500 bash tools/profile_round.sh r02 2>&1 | tail -8
