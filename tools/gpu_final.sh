#!/bin/bash
# Round-end regression on one B200: build check + smoke, the whole GPU suite, the default bench line, a full config-5 run.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6
echo "== default bench"; timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
python - <<P
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "us/step", d["roofline"]["us_per_step_avg"], "cpu", d.get("cpu_baseline",{}).get("value"), d["clocks"])
for k,v in (d.get("extra") or {}).items():
    print(k, [(c["context"], round(c["ms_per_step"],3), round(c["tokens_per_s"]), round(c["frac"],3)) for c in v["contexts"]])
P
echo "== config 5"; timeout 900 python bench.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline --lean > gpurun_out/bench_cfg5_final.json 2> gpurun_out/bench_cfg5_final.err; tail -2 gpurun_out/bench_cfg5_final.err
python - <<P
import json
d=json.loads(open("gpurun_out/bench_cfg5_final.json").read().strip().splitlines()[-1])
print("cfg5 value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["us_per_step_avg"])
P
