#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== drop-in e2e"; timeout 900 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -s -k "forward_drop_in" 2>&1 | grep -E "end to end|passed|failed|Error|assert" | head
echo "== decoder tc"; timeout 900 python -m pytest tests/test_gpu_decoder.py -x -q -m gpu -s -k "tensor_core or sampling" 2>&1 | tail -3
echo "== default bench (config 2, with extras)"; timeout 1500 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -2 gpurun_out/bench_r02.err; python - <<P
import json
d=json.load(open("gpurun_out/bench_r02.json"))
r=d["roofline"]
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", r["frac"], "us/step", r["us_per_step_avg"], "short", r["short_context"]["us_per_step"], r["short_context"]["frac"], d["check"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"])
for k,v in (d.get("extra") or {}).items():
    print(k, v if "error" in k else [(c["context"], round(c["ms_per_step"],3), round(c["tokens_per_s"]), round(c["frac"],3)) for c in v["contexts"]], v.get("tokens_per_s_over_contexts") if isinstance(v, dict) else "")
P
echo "== config 3 (batch 64, F=800, sampling)"; timeout 1500 python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r02_cfg3.json 2> gpurun_out/bench_r02_cfg3.err; tail -2 gpurun_out/bench_r02_cfg3.err; python - <<P
import json
d=json.load(open("gpurun_out/bench_r02_cfg3.json"))
print("cfg3 value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "stage", d["config"]["stage_ms"], "frac", d["roofline"]["frac"], d["roofline"]["us_per_step_avg"], d["roofline"]["short_context"])
P
