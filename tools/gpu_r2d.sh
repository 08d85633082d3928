#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== ops ws"; timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "weight_streaming or tensor_core" 2>&1 | tail -12
echo "== decoder tc"; timeout 900 python -m pytest tests/test_gpu_decoder.py -x -q -m gpu -s -k "tensor_core or sampling or greedy_bit_exact or timeout or batch_invariance" 2>&1 | tail -12
echo "== batched kernels"; timeout 600 python tools/bench_batched.py --skip-attention > gpurun_out/batched_r02.json 2> gpurun_out/batched_r02.err; tail -2 gpurun_out/batched_r02.err; python - <<P
import json
d=json.load(open("gpurun_out/batched_r02.json"))
for r in d["linear"]:
    print(r["name"], {k:(v.get("us"), v.get("weight_GBps")) if isinstance(v,dict) else v for k,v in r.items() if k in ("canon","tcgen05","tcgen05_ws")})
P
echo "== trace short"; timeout 300 python tools/trace_mega.py 30 > gpurun_out/mega_trace_r02d.txt 2>&1; head -13 gpurun_out/mega_trace_r02d.txt
echo "== bench b64 sampling f100"; timeout 900 python bench.py --batch 64 --faces 100 --sampling --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_b64_r02.json 2> gpurun_out/bench_b64_r02.err; tail -2 gpurun_out/bench_b64_r02.err; python - <<P
import json
d=json.load(open("gpurun_out/bench_b64_r02.json"))
print("b64 f100 sampling: value", d["value"], "ms/step", d["ms_per_step"], "stage", d["config"]["stage_ms"], "roofline", d["roofline"]["frac"], d["roofline"]["us_per_step_avg"])
P
