#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== batched kernels"; timeout 600 python tools/bench_batched.py > gpurun_out/batched_r02.json 2> gpurun_out/batched_r02.err; tail -2 gpurun_out/batched_r02.err; python - <<P
import json
d=json.load(open("gpurun_out/batched_r02.json"))
for r in d["attention"]: print(r)
for r in d["linear"]:
    print(r["name"], {k:(v.get("us"), v.get("weight_GBps")) if isinstance(v,dict) else v for k,v in r.items() if k in ("canon","tcgen05","tcgen05_ws")})
P
echo "== trace"; timeout 300 python tools/trace_mega.py 30 > gpurun_out/mega_trace_r02e.txt 2>&1; head -13 gpurun_out/mega_trace_r02e.txt
