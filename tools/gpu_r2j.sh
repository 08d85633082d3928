#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== ops"; timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention_decode_stream or weight_streaming" -x 2>&1 | tail -5
echo "== decoder"; timeout 1200 python -m pytest tests/test_gpu_decoder.py -q -m gpu -x 2>&1 | tail -4
echo "== batched kernels 64"; timeout 900 python tools/bench_batched.py --batch 64 > gpurun_out/batched_r02b_64.json 2> gpurun_out/batched_r02b.err || tail -5 gpurun_out/batched_r02b.err
python - <<P
import json; d=json.load(open('gpurun_out/batched_r02b_64.json'))
for a in d['attention']: print(a)
for a in d['linear']: print(a['name'], {k:(v.get('us') if isinstance(v,dict) else v) for k,v in a.items() if k not in ('name','N','K')})
P
echo "== batched kernels 8"; timeout 900 python tools/bench_batched.py --batch 8 > gpurun_out/batched_r02b_8.json 2>> gpurun_out/batched_r02b.err
python - <<P
import json; d=json.load(open('gpurun_out/batched_r02b_8.json'))
for a in d['attention']: print(a)
for a in d['linear']: print(a['name'], {k:(v.get('us') if isinstance(v,dict) else v) for k,v in a.items() if k not in ('name','N','K')})
P
