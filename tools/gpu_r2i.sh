#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== stream attention op"; timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention_decode_stream or attention_bit_exact" -x 2>&1 | tail -5
echo "== decoder"; timeout 1200 python -m pytest tests/test_gpu_decoder.py -q -m gpu -x 2>&1 | tail -6
echo "== batched kernels"; timeout 600 python tools/bench_batched.py --batch 64 --skip-linear > gpurun_out/batched_attn_r02b.json 2> gpurun_out/batched_attn_r02b.err || tail -5 gpurun_out/batched_attn_r02b.err
python -c "
import json; d=json.load(open('gpurun_out/batched_attn_r02b.json'))
for a in d['attention']: print(a)
"
echo "== batch 8"; timeout 600 python tools/bench_batched.py --batch 8 --skip-linear 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
for a in d['attention']: print(a)
"
