#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== decoder"; timeout 1200 python -m pytest tests/test_gpu_decoder.py -q -m gpu -x 2>&1 | tail -3
echo "== pipeline"; timeout 1200 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -x -k "drop_in or continuous" 2>&1 | tail -3
for v in 0 1; do
echo "== config 3, F=100, NO_PDL=$v"; MA_B200_NO_PDL=$v timeout 600 python bench.py --config 3 --faces 100 --steps 3 --warmup 2 --no-cpu-baseline --lean 2> gpurun_out/pdl_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], 'us/step', d['roofline']['us_per_step_avg'], 'frac', d['roofline']['frac'])
"
done
echo "== profiles"; timeout 900 bash tools/profile_batched.sh r02 2>&1 | tail -4
