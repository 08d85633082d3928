#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for g in 0 200 400 700 1000; do
echo "== poll gap $g"; MA_B200_MEGA_POLL_GAP=$g timeout 300 python bench.py --faces 100 --steps 4 --warmup 2 --no-extra --no-cpu-baseline --lean 2> gpurun_out/gap_$g.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'us/step', round(d['roofline']['us_per_step_avg'],2), 'frac', round(d['roofline']['frac'],4), d['check'])
"
done
echo "== goldens with gap 400"; MA_B200_MEGA_POLL_GAP=400 timeout 900 python -m pytest tests/test_gpu_decoder.py -q -m gpu -x -k "golden or persistent or mega or F1600 or oracle" 2>&1 | tail -3
