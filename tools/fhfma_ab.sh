#!/bin/bash
# A/B of the opt-in FHFMA build.  Build both variants HERE first (no GPU needed):
#     python -m meshanything_b200.build ; MA_B200_FHFMA=1 python -m meshanything_b200.build
# then on the GPU box:  gpurun --timeout 900 -- 'bash tools/fhfma_ab.sh'
# 1. the bit-exact suites (oracle parity of every canonical kernel) with the FHFMA library: they prove or refute that
#    FHFMA == convert + FFMA on this hardware;  2. bench.py and the batched per-kernel timings for both builds.
set -u
mkdir -p gpurun_out
echo "== FHFMA build: bit-exact suites"
MA_B200_FHFMA=1 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -5
for v in 0 1; do
  echo "== bench, MA_B200_FHFMA=$v"
  MA_B200_FHFMA=$v timeout 300 python bench.py --steps 3 --warmup 3 2> gpurun_out/bench_fhfma$v.err | tail -1 > gpurun_out/bench_fhfma$v.json
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_fhfma$v.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], "short-ctx us", d["roofline"]["short_context"]["us_per_step"], d["check"])
PY
  MA_B200_FHFMA=$v timeout 120 python tools/bench_batched.py --skip-attention 2>/dev/null | tail -1 > gpurun_out/batched_fhfma$v.json
  python - <<PY
import json
d = json.load(open("gpurun_out/batched_fhfma$v.json"))
print(" ".join("%s:%.1fus" % (r["name"], r["canon"]["us"]) for r in d["linear"]))
PY
done
