#!/bin/bash
# launch list of the batched decode step (config 3, short run): the kernels' shares of a step
R=${1:-r02}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_${R}_cfg3.csv \
    python bench.py --config 3 --faces 16 --steps 1 --warmup 1 --no-cpu-baseline --lean > gpurun_out/launches_${R}_cfg3.log 2>&1
wc -l gpurun_out/launches_${R}_cfg3.csv
