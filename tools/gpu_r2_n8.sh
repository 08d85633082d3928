#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --config 4 --gpus 8 --steps 1 --warmup 1 --lean --no-cpu-baseline > gpurun_out/bench_r02_cfg4_n8.json 2> gpurun_out/bench_r02_cfg4_n8.err
echo rc=$?; tail -3 gpurun_out/bench_r02_cfg4_n8.err; cat gpurun_out/bench_r02_cfg4_n8.json | head -c 1500
