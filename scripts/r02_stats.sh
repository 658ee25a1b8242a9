#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/stats_step_bench.py > gpurun_out/r02_o_stats_step_${N}gpu.json 2> gpurun_out/r02_o_stats_step_${N}gpu.err
echo "exit $?"; cat gpurun_out/r02_o_stats_step_${N}gpu.json; tail -5 gpurun_out/r02_o_stats_step_${N}gpu.err
