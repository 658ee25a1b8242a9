#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + full captures of the hot kernels (selftest shapes = ACE, batch 8)
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 420 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --precision bf16 > gpurun_out/ncu_bench.log 2>&1
echo "launch list exit $?"
CASES="fc1 6 8 256 512 64800 1 3
fc2 6 8 512 256 64800 0 5
skip 6 8 256 256 64800 1 1
idft 4 8 256 180 360 181 7
dft 0 8 256 180 360 181 0
leg 1 8 256 180 180 181 0
ileg 3 8 256 180 180 181 0
dhconv 2 8 256 180 181 0 0" bash scripts/ncu_selftest.sh
