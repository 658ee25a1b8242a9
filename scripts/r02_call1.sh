#!/bin/bash
# Round-2 GPU call 1: (a) compute-sanitizer memcheck + racecheck over small tensor-core selftest cases,
# (b) A/B of the paired-TMEM-load drain loop (build/ldtm_pair) against the shipped library.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.csv 2>&1
SAN=/usr/local/cuda/bin/compute-sanitizer
: > gpurun_out/r02_sanitizer.txt
while read -r name args; do
  [ -z "$name" ] && continue
  for tool in memcheck racecheck; do
    echo "== $tool $name ($args)" >> gpurun_out/r02_sanitizer.txt
    timeout 600 $SAN --tool $tool --print-limit 5 python tests/tc_selftest_cli.py $args >> gpurun_out/r02_sanitizer.txt 2>&1
    echo "exit $?" >> gpurun_out/r02_sanitizer.txt
  done
done <<CASES
dft 0 1 8 36 72 37 0
leg_tri 1 1 8 36 36 37 1
dhconv_tri 2 2 16 36 37 1 0
ileg_tri 3 1 8 36 36 37 3
idft_epi7 4 1 8 36 72 37 7
conv_res 6 1 64 64 2592 1 7
conv_drop 6 1 64 128 2592 0 19
CASES
grep -E "^==|ERROR SUMMARY|exit|RACECHECK SUMMARY" gpurun_out/r02_sanitizer.txt | tail -60
bash scripts/ab_variant.sh ldtm_pair
