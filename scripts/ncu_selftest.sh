#!/bin/bash
# ncu --set full on tensor-core selftest cases; reports come back under gpurun_out/
set -u
mkdir -p gpurun_out
i=0
while read -r name args; do
  [ -z "$name" ] && continue
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 2 -f -o gpurun_out/prof_$name \
     python tests/tc_selftest_cli.py $args > gpurun_out/ncu_$name.log 2>&1
  echo "$name exit $?"
done <<CASES
${CASES}
CASES
ls -la gpurun_out/*.ncu-rep
