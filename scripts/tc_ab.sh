#!/bin/bash
# A/B timing of tensor-core engine variants through tc_debug bits (64: LDS/STG epilogue, 256: single-M tiles)
mkdir -p gpurun_out
: > gpurun_out/tc_ab.txt
while read -r name args; do
  [ -z "$name" ] && continue
  for dbg in ${DBGS:-0 256}; do
    SFNO_TC_DEBUG=$dbg timeout 120 python tests/tc_selftest_cli.py $args | python -c "
import sys, json
r = json.loads(sys.stdin.readline())
print('$name', 'dbg=$dbg', 'ms', round(r['ms'], 4), 'err', r['max_err'], 'status', r['status'], r['error'][:80])
" | tee -a gpurun_out/tc_ab.txt
  done
done <<CASES
${CASES:-dft 0 8 256 180 360 181 0
leg_tri 1 8 256 180 180 181 1
leg_full 1 8 256 180 180 181 0
dhconv_tri 2 8 256 180 181 1 0
ileg_tri 3 8 256 180 180 181 2
ileg_x 3 2 64 180 180 181 3}
CASES
