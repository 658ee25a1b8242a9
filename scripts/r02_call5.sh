#!/bin/bash
set -u
mkdir -p gpurun_out
for dbg in 0 2048; do
  echo "== dbg $dbg"
  SFNO_TC_DEBUG=$dbg timeout 120 python tests/tc_selftest_cli.py 4 2 16 180 360 181 0 2>&1 | tail -1 | cut -c1-300
  SFNO_TC_DEBUG=$dbg timeout 120 python tests/tc_selftest_cli.py 4 1 8 64 128 65 0 2>&1 | tail -1 | cut -c1-300
done
timeout 120 compute-sanitizer --tool memcheck --print-limit 3 python tests/tc_selftest_cli.py 4 1 8 64 128 65 0 2>&1 | grep -v "^=========     " | head -30
