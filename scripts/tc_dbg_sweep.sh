#!/bin/bash
# timing-only experiment: which resource bounds each tensor-core kernel?
# SFNO_TC_DEBUG bits: 1 skip A loads, 2 skip B loads, 4 skip global stores, 8 skip epilogue math, 16 skip MMAs,
# 32 skip TMEM loads (results are WRONG by construction).
mkdir -p gpurun_out
: > gpurun_out/dbg_sweep.txt
while read -r name args; do
  [ -z "$name" ] && continue
  for dbg in ${DBGS:-0 3 4 8 16 32 19 23 31 63}; do
    ms=$(SFNO_TC_DEBUG=$dbg timeout 120 python tests/tc_selftest_cli.py $args | python -c "import sys,json; print(round(json.loads(sys.stdin.readline())['ms'],4))")
    echo "$name dbg=$dbg ms=$ms" | tee -a gpurun_out/dbg_sweep.txt
  done
done <<CASES
${CASES:-fc1 6 8 256 512 64800 1 3
idft_epi7 4 8 256 180 360 181 7
dft 0 8 256 180 360 181 0}
CASES
