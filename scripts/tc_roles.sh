#!/bin/bash
# role-wait profile of the tensor-core kernels (tc_debug bit 7): which role of the pipeline is blocked on which?
mkdir -p gpurun_out
: > gpurun_out/tc_roles.txt
while read -r name args; do
  [ -z "$name" ] && continue
  SFNO_TC_DEBUG=${DBG:-128} timeout 120 python tests/tc_selftest_cli.py $args | python -c "
import sys, json
r = json.loads(sys.stdin.readline()); c = r['counters']
print('$name', 'ms', round(r['ms'], 4), ' '.join(f'{k}={v:.3f}' if k not in ('cta_cycles','ctas') else f'{k}={v:.0f}' for k, v in c.items()))
" | tee -a gpurun_out/tc_roles.txt
done <<CASES
${CASES:-dft 0 8 256 180 360 181 0
leg_tri 1 8 256 180 180 181 1
dhconv_tri 2 8 256 180 181 1 0
ileg_tri 3 8 256 180 180 181 2
idft_epi7 4 8 256 180 360 181 7
fc1 6 8 256 512 64800 1 3
fc2 6 8 512 256 64800 0 5
skip 6 8 256 256 64800 1 1}
CASES
