#!/bin/bash
# quick GPU iteration: tensor-core selftests (+ optional short bench)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -rA -p no:cacheprovider --timeout=600 > gpurun_out/pytest_tc.log 2>&1
echo "exit $?" >> gpurun_out/pytest_tc.log
grep -E "^\{" gpurun_out/pytest_tc.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['op'], r['dims'], 'err', r['max_err'], 'ref', round(r['max_ref'],3), 'tc', r['tc_used'], 'ms', round(r['ms'],4))
"
grep -E "FAILED|passed|failed|exit" gpurun_out/pytest_tc.log | tail -12
if [ "${BENCH:-1}" = "1" ]; then
  timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
  python -c "
import json
r=json.load(open('gpurun_out/bench_bf16.json'))
print('value', r['value'], 'ms/step', r['ms_per_step'], 'e2e', r['e2e']['value'])
for k,v in r['roofline']['per_kernel_ms'].items(): print(' ', k, v)
"
fi
