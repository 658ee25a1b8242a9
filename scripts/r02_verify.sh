#!/bin/bash
# last verification of the tree: full -m gpu suite + smoke + a short bench (no rollout / cpu legs)
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=1500 > gpurun_out/r02_zzz_pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r02_zzz_pytest_gpu.log
tail -4 gpurun_out/r02_zzz_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-rollout > gpurun_out/r02_zzz_bench_short.json 2> gpurun_out/r02_zzz_bench_short.err; echo "bench exit $?"
python -c "
import json; r=json.loads(open('gpurun_out/r02_zzz_bench_short.json').read().strip().splitlines()[-1]); print(round(r['value'],1), round(r['ms_per_step'],3), round(r['e2e']['value'],1), r['clocks'])"
