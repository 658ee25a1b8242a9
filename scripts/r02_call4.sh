#!/bin/bash
# Round-2 GPU call 4: 32-row TMA boxes in the epilogue + MN-major tf32 swizzle fix: all tensor-core selftests,
# tf32 parity, captured window, golden / ACE parity, bench bf16 (with the ensemble-rollout leg) and tf32.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tc.py -m gpu -q -rA -s -p no:cacheprovider --timeout=600 > gpurun_out/r02_d_pytest_tc.log 2>&1
echo "tc exit $?"; grep -E "passed|failed" gpurun_out/r02_d_pytest_tc.log | tail -2
grep -E "^FAILED" gpurun_out/r02_d_pytest_tc.log | head -30
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_sampler.py -m gpu -q -rA -s -p no:cacheprovider --timeout=900 > gpurun_out/r02_d_pytest_parity.log 2>&1
echo "parity exit $?"; grep -E "passed|failed" gpurun_out/r02_d_pytest_parity.log | tail -2
grep -E "^FAILED|tf32" gpurun_out/r02_d_pytest_parity.log | head -30
for prec in bf16 tf32; do
  extra=""; [ $prec = tf32 ] && extra="--no-rollout"
  timeout 900 python bench.py --steps 40 --warmup 3 --precision $prec --no-cpu-baseline $extra > gpurun_out/r02_d_bench_$prec.json 2> gpurun_out/r02_d_bench_$prec.err
  echo "bench $prec exit $?"
  python - $prec <<'PY'
import json, sys
try:
    r = json.loads(open(f"gpurun_out/r02_d_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(r["value"], 1), "ms/step", round(r["ms_per_step"], 3), "e2e", round(r["e2e"]["value"], 1), r["clocks"])
    print({k: v for k, v in r["roofline"]["per_kernel_ms"].items() if v > 0.3})
    print("rollout", r.get("ensemble_rollout"))
except Exception as exc:
    print("bench parse failed", exc); print(open(f"gpurun_out/r02_d_bench_{sys.argv[1]}.err").read()[-1500:])
PY
done
