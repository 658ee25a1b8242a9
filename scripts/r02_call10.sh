#!/bin/bash
# Round-2 GPU call 10: stationary-A kernels (DFT, conv): selftests, parity, A/B bench (tc_debug 4096 = streaming kernels)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tc.py -m gpu -q -rA -s -p no:cacheprovider --timeout=600 > gpurun_out/r02_h_pytest_tc.log 2>&1
echo "tc exit $?"; grep -E "passed|failed" gpurun_out/r02_h_pytest_tc.log | tail -2
grep -E "^FAILED" gpurun_out/r02_h_pytest_tc.log | head -30
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_sampler.py -m gpu -q -rA -s -p no:cacheprovider --timeout=900 > gpurun_out/r02_h_pytest_parity.log 2>&1
echo "parity exit $?"; grep -E "passed|failed" gpurun_out/r02_h_pytest_parity.log | tail -2
grep -E "^FAILED" gpurun_out/r02_h_pytest_parity.log | head -30
for rnd in 1 2; do
for dbg in 0 4096; do
  SFNO_TC_DEBUG=$dbg timeout 900 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-rollout > gpurun_out/r02_h_bench_dbg${dbg}_$rnd.json 2> gpurun_out/r02_h_bench_dbg${dbg}_$rnd.err
  python - $dbg $rnd <<'PY'
import json, sys
try:
    r = json.loads(open(f"gpurun_out/r02_h_bench_dbg{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
    k = r["roofline"]["per_kernel_ms"]
    print("dbg", sys.argv[1], "value", round(r["value"], 1), "ms/step", round(r["ms_per_step"], 3), r["clocks"]["sm_mhz"], {n: k[n] for n in ("dft_fwd", "mlp_fc1", "inner_skip", "encoder1", "decoder0", "dft_inv", "mlp_fc2") if n in k})
except Exception as exc:
    print("bench parse failed", exc); print(open(f"gpurun_out/r02_h_bench_dbg{sys.argv[1]}_{sys.argv[2]}.err").read()[-1500:])
PY
done
done
