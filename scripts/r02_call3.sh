#!/bin/bash
# Round-2 GPU call 3: tf32 mode (selftests, parity, bench) + re-run of the three tests fixed after call 2 +
# ncu --set full source-level captures of the epilogue-bound / load-bound kernels.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tc.py -m gpu -q -rA -s -p no:cacheprovider -k "tf32" --timeout=600 > gpurun_out/r02_c_pytest_tf32_tc.log 2>&1
echo "tc tf32 exit $?"; grep -E "passed|failed" gpurun_out/r02_c_pytest_tf32_tc.log | tail -2
grep -E "^FAILED" gpurun_out/r02_c_pytest_tf32_tc.log | head -30
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_sampler.py tests/test_gpu_configs.py -m gpu -q -rA -s -p no:cacheprovider \
   -k "tf32 or captured_window or batch8 or sampling_window or resync or dropout_statistics" --timeout=900 > gpurun_out/r02_c_pytest_misc.log 2>&1
echo "misc exit $?"; grep -E "passed|failed" gpurun_out/r02_c_pytest_misc.log | tail -2
grep -E "^FAILED|tf32|ACE " gpurun_out/r02_c_pytest_misc.log | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?"; tail -3 gpurun_out/smoke.log
for prec in tf32 bf16; do
  timeout 600 python bench.py --steps 20 --warmup 3 --precision $prec --no-cpu-baseline > gpurun_out/r02_c_bench_$prec.json 2> gpurun_out/r02_c_bench_$prec.err
  echo "bench $prec exit $?"
  python - $prec <<'PY'
import json, sys
try:
    r = json.loads(open(f"gpurun_out/r02_c_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(r["value"], 1), "ms/step", round(r["ms_per_step"], 3), "e2e", round(r["e2e"]["value"], 1), r["clocks"])
    print({k: v for k, v in r["roofline"]["per_kernel_ms"].items() if v > 0.3})
except Exception as exc:
    print("bench parse failed", exc); print(open(f"gpurun_out/r02_c_bench_{sys.argv[1]}.err").read()[-1500:])
PY
done
CASES="fc1 6 8 256 512 64800 1 3
dft 0 8 256 180 360 181 0
idft_epi7 4 8 256 180 360 181 7" bash scripts/ncu_selftest.sh
