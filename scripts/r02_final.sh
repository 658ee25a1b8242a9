#!/bin/bash
# final check of the round on one GPU: full -m gpu suite, smoke, default bench (what the driver runs), reference arm
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rA -s -p no:cacheprovider --timeout=1500 --durations=12 > gpurun_out/r02_r_pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r02_r_pytest_gpu.log
grep -E "passed|failed|error" gpurun_out/r02_r_pytest_gpu.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/r02_r_pytest_gpu.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?"; tail -3 gpurun_out/smoke.log
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_r_bench_1gpu.json 2> gpurun_out/r02_r_bench_1gpu.err
echo "bench exit $?"
python - <<'PY'
import json
try:
    r = json.loads(open("gpurun_out/r02_r_bench_1gpu.json").read().strip().splitlines()[-1])
    print("value", round(r["value"], 1), "ms/step", round(r["ms_per_step"], 3), "e2e", round(r["e2e"]["value"], 1), r["clocks"], "launches", r["gpu_launches"])
    ro = r["roofline"]; print(ro["kernel"], round(ro["frac"], 3), ro["unit"], round(ro["achieved"], 1), "traffic", ro.get("traffic"))
    print({k: (v["frac_hbm"], v.get("frac_hbm_executed")) for k, v in ro["per_kernel_roofline"].items()})
    print(ro["tensor_pipe"])
    print("cpu", r.get("cpu_baseline")); print("rollout", {k: v for k, v in r.get("ensemble_rollout", {}).items() if k != "workload"})
except Exception as exc:
    print("bench parse failed", exc); print(open("gpurun_out/r02_r_bench_1gpu.err").read()[-1500:])
PY
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02_r_bench_reference.json 2>&1; tail -c 600 gpurun_out/r02_r_bench_reference.json

# one simulated year (1460 six-hour steps = 244 windows) of the 25-member ensemble on this GPU, statistics every step
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --rollout-windows 244 > gpurun_out/r02_r_bench_1gpu_1year.json 2> gpurun_out/r02_r_bench_1gpu_1year.err
echo "1-year exit $?"
python - <<'PY'
import json
try:
    r = json.loads(open("gpurun_out/r02_r_bench_1gpu_1year.json").read().strip().splitlines()[-1])
    print("1-year rollout", {k: v for k, v in r["ensemble_rollout"].items() if k != "workload"})
except Exception as exc:
    print("parse failed", exc); print(open("gpurun_out/r02_r_bench_1gpu_1year.err").read()[-1500:])
PY
