#!/bin/bash
# End-of-round verification in ONE gpurun call: GPU tests, smoke, bench (bf16 + fp32), sampling window, interpolator
# profile, ncu launch list.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
NCU=${NCU:-1} PRECS="bf16 fp32" STEPS=${STEPS:-20} bash scripts/gpu_check.sh > gpurun_out/final_check.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu.log | tail -4
tail -2 gpurun_out/smoke.log
timeout 600 python bench_extra.py window > gpurun_out/extra_window.json 2>/dev/null
timeout 600 python scripts/profile_interp.py > gpurun_out/interp.json 2>/dev/null
python - <<'PY'
import json
for p in ("bf16", "fp32"):
    try:
        r = json.load(open(f"gpurun_out/bench_{p}.json"))
        print(p, round(r["value"], 1), "samples/s", round(r["ms_per_step"], 3), "ms  e2e", round(r["e2e"]["value"], 1), r["clocks"],
              r["roofline"]["kernel"], round(r["roofline"]["frac"], 3))
    except Exception as exc:
        print(p, "bench failed:", exc)
try:
    w = json.loads(open("gpurun_out/extra_window.json").read().strip().splitlines()[-1])
    print("window ms", round(w["ms_per_window"], 1))
    for line in open("gpurun_out/interp.json"):
        r = json.loads(line)
        print(r["model"], r["ms_total"], dict(list(r["per_kernel_ms"].items())[:4]))
except Exception as exc:
    print("extras failed:", exc)
PY
