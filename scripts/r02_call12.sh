#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -p no:cacheprovider -k "conv or dft" --timeout=600 2>&1 | tail -2
for rnd in 1 2; do
for dbg in 0 4096; do
  SFNO_TC_DEBUG=$dbg timeout 900 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-rollout > gpurun_out/ab_$dbg.json 2> gpurun_out/ab_$dbg.err
  python - $dbg $rnd <<'PY'
import json, sys
try:
    r = json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    k = r["roofline"]["per_kernel_ms"]
    print("dbg", sys.argv[1], "ms/step", round(r["ms_per_step"], 3), r["clocks"]["sm_mhz"], {n: round(k[n], 3) for n in ("dft_fwd", "dft_inv", "inner_skip", "mlp_fc1", "mlp_fc2", "encoder1", "decoder0", "decoder1", "encoder0") if n in k})
except Exception as exc:
    print("bench parse failed", exc); print(open(f"gpurun_out/ab_{sys.argv[1]}.err").read()[-800:])
PY
done
done
