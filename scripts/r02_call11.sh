#!/bin/bash
# A/B of engine switches through bench.py per-kernel times: 0 = shipped, 4096 = no stationary A, 8192 = L2 prefetch of the next
# tile's B operand, 12288 = both
set -u
mkdir -p gpurun_out
for rnd in 1 2; do
for dbg in 0 4096 8192 12288; do
  SFNO_TC_DEBUG=$dbg timeout 900 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-rollout > gpurun_out/ab_$dbg.json 2> gpurun_out/ab_$dbg.err
  python - $dbg $rnd <<'PY'
import json, sys
try:
    r = json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    k = r["roofline"]["per_kernel_ms"]
    print("dbg", sys.argv[1], "ms/step", round(r["ms_per_step"], 3), r["clocks"]["sm_mhz"], {n: round(k[n], 3) for n in ("dft_fwd", "legendre_fwd", "dhconv", "legendre_inv", "dft_inv", "inner_skip", "mlp_fc1", "mlp_fc2", "decoder0") if n in k})
except Exception as exc:
    print("bench parse failed", exc); print(open(f"gpurun_out/ab_{sys.argv[1]}.err").read()[-800:])
PY
done
done
