#!/bin/bash
# bench-only A/B of several library variants against the shipped library, interleaved: bash scripts/ab_multi.sh v1 v2 ...
set -u
mkdir -p gpurun_out
for round in 1 2; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-rollout > gpurun_out/ab_base_$round.json 2>/dev/null
  for NAME in "$@"; do
    SFNO_B200_LIB=build/$NAME/pkg/libsfno_b200.so timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-rollout > gpurun_out/ab_${NAME}_$round.json 2>/dev/null
  done
done
python - "$@" <<'PY'
import json, sys
for tag in ["base"] + sys.argv[1:]:
    for rnd in (1, 2):
        try:
            r = json.loads(open(f"gpurun_out/ab_{tag}_{rnd}.json").read().strip().splitlines()[-1])
            k = r["roofline"]["per_kernel_ms"]
            print(tag, rnd, round(r["ms_per_step"], 3), "ms", r["clocks"]["sm_mhz"], "MHz",
                  {n: k[n] for n in ("dft_inv", "mlp_fc1", "mlp_fc2", "inner_skip", "dft_fwd", "legendre_fwd", "legendre_inv", "dhconv") if n in k})
        except Exception as exc:
            print(tag, rnd, "failed:", exc)
PY
