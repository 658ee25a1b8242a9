#!/bin/bash
# bench-only A/B of a library variant (scripts/build_variant.sh) against the shipped library: bash scripts/ab_bench_only.sh <name>
set -u
NAME=$1
LIBV=build/$NAME/pkg/libsfno_b200.so
mkdir -p gpurun_out
for round in 1 2; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-rollout > gpurun_out/ab_base_$round.json 2>/dev/null
  SFNO_B200_LIB=$LIBV timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-rollout > gpurun_out/ab_${NAME}_$round.json 2>/dev/null
done
SFNO_B200_LIB=$LIBV timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ace_sized or golden" -x > gpurun_out/ab_${NAME}_pytest.log 2>&1; tail -2 gpurun_out/ab_${NAME}_pytest.log
python - "$NAME" <<'PY'
import json, sys
name = sys.argv[1]
for tag in ("base", name):
    for rnd in (1, 2):
        try:
            r = json.loads(open(f"gpurun_out/ab_{tag}_{rnd}.json").read().strip().splitlines()[-1])
            k = r["roofline"]["per_kernel_ms"]
            print(tag, rnd, round(r["ms_per_step"], 3), "ms", r["clocks"]["sm_mhz"], "MHz",
                  {n: k[n] for n in ("dft_inv", "mlp_fc1", "mlp_fc2", "inner_skip", "encoder1", "decoder0") if n in k})
        except Exception as exc:
            print(tag, rnd, "failed:", exc)
PY
