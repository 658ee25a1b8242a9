#!/bin/bash
# On the GPU box (via gpurun): parity tests and the per-kernel bench profile with the shipped library and with a variant
# built by scripts/build_variant.sh.   usage: bash scripts/ab_variant.sh ldtm_pair
set -u
NAME=$1
LIBV=build/$NAME/pkg/libsfno_b200.so
mkdir -p gpurun_out
[ -f "$LIBV" ] || { echo "$LIBV missing: run scripts/build_variant.sh $NAME <flags> in the build container first"; exit 1; }
SFNO_B200_LIB=$LIBV timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/ab_${NAME}_pytest.log 2>&1
echo "variant pytest exit $?"; tail -3 gpurun_out/ab_${NAME}_pytest.log
for round in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_base_$round.json 2>/dev/null
  SFNO_B200_LIB=$LIBV timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_${NAME}_$round.json 2>/dev/null
done
python - "$NAME" <<'PY'
import json, sys
name = sys.argv[1]
for tag in ("base", name):
    for rnd in (1, 2):
        try:
            r = json.loads(open(f"gpurun_out/ab_{tag}_{rnd}.json").read().strip().splitlines()[-1])
            k = r["roofline"]["per_kernel_ms"]
            print(tag, rnd, round(r["ms_per_step"], 3), "ms", r["clocks"]["sm_mhz"], "MHz",
                  {n: k[n] for n in ("dft_inv", "dft_fwd", "legendre_inv", "mlp_fc1") if n in k})
        except Exception as exc:
            print(tag, rnd, "failed:", exc)
PY
