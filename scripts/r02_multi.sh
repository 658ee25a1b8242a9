#!/bin/bash
# multi-GPU bench (forward throughput + sharded ensemble rollout with the NCCL statistics gather); N from $1
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r02_r_bench_${N}gpu.json 2> gpurun_out/r02_r_bench_${N}gpu.err
echo "bench N=$N exit $?"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    r = json.loads(open(f"gpurun_out/r02_r_bench_{n}gpu.json").read().strip().splitlines()[-1])
    print("N", n, "value", round(r["value"], 1), "ms/step", round(r["ms_per_step"], 3), "e2e", round(r["e2e"]["value"], 1))
    print("rollout", json.dumps(r.get("ensemble_rollout")))
except Exception as exc:
    print("parse failed", exc); print(open(f"gpurun_out/r02_r_bench_{n}gpu.err").read()[-2500:])
PY

if [ "${YEAR:-0}" = "1" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --rollout-windows 244 > gpurun_out/r02_r_bench_${N}gpu_1year.json 2> gpurun_out/r02_r_bench_${N}gpu_1year.err
  echo "1-year N=$N exit $?"
  python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    r = json.loads(open(f"gpurun_out/r02_r_bench_{n}gpu_1year.json").read().strip().splitlines()[-1])
    print("1-year rollout", {k: v for k, v in r["ensemble_rollout"].items() if k != "workload"})
except Exception as exc:
    print("parse failed", exc)
PY
fi
