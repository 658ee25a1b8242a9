#!/bin/bash
# Runs on the GPU box (via gpurun): GPU parity tests, smoke, a short bench and an ncu launch list.
# Everything is written under gpurun_out/ so it comes back to the build container.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.csv 2>&1
nproc > gpurun_out/nproc.txt
STEPS=${STEPS:-5}
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
for prec in ${PRECS:-bf16 fp32}; do
  timeout 900 python bench.py --steps $STEPS --warmup 3 --precision $prec > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err
  echo "bench $prec exit: $?"; tail -c 1500 gpurun_out/bench_$prec.json
done
if [ "${NCU:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --precision ${NCU_PREC:-bf16} > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit: $?"
fi
