#!/bin/bash
# Round-2 GPU call 2: full -m gpu suite (new: config tests, graph-captured window, ensemble moments, fingerprints)
set -u
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt; free -g | head -2 >> gpurun_out/nproc.txt
timeout 2400 python -m pytest tests -m gpu -q -rA -s -p no:cacheprovider --timeout=1500 --durations=15 > gpurun_out/r02_b_pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r02_b_pytest_gpu.log
grep -E "passed|failed|error" gpurun_out/r02_b_pytest_gpu.log | tail -5
grep -E "^FAILED|^ERROR" gpurun_out/r02_b_pytest_gpu.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?"; tail -2 gpurun_out/smoke.log
