#!/bin/bash
# dropout masks by a dedicated kernel: parity of the dropout paths + interpolator timing + window timing
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_sampler.py tests/test_gpu_tc.py -m gpu -q -p no:cacheprovider -k "dropout or captured or rollout or conv1x1_ex or 21 or 19" --timeout=600 2>&1 | tail -3
timeout 600 python scripts/profile_interp.py bf16 > gpurun_out/r02_j_interp.json 2>gpurun_out/r02_j_interp.err; cat gpurun_out/r02_j_interp.json; tail -3 gpurun_out/r02_j_interp.err
timeout 600 python bench_extra.py window --steps 5 > gpurun_out/r02_j_window.json 2>&1; tail -1 gpurun_out/r02_j_window.json | cut -c1-600
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -s -p no:cacheprovider -k "triangular or injected or identical" --timeout=600 2>&1 | grep -E "passed|failed|rel-L2|dropout masks|^E " | head
