#!/bin/bash
# Round-2 GPU call 8: sub-module boundary tests, step glue, ensemble statistics, captured window; then full suite timing.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_sampler.py -m gpu -q -rA -s -p no:cacheprovider \
   -k "spectral_conv_module or conv1x1_ex or step_glue or ensemble or captured_window or cold_update or resync" --timeout=900 > gpurun_out/r02_f_pytest_boundary.log 2>&1
echo "boundary exit $?"; grep -E "passed|failed" gpurun_out/r02_f_pytest_boundary.log | tail -2
grep -E "^FAILED|SpectralConvS2\[|conv1x1_ex\[" gpurun_out/r02_f_pytest_boundary.log | head -60
grep -E "^E  " gpurun_out/r02_f_pytest_boundary.log | head -20
