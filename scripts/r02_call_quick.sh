#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
SFNO_NVTX=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-rollout 2>/dev/null | python -c "
import json,sys; r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nvtx on:', round(r['value'],1), round(r['ms_per_step'],3))"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_sampler.py -q -m gpu -k "golden_fp32 or capturable or captured or window" 2>&1 | tail -2
