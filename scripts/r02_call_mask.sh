#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/mask_overlap_ab.py > gpurun_out/r02_q_mask_overlap_ab.json 2> gpurun_out/r02_q_mask_overlap_ab.err
echo "ab exit $?"; cat gpurun_out/r02_q_mask_overlap_ab.json; tail -5 gpurun_out/r02_q_mask_overlap_ab.err
timeout 900 python -m pytest tests/test_sampler.py tests/test_gpu_parity.py -q -m gpu -k "dropout or mask or captured or window or rollout" > gpurun_out/r02_q_pytest_dropout.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/r02_q_pytest_dropout.log
