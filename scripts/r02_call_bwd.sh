#!/bin/bash
# Round-2 GPU call: backward-pass gradient tests (row f4) + training-step timing at the ACE size.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_backward.py -q -m gpu -s > gpurun_out/r02_p_pytest_backward.log 2>&1
echo "backward exit $?"
tail -3 gpurun_out/r02_p_pytest_backward.log
grep -E "worst|fused spectral|conv1x1_ex grad|FAILED|Error|assert" gpurun_out/r02_p_pytest_backward.log | head -90
timeout 600 python scripts/train_step_timing.py > gpurun_out/r02_p_train_step.log 2>&1
echo "train-step exit $?"; tail -6 gpurun_out/r02_p_train_step.log
