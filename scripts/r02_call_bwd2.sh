#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_parity.py -q -m gpu -k "backward or gradients or instance_norm or tf32 or training" > gpurun_out/r02_z_pytest.log 2>&1
echo "exit $?"; tail -2 gpurun_out/r02_z_pytest.log
timeout 600 python scripts/train_step_timing.py > gpurun_out/r02_z_train_step.log 2>&1
echo "train-step exit $?"; cut -c1-420 gpurun_out/r02_z_train_step.log | tail -4
