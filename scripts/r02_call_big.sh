#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu -s -k "2_31" > gpurun_out/r02_y_pytest_big.log 2>&1
echo "exit $?"; grep -E "720x1440|passed|failed|Error|error" gpurun_out/r02_y_pytest_big.log | tail -12; nvidia-smi --query-gpu=memory.used --format=csv | tail -1
