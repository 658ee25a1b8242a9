#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k ragged > gpurun_out/r02_x_pytest_ragged.log 2>&1
echo "exit $?"; grep -E "ragged|passed|failed|Error" gpurun_out/r02_x_pytest_ragged.log | tail -24
