#!/bin/bash
# compute-sanitizer over the tensor-core kernels added for the backward pass (weight gradients, adjoint contraction)
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02_w_sanitizer_backward.txt
: > $OUT
K='(fused_spectral_conv_gradients and 16-32 and dhconv-64-64 and bf16) or (fused_spectral_conv_gradients and 24-48 and dhconv-64-64 and tf32) or (conv1x1_ex_gradients and 2-64-96-30-64 and (bf16 or tf32))'
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool :: pytest tests/test_gpu_backward.py -k \"$K\"" >> $OUT
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_backward.py -q -m gpu -p no:cacheprovider -k "$K" > gpurun_out/san_$tool.log 2>&1
  echo "exit $?" >> $OUT
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/san_$tool.log | tail -6 >> $OUT
done
cat $OUT
