#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"OpWgradTc" -c 1 -f -o gpurun_out/prof_conv_wgrad python scripts/bwd_kernels_once.py > gpurun_out/ncu_bwd1.log 2>&1; echo "exit $?"
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"OpDhconvWgrad" -c 1 -f -o gpurun_out/prof_dhconv_wgrad python scripts/bwd_kernels_once.py > gpurun_out/ncu_bwd2.log 2>&1; echo "exit $?"
python scripts/summarize_ncu.py gpurun_out/r02_zb_ncu_backward_kernels.md
rm -f gpurun_out/prof_*.ncu-rep
