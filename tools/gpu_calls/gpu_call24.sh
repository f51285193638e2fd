#!/bin/bash
# round 2, GPU call 24: ncu of conv1_1 (CUDA-core first layer) and of the digit-split kernels
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_first_kernel|act_digits_kernel|pool_digits_kernel" -c 3 -o gpurun_out/r2_conv_first_full python tools/one_pair.py 700 1 > gpurun_out/c24.log 2>&1; echo "ncu rc=$?"
