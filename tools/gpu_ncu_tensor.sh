#!/bin/bash
# ncu --set full of the tensor-bound kernels: the M = 24000 weight-only GEMM and the conv2 implicit GEMM.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/build.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:woq_gemm_tc_kernel -s 2 -c 1 -f \
   -o gpurun_out/prof_gemm_m24000 python tools/gemm_one.py 24000 1280 3840 > gpurun_out/ncu_m24000.log 2>&1; echo "exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc_kernel -s 3 -c 1 -f \
   -o gpurun_out/prof_conv2 python tools/conv_bench.py > gpurun_out/ncu_conv2.log 2>&1; echo "exit $?"
