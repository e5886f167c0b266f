#!/bin/bash
# ncu --set full of the persistent large-M GEMM (M = 24000): plain 1280 -> 3840 and 1280 -> 5120 with bias + GELU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:woq_gemm_large_kernel -s 2 -c 1 -f \
   -o gpurun_out/prof_r2_gemm_large python tools/gemm_one.py 24000 1280 3840 > gpurun_out/ncu_r2_large.log 2>&1; echo "exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:woq_gemm_large_kernel -s 2 -c 1 -f \
   -o gpurun_out/prof_r2_gemm_large_gelu python tools/gemm_one.py 24000 1280 5120 gelu > gpurun_out/ncu_r2_large_gelu.log 2>&1; echo "exit $?"
