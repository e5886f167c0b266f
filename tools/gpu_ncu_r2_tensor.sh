#!/bin/bash
# Round 2: fresh ncu --set full captures of the tensor-bound encoder kernels on the final tree: the M = 24000 weight-only
# GEMMs (plain and GELU) and the tcgen05 attention.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python __graft_entry__.py build > gpurun_out/build.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:woq_gemm_tc_kernel -s 2 -c 1 -f \
   -o gpurun_out/prof_r2_gemm_m24000 python tools/gemm_one.py 24000 1280 3840 > gpurun_out/ncu_r2_m24000.log 2>&1; echo "exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:woq_gemm_tc_kernel -s 2 -c 1 -f \
   -o gpurun_out/prof_r2_gemm_m24000_gelu python tools/gemm_one.py 24000 1280 5120 gelu > gpurun_out/ncu_r2_m24000_gelu.log 2>&1; echo "exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_bidir_tc -s 2 -c 1 -f \
   -o gpurun_out/prof_r2_encattn python tools/enc_attn_time.py > gpurun_out/ncu_r2_encattn.log 2>&1; echo "exit $?"
