#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/build.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:cross_attention_kernel -c 2 \
   -f -o gpurun_out/prof_xattn python bench.py --profile > gpurun_out/ncu_xattn.log 2>&1; echo "exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:mmha_generation_kernel -c 2 \
   -f -o gpurun_out/prof_mmha python bench.py --profile > gpurun_out/ncu_mmha.log 2>&1; echo "exit $?"
