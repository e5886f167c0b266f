#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/build.log 2>&1
for cfg in ${XA_CFGS:-C}; do
B200_XA_CFG=$cfg timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:cross_attention_kernel -c 1 \
   -f -o gpurun_out/prof_xattn_$cfg python bench.py --profile > gpurun_out/ncu_xattn.log 2>&1; echo "exit $?"
done
