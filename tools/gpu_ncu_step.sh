#!/bin/bash
# ncu --set full capture (with source-level stall samples) of ONE launch of the persistent step kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python __graft_entry__.py build > gpurun_out/build.log 2>&1; tail -n 1 gpurun_out/build.log
B200_STEP_KERNEL=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_step -c 1 -o gpurun_out/prof_step_r2 -f python bench.py --profile > gpurun_out/ncu_step_r2.log 2>&1; echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_step_r2.log
