#!/bin/bash
# timeline + ncu of the q-fused cross-attention inside the decoder step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LAYERS=6 B200_FUSE_XQ=1 timeout 300 python tools/step_timeline.py > gpurun_out/timeline_xq1.txt 2>&1
LAYERS=6 B200_FUSE_XQ=0 timeout 300 python tools/step_timeline.py > gpurun_out/timeline_xq0.txt 2>&1
B200_FUSE_XQ=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:cross_attention_qproj -s 3 -c 1 \
   -f -o gpurun_out/prof_xq python bench.py --profile > gpurun_out/ncu_xq.log 2>&1; echo "ncu exit $?"
for v in "B200_FUSE_XQ=1" "B200_FUSE_XQ=0"; do echo "[$v] $(env $v timeout 240 python bench.py --steps 64 --warmup 4 --no-cpu-baseline 2>&1 | grep -o "\"ms_per_step\": [0-9.]*" | head -1)"; done
