#!/bin/bash
# persistent large-M GEMM: parity, sweep at M = 24000 with / without it, encoder timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_woq_large_gpu.py -x -q 2>&1 | tail -15
for thr in 4096 0; do
  echo "=== B200_LARGE_M=$thr"
  B200_LARGE_M=$thr SWEEP_SHAPES=1280x1280,1280x3840,1280x5120,5120x1280 SWEEP_M=6000,24000 timeout 300 python tools/gemm_sweep.py 2>&1 | tail -9
  B200_LARGE_M=$thr timeout 300 python tools/encoder_bench.py 2>&1 | tail -12
done
