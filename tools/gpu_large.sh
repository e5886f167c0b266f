#!/bin/bash
# persistent large-M GEMM: parity, sweep at M = 24000 (multicast clusters / plain persistent / per-tile kernel), encoder timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_woq_large_gpu.py -q 2>&1 | tail -5
B200_LARGE_MC=0 timeout 600 python -m pytest tests/test_woq_large_gpu.py -q 2>&1 | tail -3
for v in ${VARIANTS:-"B200_LARGE_MC=1" "B200_LARGE_MC=0"}; do
  echo "=== $v"
  env $v SWEEP_SHAPES=1280x1280,1280x3840,1280x5120,5120x1280 SWEEP_M=24000 timeout 300 python tools/gemm_sweep.py 2>&1 | tail -5
  env $v timeout 300 python tools/encoder_bench.py 2>&1 | tail -14
done
