#!/bin/bash
# One GPU session: per-family test processes (a sticky CUDA error only poisons its own process), smoke, bench.
# Usage (under gpurun): bash tools/gpu_round.sh [quick]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, args...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to python -m pytest -q -m gpu -p no:cacheprovider --timeout 240 "$@" > gpurun_out/$name.log 2>&1
  echo "exit $?" >> gpurun_out/$name.log
  tail -n 3 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
python __graft_entry__.py build > gpurun_out/build.log 2>&1; tail -n 2 gpurun_out/build.log | tee -a gpurun_out/summary.txt
run quant 600 tests/test_quantize_gpu.py
run woq_simt 600 tests/test_woq_matmul_gpu.py -k "simt or errors"
run woq_tc 900 tests/test_woq_matmul_gpu.py -k "tc or auto or linearity"
run attention 600 tests/test_attention_gpu.py
run glue 600 tests/test_glue_conv_gpu.py
run decoder 900 tests/test_decoder_gpu.py
echo "=== smoke" | tee -a gpurun_out/summary.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
tail -n 4 gpurun_out/smoke.log | tee -a gpurun_out/summary.txt
echo "=== bench" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --steps 64 --warmup 4 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
tail -n 5 gpurun_out/bench.log | cut -c1-3000 | tee -a gpurun_out/summary.txt
grep -h -E "^(FAILED|ERROR)" gpurun_out/*.log | head -60 | tee -a gpurun_out/summary.txt
