#!/bin/bash
# GPU session for the log-Mel front end: parity tests, conv tests (256-step tile), timing.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 240 tests/test_log_mel_gpu.py > gpurun_out/logmel.log 2>&1
echo "exit $?" >> gpurun_out/logmel.log; tail -n 25 gpurun_out/logmel.log
timeout 300 python tools/log_mel_bench.py > gpurun_out/logmel_bench.txt 2>&1; cat gpurun_out/logmel_bench.txt
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 240 tests/test_glue_conv_gpu.py -k conv1d 2>&1 | tail -n 3
