#!/bin/bash
# GPU session: waveform -> tokens pipeline tests (+ log-mel tests again).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 240 -x tests/test_pipeline_gpu.py tests/test_log_mel_gpu.py > gpurun_out/pipeline.log 2>&1
echo "exit $?" >> gpurun_out/pipeline.log; tail -n 40 gpurun_out/pipeline.log
