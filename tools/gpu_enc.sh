#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider --timeout 240 -x tests/test_attention_gpu.py tests/test_encoder_gpu.py tests/test_pipeline_gpu.py 2>&1 | tail -n 8
timeout 300 python tools/encoder_bench.py > gpurun_out/encoder_bench.txt 2>&1; grep -v arn gpurun_out/encoder_bench.txt | tail -n 18
