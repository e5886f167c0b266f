#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider --timeout 240 -x tests/test_glue_conv_gpu.py tests/test_encoder_gpu.py tests/test_pipeline_gpu.py tests/test_woq_matmul_gpu.py tests/test_decoder_gpu.py 2>&1 | tail -n 12
