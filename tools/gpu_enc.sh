#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider --timeout 240 -x tests/test_woq_matmul_gpu.py tests/test_encoder_gpu.py tests/test_decoder_gpu.py 2>&1 | tail -n 8
SWEEP_M=128,256,300,512,1500 timeout 300 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_midM.txt 2>&1; tail -n 20 gpurun_out/gemm_sweep_midM.txt
