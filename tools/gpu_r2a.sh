#!/bin/bash
# Round-2 GPU session A: whole GPU suite, smoke, bench, sweep row of the logits shape, ncu of MMHA.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python __graft_entry__.py build > gpurun_out/build.log 2>&1; tail -n 2 gpurun_out/build.log
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/all_gpu.log 2>&1; echo "pytest exit $?"
tail -n 15 gpurun_out/all_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 64 --warmup 4 > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench.log | cut -c1-3000
SWEEP_SHAPES=1280x51904 timeout 600 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_logits.txt 2>&1; tail -n 12 gpurun_out/gemm_sweep_logits.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmha_generation -c 3 -o gpurun_out/prof_mmha_r2 -f python bench.py --profile > gpurun_out/ncu_mmha_r2.log 2>&1; echo "ncu exit $?"
