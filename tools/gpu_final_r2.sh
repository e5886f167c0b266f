#!/bin/bash
# Round-2 final session: whole GPU suite, smoke, bench (+ launch list under ncu), encoder / pipeline benches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python __graft_entry__.py build > gpurun_out/build.log 2>&1; tail -n 2 gpurun_out/build.log
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/all_gpu.log 2>&1; echo "pytest exit $?"
tail -n 6 gpurun_out/all_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 64 --warmup 4 > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench.log | cut -c1-600
timeout 300 python tools/encoder_bench.py > gpurun_out/encoder_bench.txt 2>&1; grep "encoder large\|layer total\|dur " gpurun_out/encoder_bench.txt
timeout 300 python tools/pipeline_bench.py > gpurun_out/pipeline_bench.txt 2>&1; tail -n 8 gpurun_out/pipeline_bench.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile > gpurun_out/ncu_bench.log 2>&1; echo "ncu exit $?"; wc -l gpurun_out/launches.csv
