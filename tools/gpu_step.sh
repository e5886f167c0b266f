#!/bin/bash
# Step-kernel session: step-kernel tests, (optional) phase timeline with the debug build, bench of both paths.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python __graft_entry__.py build > gpurun_out/build.log 2>&1; tail -n 2 gpurun_out/build.log
if [ "${1:-tests}" != "notests" ]; then
timeout 900 python -m pytest tests/test_decoder_step_gpu.py -x -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/step_tests.log 2>&1; echo "pytest exit $?"
tail -n 12 gpurun_out/step_tests.log
fi
B200_STEP_KERNEL=1 timeout 600 python bench.py --steps 64 --warmup 4 --no-cpu-baseline --no-extras > gpurun_out/bench_step.log 2>&1
echo "step kernel: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_step.log | head -1) $(tail -n 2 gpurun_out/bench_step.log | grep -v '^{' | cut -c1-300)"
B200_DS_DEBUG=1 python -m b200_whisper._build > gpurun_out/dbg_build.log 2>&1
B200_STEP_KERNEL=1 timeout 300 python tools/step_phases.py > gpurun_out/step_phases.txt 2>&1; cat gpurun_out/step_phases.txt | cut -c1-250
