#!/bin/bash
# Round-end rehearsal: the driver's own commands (whole GPU suite in ONE process, smoke, bench), then ncu captures of the
# kernels added or retuned last (log-Mel front end, conv stem with the 256-step tile).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
: > gpurun_out/summary.txt
echo "=== pytest -m gpu (one process)" | tee -a gpurun_out/summary.txt
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/all_gpu.log 2>&1; echo "exit $?" >> gpurun_out/all_gpu.log
tail -n 6 gpurun_out/all_gpu.log | tee -a gpurun_out/summary.txt
echo "=== smoke" | tee -a gpurun_out/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
tail -n 3 gpurun_out/smoke.log | tee -a gpurun_out/summary.txt
echo "=== bench" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py $BENCH_ARGS > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
tail -n 3 gpurun_out/bench.log | cut -c1-6000 | tee -a gpurun_out/summary.txt
if [ "$1" != "noncu" ]; then
echo "=== ncu" | tee -a gpurun_out/summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:log_mel_power_kernel -s 3 -c 1 -f \
   -o gpurun_out/prof_log_mel python tools/log_mel_bench.py > gpurun_out/ncu_log_mel.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc_kernel -s 3 -c 1 -f \
   -o gpurun_out/prof_conv2_nt256 python tools/conv_bench.py > gpurun_out/ncu_conv2_nt256.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
fi
