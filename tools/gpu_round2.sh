#!/bin/bash
# Partial GPU session: selected families + smoke + bench (+ optional ncu launch list).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
: > gpurun_out/summary.txt
python __graft_entry__.py build > gpurun_out/build.log 2>&1; tail -n 2 gpurun_out/build.log | tee -a gpurun_out/summary.txt
run() { local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to python -m pytest -q -m gpu -p no:cacheprovider --timeout 240 "$@" > gpurun_out/$name.log 2>&1
  echo "exit $?" >> gpurun_out/$name.log
  tail -n 3 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
for fam in "$@"; do
  case $fam in
    quant) run quant 600 tests/test_quantize_gpu.py;;
    woq) run woq 900 tests/test_woq_matmul_gpu.py;;
    attention) run attention 600 tests/test_attention_gpu.py;;
    glue) run glue 600 tests/test_glue_conv_gpu.py;;
    decoder) run decoder 900 tests/test_decoder_gpu.py;;
    plugin) run plugin 600 tests/test_plugin_gpu.py;;
    refk) run refk 600 tests/test_reference_kernels_gpu.py;;
    filters) run filters 600 tests/test_logit_filters_gpu.py;;
    encoder) run encoder 600 tests/test_encoder_gpu.py;;
    smoke) echo "=== smoke" | tee -a gpurun_out/summary.txt
       timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
       tail -n 4 gpurun_out/smoke.log | tee -a gpurun_out/summary.txt;;
    bench) echo "=== bench" | tee -a gpurun_out/summary.txt
       timeout 900 python bench.py --steps 64 --warmup 4 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
       tail -n 5 gpurun_out/bench.log | cut -c1-4000 | tee -a gpurun_out/summary.txt;;
    ncu) echo "=== ncu launch list" | tee -a gpurun_out/summary.txt
       timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
          python bench.py --profile > gpurun_out/ncu_bench.log 2>&1
       echo "exit $?" | tee -a gpurun_out/summary.txt; wc -l gpurun_out/launches.csv | tee -a gpurun_out/summary.txt;;
    benchvar) echo "=== bench variants (PDL x cross-KV L2 prefetch)" | tee -a gpurun_out/summary.txt
       IFS=";" read -ra VARS <<< "${BENCH_VARS:-0 C 1 rowhead;0 A 1 rowhead;0 C 1 split;0 E 1 split;0 D 1 split}"; for cfg in "${VARS[@]}"; do
         set -- $cfg
         B200_FUSE_LN=$1 B200_XA_CFG=$2 B200_STATIC_KV=$3 B200_XA_MODE=${4:-auto} B200_CHAINS=${5:-1} timeout 600 python bench.py --steps 64 --warmup 4 --no-cpu-baseline > gpurun_out/bench_var.log 2>&1
         echo "fuse_ln=$1 xa_cfg=$2 static_kv=$3 xa_mode=${4:-auto} chains=${5:-1}: $(grep -o '"value": [0-9.]*' gpurun_out/bench_var.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_var.log | head -1) $(grep -o '"frac": [0-9.]*' gpurun_out/bench_var.log | head -1) $(tail -n 2 gpurun_out/bench_var.log | grep -v '^{' | cut -c1-200)" | tee -a gpurun_out/summary.txt
       done;;
    ncufull) echo "=== ncu full" | tee -a gpurun_out/summary.txt
       timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:cross_attention -c 2 \
          -f -o gpurun_out/prof_xattn python bench.py --profile > gpurun_out/ncu_xattn.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
       timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:woq_gemm_tc_kernel -c 6 \
          -f -o gpurun_out/prof_gemm python bench.py --profile > gpurun_out/ncu_gemm.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt;;
  esac
done
grep -h -E "^(FAILED|ERROR)" gpurun_out/*.log | head -40 | tee -a gpurun_out/summary.txt
