#!/bin/bash
# A/B of environment-variable variants of the working tree on one box, interleaved twice.
# usage: tools/ab_env.sh "" "B200_CLUSTER16=1" "B200_XA_CFG=F" ...
cd "$(dirname "$0")/.."
python __graft_entry__.py build > /dev/null 2>&1
for rep in 1 2; do
  for v in "$@"; do
    echo "[${v:-default}] $(env $v timeout 240 python bench.py --steps 64 --warmup 4 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*')"
  done
done
