#!/bin/bash
# A/B of whole-repo snapshots under ab/<name>/ against the working tree, same box, interleaved twice.
cd "$(dirname "$0")/.."
ROOT=$PWD
for v in "$@"; do (cd ab/$v && python __graft_entry__.py build > /dev/null 2>&1); done
python __graft_entry__.py build > /dev/null 2>&1
for rep in 1 2; do
  for v in "$@" cur; do
    if [ $v = cur ]; then d=$ROOT; else d=$ROOT/ab/$v; fi
    echo "$v: $(cd $d && timeout 240 python bench.py --steps 64 --warmup 4 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*')"
  done
done
