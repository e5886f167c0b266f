#!/bin/bash
# configs[3] (64 utterances sharded) + weak scaling on N GPUs of one box: one bench run per call
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python __graft_entry__.py build > gpurun_out/build.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 64 --warmup 4 --no-cpu-baseline > gpurun_out/bench_n$N.log 2>&1
echo "exit $?"; grep '^{' gpurun_out/bench_n$N.log | tail -1 > gpurun_out/bench_n$N.json; python - <<PY
import json
d = json.load(open("gpurun_out/bench_n$N.json"))
print("N=%d weak: %.0f tok/s (%.3f ms/step, batch 16 per GPU); strong (64 utterances): %s" % (d["n_gpus"], d["value"], d["ms_per_step"], json.dumps(d.get("strong_scaling"))))
PY
