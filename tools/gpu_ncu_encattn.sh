#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python __graft_entry__.py build > gpurun_out/build.log 2>&1
cat > /tmp/enc_attn_one.py <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
from b200_whisper.functional import bidirectional_attention
torch.manual_seed(0)
qkv = (torch.randn((16, 1500, 3 * 20 * 64), device="cuda") * 1.2).half()
for _ in range(3):
    out = bidirectional_attention(qkv, 20, 64)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = bidirectional_attention(qkv, 20, 64)
e1.record(); torch.cuda.synchronize()
print("attention B=16 S=1500 H=20: %.1f us per launch" % (e0.elapsed_time(e1) * 100))
PY
python /tmp/enc_attn_one.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_bidir_tc -c 1 -o gpurun_out/prof_encattn_r2 -f python /tmp/enc_attn_one.py > gpurun_out/ncu_encattn_r2.log 2>&1; echo "ncu exit $?"
