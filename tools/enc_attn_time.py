"""Times b200_attention_bidirectional_fp16 at the large-v2 encoder shape (B x 1500 frames x 20 heads x 64)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from b200_whisper.functional import bidirectional_attention  # noqa: E402

B = int(os.environ.get("BATCH", "16"))
torch.manual_seed(0)
qkv = (torch.randn((B, 1500, 3 * 20 * 64), device="cuda") * 1.2).half()
for _ in range(3):
    out = bidirectional_attention(qkv, 20, 64)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = bidirectional_attention(qkv, 20, 64)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 50
flop = 4.0 * B * 20 * 1500 * 1500 * 64
q, k, v = [t.float().view(B, 1500, 20, 64).permute(0, 2, 1, 3)[:1] for t in qkv.split(20 * 64, dim=-1)]
ref = (torch.softmax((q @ k.transpose(-1, -2)) * 0.125, dim=-1) @ v).permute(0, 2, 1, 3).reshape(1, 1500, 1280)
err = (out[:1].float() - ref).abs().max().item()
print(f"attention B={B} S=1500 H=20 ({os.environ.get('B200_ENC_ATTN', 'tcgen05')}): {us:.1f} us per launch, {flop / us / 1e6:.0f} TFLOP/s, "
      f"max err vs fp32 {err:.2e}")
