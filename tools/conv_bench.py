"""Times the Whisper large-v2 conv stem (conv1 80->1280 s1, conv2 1280->1280 s2, GELU fused) for a batch of utterances:
tcgen05 implicit GEMM vs the CUDA-core direct convolution."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

from b200_whisper.functional import conv1d

B = int(os.environ.get("BATCH", "16"))
torch.manual_seed(0)
x = torch.randn((B, 80, 3000), device="cuda").half()
w1 = (torch.randn((1280, 80, 3), device="cuda") / 240 ** 0.5).half()
b1 = torch.zeros(1280, device="cuda").half()
w2 = (torch.randn((1280, 1280, 3), device="cuda") / 3840 ** 0.5).half()
b2 = torch.zeros(1280, device="cuda").half()
for impl in ("tc", "simt"):
    for _ in range(2):
        h = conv1d(x, w1, b1, stride=1, padding=1, activation="gelu", impl=impl)
        y = conv1d(h, w2, b2, stride=2, padding=1, activation="gelu", impl=impl)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    h = conv1d(x, w1, b1, stride=1, padding=1, activation="gelu", impl=impl)
    e[1].record()
    y = conv1d(h, w2, b2, stride=2, padding=1, activation="gelu", impl=impl)
    e[2].record()
    torch.cuda.synchronize()
    t1, t2 = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    f1 = 2 * 1280 * 240 * 3000 * B / 1e12
    f2 = 2 * 1280 * 3840 * 1500 * B / 1e12
    print(f"{impl:5s} batch {B}: conv1 {t1:7.3f} ms ({f1 / t1 * 1e3:7.1f} TFLOP/s)  conv2 {t2:7.3f} ms ({f2 / t2 * 1e3:7.1f} TFLOP/s)"
          f"  (includes the weight re-layout and input transpose of the tc path)")
