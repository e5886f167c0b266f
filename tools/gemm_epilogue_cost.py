"""Cost of the fused epilogues of the large-M weight-only GEMM: python tools/gemm_epilogue_cost.py [M K N]
Times plain / +bias / +bias+GELU / +bias+residual(in place) with CUDA events; VARIANT=<name> runs only that one
(for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

import b200_whisper as bw
from b200_whisper import _lib

m, k, n = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (24000, 1280, 5120)
lib = _lib.load()
dev = torch.device("cuda")
w = ((torch.rand((k, n), device=dev) * 2 - 1) * 0.05).half()
p, s = bw.ops.symmetric_quantize_last_axis_of_batched_matrix(w, torch.int8)
x = (torch.rand((m, k), device=dev) * 2 - 1).half()
bias = (torch.rand((n,), device=dev) - 0.5).half()
o = torch.zeros((m, n), dtype=torch.float16, device=dev)
wk = torch.empty((max(lib.b200_woq_workspace_bytes(m, n, k), 1 << 20),), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
variants = {
    "plain": (None, _lib.ACT_NONE, None),
    "bias": (bias, _lib.ACT_NONE, None),
    "bias+gelu": (bias, _lib.ACT_GELU_ERF, None),
    "bias+residual(in place)": (bias, _lib.ACT_NONE, o),
}
only = os.environ.get("VARIANT")
for name, (b, act, res) in variants.items():
    if only and name != only:
        continue

    def run():
        _lib.check(lib.b200_woq_int8_gemm_fused(x.data_ptr(), m, k, p.data_ptr(), s.data_ptr(), n,
                                                b.data_ptr() if b is not None else None, act,
                                                res.data_ptr() if res is not None else None, o.data_ptr(), wk.data_ptr(),
                                                wk.numel(), st))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    print(f"{k}->{n} M={m} {name:26s} {t * 1e3:8.1f} us  {2.0 * m * n * k / t / 1e9:7.1f} TFLOP/s")
