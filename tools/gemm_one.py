"""One weight-only GEMM launch of a given shape (for ncu captures): python tools/gemm_one.py M K N [gelu]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

import b200_whisper as bw
from b200_whisper import _lib

m, k, n = (int(a) for a in sys.argv[1:4])
lib = _lib.load()
dev = torch.device("cuda")
w = ((torch.rand((k, n), device=dev) * 2 - 1) * 0.05).half()
p, s = bw.ops.symmetric_quantize_last_axis_of_batched_matrix(w, torch.int8)
x = (torch.rand((m, k), device=dev) * 2 - 1).half()
o = torch.empty((m, n), dtype=torch.float16, device=dev)
wk = torch.empty((lib.b200_woq_workspace_bytes(m, n, k),), dtype=torch.uint8, device=dev)
gelu = len(sys.argv) > 4 and sys.argv[4] == "gelu"
bias = (torch.rand((n,), device=dev) * 0.1).half()
for _ in range(3):
    if gelu:  # the fc1 launch of the encoder: bias + erf GELU in the epilogue
        lib.b200_woq_int8_gemm_fused(x.data_ptr(), m, k, p.data_ptr(), s.data_ptr(), n, bias.data_ptr(), _lib.ACT_GELU_ERF, None,
                                     o.data_ptr(), wk.data_ptr(), wk.numel(), torch.cuda.current_stream().cuda_stream)
    else:
        lib.b200_woq_int8_gemm(x.data_ptr(), m, k, p.data_ptr(), s.data_ptr(), n, o.data_ptr(), wk.data_ptr(), wk.numel(),
                               torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("done")
