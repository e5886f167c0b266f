"""SIMT GEMV vs tcgen05 GEMM at small M on the decoder shapes: CUDA-graph replays over 32 distinct weight matrices."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

import b200_whisper as bw
from b200_whisper import _lib

lib = _lib.load()
dev = torch.device("cuda")
L = 32


def graph_ms(body, reps=5):
    cur = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        body()
    cur.wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for k, n in ((1280, 3840), (1280, 1280), (1280, 5120), (5120, 1280)):
    ws_ = []
    for i in range(L):
        w = ((torch.rand((k, n), device=dev) * 2 - 1) * 0.05).half()
        ws_.append(bw.ops.symmetric_quantize_last_axis_of_batched_matrix(w, torch.int8))
    line = f"K={k:5d} N={n:5d}:"
    for m in (1, 2, 4, 8):
        x = (torch.rand((m, k), device=dev) * 2 - 1).half()
        o = torch.empty((m, n), dtype=torch.float16, device=dev)
        wk = torch.empty((lib.b200_woq_workspace_bytes(m, n, k),), dtype=torch.uint8, device=dev)

        def run():
            st = torch.cuda.current_stream().cuda_stream
            for p, s in ws_:
                lib.b200_woq_int8_gemm(x.data_ptr(), m, k, p.data_ptr(), s.data_ptr(), n, o.data_ptr(), wk.data_ptr(),
                                       wk.numel(), st)
        res = []
        for policy in (1, 2):
            if policy == 1 and m > 4:
                res.append(float("nan"))
                continue
            lib.b200_woq_set_kernel_policy(policy)
            res.append(1e3 * graph_ms(run) / L)
        lib.b200_woq_set_kernel_policy(0)
        line += f"  M={m}: simt {res[0]:6.2f} tc {res[1]:6.2f} us"
    print(line)
