"""BASELINE config 5: isolated fp16 x int8 GEMM sweep at the large-v2 shapes, M = 1 .. 256 (+ the encoder's M = 1500 and
24000), CUDA-graph replays over distinct weight matrices (no L2 reuse of weights), CUDA events."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

import b200_whisper as bw
from b200_whisper import _lib

lib = _lib.load()
dev = torch.device("cuda")


def graph_ms(body, reps=3):
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


print(f"{'K->N':>12s} {'M':>6s} {'us/launch':>10s} {'weight GB/s':>12s} {'TFLOP/s':>9s}")
# the fourth shape of config 5 is the logits projection 1280 -> 51865, padded to 51904 (a multiple of 64) for the int8 layout
SHAPES = [tuple(int(v) for v in s.split("x")) for s in os.environ.get("SWEEP_SHAPES", "1280x3840,1280x5120,5120x1280,1280x51904").split(",")]
for k, n in SHAPES:
    L = 16 if k * n < (32 << 20) else 4   # distinct weight sets per graph: > L2 in total either way
    if n > 6000 and "SWEEP_M" not in os.environ:
        os.environ["SWEEP_M_"] = "1,2,4,8,16,32,64,128,256"
    ws_ = []
    for i in range(L):
        w = ((torch.rand((k, n), device=dev) * 2 - 1) * 0.05).half()
        ws_.append(bw.ops.symmetric_quantize_last_axis_of_batched_matrix(w, torch.int8))
    for m in [int(v) for v in os.environ.get("SWEEP_M", os.environ.get("SWEEP_M_", "1,2,4,8,16,32,64,128,256,1500,24000")).split(",")]:
        x = (torch.rand((m, k), device=dev) * 2 - 1).half()
        o = torch.empty((m, n), dtype=torch.float16, device=dev)
        wk = torch.empty((lib.b200_woq_workspace_bytes(m, n, k),), dtype=torch.uint8, device=dev)

        def run():
            st = torch.cuda.current_stream().cuda_stream
            for p, s in ws_:
                lib.b200_woq_int8_gemm(x.data_ptr(), m, k, p.data_ptr(), s.data_ptr(), n, o.data_ptr(), wk.data_ptr(),
                                       wk.numel(), st)
        t = graph_ms(run) / L
        print(f"{k:5d}->{n:5d} {m:6d} {1e3 * t:10.2f} {k * n / t / 1e6:12.1f} {2.0 * m * n * k / t / 1e9:9.1f}")
