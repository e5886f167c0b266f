"""Times the REFERENCE's own GEMV and MMHA kernels (oracle/_ref/libref_gpu.so, built for sm_100a from /root/reference)
beside this repo's kernels on the same B200: CUDA-graph replays over 32 distinct weight matrices / caches (no L2 reuse),
CUDA events.  Test infrastructure only."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

import b200_whisper as bw
from b200_whisper import _lib
from oracle import woq

ref = woq.ref_gpu_lib()
assert ref is not None, "oracle/_ref/libref_gpu.so not built"
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


def st():
    return torch.cuda.current_stream().cuda_stream


print("== GEMV, M = 1 (one utterance), per-launch time and weight bandwidth")
for k, n in ((1280, 3840), (1280, 1280), (1280, 5120), (5120, 1280)):
    ws_ = []
    for i in range(L):
        w = ((torch.rand((k, n), device=dev) * 2 - 1) * 0.05).half()
        ws_.append(bw.ops.symmetric_quantize_last_axis_of_batched_matrix(w, torch.int8))
    x = (torch.rand((1, k), device=dev) * 2 - 1).half()
    o = torch.empty((1, n), dtype=torch.float16, device=dev)
    wk = torch.empty((lib.b200_woq_workspace_bytes(1, n, k),), dtype=torch.uint8, device=dev)

    def run_ref():
        for p, s in ws_:
            ref.ref_gpu_gemv(x.data_ptr(), p.data_ptr(), s.data_ptr(), None, o.data_ptr(), k, n, st())

    def run_ours():
        for p, s in ws_:
            lib.b200_woq_int8_gemm(x.data_ptr(), 1, k, p.data_ptr(), s.data_ptr(), n, o.data_ptr(), wk.data_ptr(), wk.numel(), st())
    t_ref, t_our = graph_ms(run_ref) / L, graph_ms(run_ours) / L
    print(f"  K={k:5d} N={n:5d}: reference {1e3 * t_ref:7.2f} us ({k * n / t_ref / 1e6:7.1f} GB/s)   "
          f"this repo {1e3 * t_our:7.2f} us ({k * n / t_our / 1e6:7.1f} GB/s)")

print("== MMHA generation step, int8 KV cache, H = 20, Dh = 64, Smax = 448")
for B, t in ((16, 37), (16, 200), (1, 37)):
    H, D, Smax = 20, 64, 448
    caches = [torch.randint(-127, 128, (B, 2, H, Smax, D), device=dev, dtype=torch.int8) for _ in range(L)]
    qkv = torch.randn((B, 3 * H * D), device=dev).half()
    out = torch.empty((B, H * D), dtype=torch.float16, device=dev)
    seq = torch.full((B,), t, dtype=torch.int32, device=dev)
    zeros = torch.zeros((B,), dtype=torch.int32, device=dev)
    oq = torch.tensor([30.0], dtype=torch.float32, device=dev)
    qo = torch.tensor([1 / 30.0], dtype=torch.float32, device=dev)

    def run_ref():
        for c in caches:
            ref.ref_gpu_mmha(qkv.data_ptr(), out.data_ptr(), c.data_ptr(), seq.data_ptr(), None, zeros.data_ptr(),
                             oq.data_ptr(), qo.data_ptr(), B, H, Smax, t, t, 1, 1.0, st())

    def run_ours():
        for c in caches:
            p = _lib.MmhaParams()
            p.qkv, p.qkv_bias, p.out = qkv.data_ptr(), None, out.data_ptr()
            p.kv_cache = c.data_ptr()
            p.sequence_lengths = seq.data_ptr()
            p.masked_tokens = None
            p.kv_scale_orig_quant, p.kv_scale_quant_orig = oq.data_ptr(), qo.data_ptr()
            p.batch_size, p.num_heads, p.head_size = B, H, D
            p.max_seq_len, p.past_kv_length, p.int8_kv_cache, p.q_scaling = Smax, t, 1, 1.0
            lib.b200_mmha_generation(ctypes.byref(p), st())
    t_ref, t_our = graph_ms(run_ref) / L, graph_ms(run_ours) / L
    print(f"  B={B:2d} t={t:3d}: reference {1e3 * t_ref:7.2f} us   this repo {1e3 * t_our:7.2f} us")
