"""Times b200_qkv_mmha_decode against b200_woq_int8_gemm_ln_folded + b200_mmha_generation (32 layers' worth of distinct
weights and caches per CUDA-graph replay, batch 16, large-v2 width)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from b200_whisper import _lib  # noqa: E402
from b200_whisper.runtime.whisper_decoding import _QLinear  # noqa: E402

lib = _lib.load()
_lib.check(lib.b200_init())
H, B, Smax, L, past = 20, 16, 448, 32, 37
d = H * 64
g = torch.Generator().manual_seed(0)
lins, caches = [], []
gamma = torch.ones(d).half().cuda()
beta = torch.zeros(d).half().cuda()
for i in range(L):
    lin = _QLinear(torch.randn((3 * d, d), generator=g) * d ** -0.5, torch.zeros(3 * d), "cuda")
    lin.fold_layernorm(lib, gamma, beta, torch.cuda.current_stream().cuda_stream)
    lins.append(lin)
    caches.append(torch.randint(-127, 128, (B, 2, H, Smax, 64), generator=g, dtype=torch.int8).cuda())
x = torch.randn((B, d), generator=g).half().cuda()
seq = torch.full((B,), past, dtype=torch.int32, device="cuda")
oq = torch.tensor([25.0], device="cuda")
qo = torch.tensor([0.04], device="cuda")
qkv = torch.empty((B, 3 * d), dtype=torch.float16, device="cuda")
out = torch.empty((B, d), dtype=torch.float16, device="cuda")
ws = torch.empty((1 << 22,), dtype=torch.uint8, device="cuda")
lib.b200_set_static_kv_hint(1)


def fused():
    st = torch.cuda.current_stream().cuda_stream
    for lin, c in zip(lins, caches):
        lib.b200_qkv_mmha_decode(x.data_ptr(), gamma.data_ptr(), lin.c1s.data_ptr(), lin.c2.data_ptr(), 1e-5, lin.weight.data_ptr(),
                                 lin.scales.data_ptr(), lin.bias.data_ptr(), c.data_ptr(), seq.data_ptr(), oq.data_ptr(), qo.data_ptr(),
                                 out.data_ptr(), B, H, 64, Smax, st)


def two():
    st = torch.cuda.current_stream().cuda_stream
    for lin, c in zip(lins, caches):
        lib.b200_woq_int8_gemm_ln_folded(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), lin.c1s.data_ptr(), lin.c2.data_ptr(), 1e-5,
                                         B, d, lin.weight.data_ptr(), lin.scales.data_ptr(), 3 * d, lin.bias.data_ptr(), 0, None,
                                         qkv.data_ptr(), ws.data_ptr(), ws.numel(), st)
        p = _lib.MmhaParams()
        p.qkv, p.qkv_bias, p.out = qkv.data_ptr(), None, out.data_ptr()
        p.kv_cache, p.sequence_lengths, p.masked_tokens = c.data_ptr(), seq.data_ptr(), None
        p.kv_scale_orig_quant, p.kv_scale_quant_orig = oq.data_ptr(), qo.data_ptr()
        p.batch_size, p.num_heads, p.head_size = B, H, 64
        p.max_seq_len, p.past_kv_length, p.int8_kv_cache, p.q_scaling = Smax, 0, 1, 1.0
        lib.b200_mmha_generation(ctypes.byref(p), st)


def graph_us(body, reps=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        body()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps / L


print(f"qkv projection + self-attention, batch {B}, t = {past}: fused kernel {graph_us(fused):.2f} us per layer, "
      f"two kernels {graph_us(two):.2f} us per layer")
