"""GPU tests through the TensorRT plugin classes (registry -> creator -> createPlugin -> enqueue), the way TensorRT
would call them: the reference's test_weight_only_quant_matmul / test_gpt_attention flows without an engine."""
import math

import numpy as np
import pytest
import torch

from b200_whisper.plugin import TrtPlugin
from tests.test_plugin_cpu import attn_fields, woq_fields

pytestmark = pytest.mark.gpu


def gen(shape, seed=0):
    torch.manual_seed(seed)
    return torch.rand(shape, dtype=torch.float16) * 2 - 1.0


@pytest.mark.parametrize("m,n,k", [(1, 1024, 4096), (128, 6144, 12288), (16, 3840, 1280)])
def test_woq_matmul_through_plugin(m, n, k):
    import b200_whisper as bw
    from tests.test_woq_matmul_gpu import colwise_near, tight_check
    mat1 = gen((m, k)) * 200.0
    weight = gen((k, n))
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(weight, torch.int8)
    x = mat1.cuda().view(1, m, k)
    w = proc.cuda().view(torch.float32)  # [K, N/4] float32 view, as examples/whisper/weight.py:79-80
    s = scales.cuda()
    plug = TrtPlugin.create("WeightOnlyQuantMatmul", woq_fields())
    # round-trip through serialization first, as an engine load would
    plug = TrtPlugin.deserialize("WeightOnlyQuantMatmul", plug.serialize())
    ins = [(tuple(x.shape), "float16"), (tuple(w.shape), "float32"), (tuple(s.shape), "float16")]
    out_shape = plug.output_dims(0, ins)
    assert out_shape == (1, m, n)
    outs = [(out_shape, "float16")]
    plug.configure(ins, outs)
    ws_bytes = plug.workspace_size(ins, outs)
    ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device="cuda")
    out = torch.empty(out_shape, dtype=torch.float16, device="cuda")
    rc = plug.enqueue(ins, outs, [x.data_ptr(), w.data_ptr(), s.data_ptr()], [out.data_ptr()], ws.data_ptr(),
                      torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == 0
    ref = (mat1.float() @ raw.float()) * scales.float()[None, :]
    colwise_near(ref.to(torch.float16), out.cpu().view(m, n))
    tight_check(mat1, raw, scales, out.cpu().view(m, n), "plugin")
    plug.destroy()


def test_gpt_attention_through_plugin():
    """Context step + 3 generation steps with an int8 KV cache, I/O convention of test_gpt_attention.py."""
    torch.manual_seed(42)
    B, H, D, in_len = 2, 4, 64, 8
    hidden, max_seq = H * D, 32
    dev = "cuda"
    plug = TrtPlugin.create("GPTAttention", attn_fields(num_heads=H, head_size=D))
    deq = torch.tensor([0.05], dtype=torch.float32, device=dev)
    qnt = 1.0 / deq
    cache = torch.zeros((B, 2, H, max_seq, D), dtype=torch.int8, device=dev)
    input_lengths = torch.full((B,), in_len, dtype=torch.int32, device=dev)
    masked = torch.zeros((B, max_seq), dtype=torch.int32, device=dev)
    cache_ind = torch.zeros((B, 1, max_seq), dtype=torch.int32, device=dev)
    max_in = torch.zeros((in_len,), dtype=torch.int32, device=dev)

    def run(qkv, seq_len, host_scalars):
        S = qkv.shape[1]
        out = torch.empty((B, S, hidden), dtype=torch.float16, device=dev)
        ins = [(tuple(qkv.shape), "float16"), (tuple(cache.shape), "int8"), ((B,), "int32"), ((2,), "int32"),
               ((B, max_seq), "int32"), ((B,), "int32"), ((in_len,), "int32"), ((B, 1, max_seq), "int32"),
               ((1,), "float32"), ((1,), "float32")]
        outs = [(tuple(out.shape), "float16"), (tuple(cache.shape), "int8")]
        assert plug.output_dims(0, ins) == tuple(out.shape)
        host = np.array(host_scalars, np.int32)  # HOST tensor [past_len, is_context]
        rc = plug.enqueue(ins, outs, [qkv.data_ptr(), cache.data_ptr(), seq_len.data_ptr(), host.ctypes.data,
                                      masked.data_ptr(), input_lengths.data_ptr(), max_in.data_ptr(),
                                      cache_ind.data_ptr(), qnt.data_ptr(), deq.data_ptr()],
                          [out.data_ptr(), cache.data_ptr()], None, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert rc == 0
        return out

    def heads(x):
        return x.view(B, -1, H, D).permute(0, 2, 1, 3).float()

    qkv = torch.randn((B, in_len, 3 * hidden), device=dev).half()
    out = run(qkv, torch.full((B,), in_len, dtype=torch.int32, device=dev), [0, 1])
    q, k, v = [heads(t) for t in qkv.float().split(hidden, dim=-1)]
    s = (q @ k.transpose(-1, -2)) / math.sqrt(D)
    s = s.masked_fill(~torch.tril(torch.ones(in_len, in_len, dtype=torch.bool, device=dev)), float("-inf"))
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B, in_len, hidden)
    assert (out.float() - ref).abs().max().item() <= 5e-3
    past_k = (cache[:, 0, :, :in_len].float() * deq).half().float()
    past_v = (cache[:, 1, :, :in_len].float() * deq).half().float()
    for step in range(1, 4):
        past_len = in_len + step - 1
        qkv1 = torch.randn((B, 1, 3 * hidden), device=dev).half()
        out = run(qkv1, torch.full((B,), past_len, dtype=torch.int32, device=dev), [past_len, 0])
        q1, k1, v1 = [heads(t) for t in qkv1.float().split(hidden, dim=-1)]
        k_all, v_all = torch.cat([past_k, k1], 2), torch.cat([past_v, v1], 2)
        ref = (torch.softmax((q1 @ k_all.transpose(-1, -2)) / math.sqrt(D), -1) @ v_all).permute(0, 2, 1, 3).reshape(B, 1, hidden)
        assert (out.float() - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())
        past_k = torch.cat([past_k, (cache[:, 0, :, past_len].float() * deq).half().float()[:, :, None]], 2)
        past_v = torch.cat([past_v, (cache[:, 1, :, past_len].float() * deq).half().float()[:, :, None]], 2)
    plug.destroy()


def test_gpt_attention_paged_kv_through_plugin():
    """paged_kv_cache = 1 (gptAttentionPlugin.cpp:314-326): input 1 is the block pool, input 10 the block pointers as pairs
    of int32; the plugin must produce the same bits as the contiguous-cache plugin."""
    torch.manual_seed(3)
    B, H, D, in_len = 2, 4, 64, 8
    hidden, max_seq, tpb = H * D, 32, 8
    max_blocks = max_seq // tpb
    dev = "cuda"
    lin = TrtPlugin.create("GPTAttention", attn_fields(num_heads=H, head_size=D))
    pag = TrtPlugin.create("GPTAttention", attn_fields(num_heads=H, head_size=D, paged_kv_cache=1))
    pag = TrtPlugin.deserialize("GPTAttention", pag.serialize())
    deq = torch.tensor([0.05], dtype=torch.float32, device=dev)
    qnt = 1.0 / deq
    cache = torch.zeros((B, 2, H, max_seq, D), dtype=torch.int8, device=dev)
    n_pool = B * 2 * max_blocks
    pool = torch.zeros((n_pool, 2, H, tpb, D), dtype=torch.int8, device=dev)
    ids = torch.randperm(n_pool).view(B, 1, 2, max_blocks)
    ptrs64 = (pool.data_ptr() + ids * (2 * H * tpb * D)).to(dev)               # each table entry -> pool[id, 0]
    ptrs32 = ptrs64.view(torch.int32)                                          # [B, 1, 2, 2 * max_blocks]
    assert tuple(ptrs32.shape) == (B, 1, 2, 2 * max_blocks)
    input_lengths = torch.full((B,), in_len, dtype=torch.int32, device=dev)
    masked = torch.zeros((B, max_seq), dtype=torch.int32, device=dev)
    cache_ind = torch.zeros((B, 1, max_seq), dtype=torch.int32, device=dev)
    max_in = torch.zeros((in_len,), dtype=torch.int32, device=dev)

    def run(plug, kv, qkv, seq_len, host_scalars, paged):
        S = qkv.shape[1]
        out = torch.empty((B, S, hidden), dtype=torch.float16, device=dev)
        ins = [(tuple(qkv.shape), "float16"), (tuple(kv.shape), "int8"), ((B,), "int32"), ((2,), "int32"),
               ((B, max_seq), "int32"), ((B,), "int32"), ((in_len,), "int32"), ((B, 1, max_seq), "int32"),
               ((1,), "float32"), ((1,), "float32")]
        ptr_list = [qkv.data_ptr(), kv.data_ptr(), seq_len.data_ptr(), None, masked.data_ptr(), input_lengths.data_ptr(),
                    max_in.data_ptr(), cache_ind.data_ptr(), qnt.data_ptr(), deq.data_ptr()]
        if paged:
            ins.append((tuple(ptrs32.shape), "int32"))
            ptr_list.append(ptrs32.data_ptr())
        outs = [(tuple(out.shape), "float16"), (tuple(kv.shape), "int8")]
        host = np.array(host_scalars, np.int32)
        ptr_list[3] = host.ctypes.data
        rc = plug.enqueue(ins, outs, ptr_list, [out.data_ptr(), kv.data_ptr()], None, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert rc == 0
        return out

    steps = [(torch.randn((B, in_len, 3 * hidden), device=dev).half(), in_len, [0, 1])]
    for step in range(4):
        steps.append((torch.randn((B, 1, 3 * hidden), device=dev).half(), in_len + step, [in_len + step, 0]))
    for qkv, n, host in steps:
        seq_len = torch.full((B,), n, dtype=torch.int32, device=dev)
        a = run(lin, cache, qkv, seq_len, host, False)
        b = run(pag, pool, qkv, seq_len, host, True)
        assert torch.equal(a, b)
    n = in_len + 4
    for bi in range(B):
        for kv in range(2):
            for t in range(n):
                blk = int(ids[bi, 0, kv, t // tpb])
                assert torch.equal(pool[blk, 0, :, t % tpb], cache[bi, kv, :, t])
    lin.destroy()
    pag.destroy()
