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


def test_decoder_step_through_the_plugin_classes_matches_the_fused_path():
    """runtime.PluginDecoderStep: every Linear through WeightOnlyQuantMatmulPlugin::enqueue, self-attention through
    GPTAttentionPlugin::enqueue, unfused LayerNorm / bias / GELU / residual layers in between (the reference's graph)
    -- against runtime.WhisperDecoding's fused kernels on the same weights, caches and tokens."""
    from b200_whisper.runtime import PluginDecoderStep, WhisperDecoding
    from oracle import whisper_oracle as wo
    dims = wo.ModelDimensions(80, 1500, 1280, 20, 1, 51865, 448, 1280, 20, 2)
    B, prompt, n_new = 4, [50258, 50259, 50359], 5
    sd = wo.synthetic_state_dict(dims, seed=5, decoder_only=True)
    L = dims.n_text_layer
    g = torch.Generator(device="cuda").manual_seed(7)
    cross = [torch.randint(-127, 128, (B, 2, dims.n_text_head, 200, 64), generator=g, device="cuda", dtype=torch.int8)
             for _ in range(L)]

    def run(through_plugins):
        dec = WhisperDecoding(dims, sd, B, [0.05] * L, [0.03] * L, n_audio_ctx=200)
        dec.set_cross_kv([c.clone() for c in cross])
        dec.reset()
        dec.prefill([prompt] * B)
        stepper = PluginDecoderStep(dec) if through_plugins else None
        toks, logits = [], []
        for _ in range(n_new):
            toks.append((stepper.step() if stepper else (dec._step_body() or dec.next_tokens)).clone())
            if not through_plugins:
                dec._host_len += 1
            logits.append(dec.logits.clone())
        n_enq = stepper.enqueues if stepper else 0
        if stepper:
            stepper.close()
        return torch.stack(toks, 1).cpu(), torch.stack(logits, 1).float().cpu(), n_enq

    t_fused, l_fused, _ = run(False)
    t_plug, l_plug, n_enq = run(True)
    assert n_enq == n_new * L * 7            # six matmul plugin enqueues + one attention plugin enqueue per layer
    scale = l_fused.abs().max().item()
    # same arithmetic up to the rounding points of the fused epilogues / folded LayerNorm (numerics)
    assert (l_plug - l_fused).abs().max().item() <= 2e-2 * scale
    margins = l_fused.topk(2, dim=-1).values
    clear = (margins[..., 0] - margins[..., 1]) > 0.05 * scale
    same_history = torch.cat([torch.ones(B, 1, dtype=torch.bool), (t_plug == t_fused).cumprod(1)[:, :-1].bool()], 1)
    assert ((t_plug == t_fused) | ~clear | ~same_history).all()
    assert (t_plug[:, 0] == t_fused[:, 0]).all()


def test_plugin_step_captured_in_a_cuda_graph_equals_the_eager_enqueues():
    """PluginDecoderStep.capture(): the plugin enqueues of a step recorded in one CUDA graph (how TensorRT captures an
    execution context) and replayed from host token buffers -- same tokens and logits, bit for bit, as enqueueing every
    step eagerly, over several steps (the lengths advance on the device, not in the captured host scalars)."""
    from b200_whisper.runtime import PluginDecoderStep, WhisperDecoding
    from oracle import whisper_oracle as wo
    dims = wo.ModelDimensions(80, 1500, 1280, 20, 1, 51865, 448, 1280, 20, 2)
    B, prompt, n_new = 4, [50258, 50259, 50359], 6
    sd = wo.synthetic_state_dict(dims, seed=5, decoder_only=True)
    L = dims.n_text_layer
    g = torch.Generator(device="cuda").manual_seed(7)
    cross = [torch.randint(-127, 128, (B, 2, dims.n_text_head, 200, 64), generator=g, device="cuda", dtype=torch.int8)
             for _ in range(L)]

    def run(graph):
        dec = WhisperDecoding(dims, sd, B, [0.05] * L, [0.03] * L, n_audio_ctx=200)
        dec.set_cross_kv([c.clone() for c in cross])
        dec.reset()
        host = dec.prefill([prompt] * B).cpu().numpy().copy()
        stepper = PluginDecoderStep(dec)
        if graph:
            stepper.capture()
        toks, logits = [], []
        for _ in range(n_new):
            host = (stepper.step_host_graph(host) if graph else stepper.step_host(host)).numpy().copy()
            toks.append(torch.from_numpy(host.copy()))
            logits.append(dec.logits.clone())
        n_enq = stepper.enqueues
        stepper.close()
        return torch.stack(toks, 1), torch.stack(logits, 1).cpu(), n_enq, dec.seq_len.cpu()

    t_e, l_e, n_e, len_e = run(False)
    t_g, l_g, n_g, len_g = run(True)
    assert torch.equal(t_e, t_g)
    assert torch.equal(l_e, l_g)
    assert torch.equal(len_e, len_g)
    assert n_e == n_g == n_new * L * 7


def test_two_threads_enqueue_concurrently():
    """Two plugin instances enqueueing from two host threads on two streams (two execution contexts of an engine): the
    library's switches are thread-local and its counter / scratch slots are per stream, so neither thread can disturb
    the other.  One thread additionally flips the test-only switches the whole time."""
    import threading

    import b200_whisper as bw
    from b200_whisper import _lib
    lib = _lib.load()
    m, n, k = 16, 1280, 5120                     # split-K decode shape: uses the arrival counters / cluster exchange
    torch.manual_seed(0)
    weight = gen((k, n), 3)
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(weight, torch.int8)
    w, s = proc.cuda().view(torch.float32), scales.cuda()
    xs = [(gen((m, k), 10 + t) * 4).cuda().view(1, m, k) for t in range(2)]
    refs = []
    plug0 = TrtPlugin.create("WeightOnlyQuantMatmul", woq_fields())
    ins = [((1, m, k), "float16"), (tuple(w.shape), "float32"), (tuple(s.shape), "float16")]
    outs = [((1, m, n), "float16")]
    ws0 = torch.empty((max(plug0.workspace_size(ins, outs), 16),), dtype=torch.uint8, device="cuda")
    for x in xs:                                  # single-threaded answers first
        o = torch.empty((1, m, n), dtype=torch.float16, device="cuda")
        assert plug0.enqueue(ins, outs, [x.data_ptr(), w.data_ptr(), s.data_ptr()], [o.data_ptr()], ws0.data_ptr(),
                             torch.cuda.current_stream().cuda_stream) == 0
        torch.cuda.synchronize()
        refs.append(o.clone())
    plug0.destroy()
    errors = []

    def worker(t):
        try:
            plug = TrtPlugin.create("WeightOnlyQuantMatmul", woq_fields())
            stream = torch.cuda.Stream()
            ws = torch.empty((ws0.numel(),), dtype=torch.uint8, device="cuda")
            o = torch.empty((1, m, n), dtype=torch.float16, device="cuda")
            with torch.cuda.stream(stream):
                for it in range(200):
                    if t == 1:                    # thread-local: must not leak into thread 0's launches
                        lib.b200_set_static_kv_hint(it & 1)
                        lib.b200_woq_set_kernel_policy(2 if it & 1 else 0)
                    rc = plug.enqueue(ins, outs, [xs[t].data_ptr(), w.data_ptr(), s.data_ptr()], [o.data_ptr()],
                                      ws.data_ptr(), stream.cuda_stream)
                    if rc != 0:
                        raise RuntimeError(f"enqueue rc {rc}")
                    if it % 50 == 49:
                        stream.synchronize()
                        if not torch.equal(o, refs[t]):
                            raise AssertionError(f"thread {t}, iteration {it}: output differs from the single-threaded run")
            stream.synchronize()
            plug.destroy()
            if t == 1:
                lib.b200_set_static_kv_hint(0)
                lib.b200_woq_set_kernel_policy(0)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
