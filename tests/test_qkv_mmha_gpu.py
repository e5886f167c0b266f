"""b200_qkv_mmha_decode (csrc/qkv_mmha.cu: LayerNorm + qkv projection + masked self-attention of the generation step as ONE
kernel, a thread-block cluster per head) against the two operators it replaces -- b200_woq_int8_gemm_ln_folded (the
weight-only matmul plugin with folded LayerNorm and bias) followed by b200_mmha_generation (the GPTAttention plugin's
generation kernel; reference: weightOnlyQuantMatmulPlugin.cpp:162-222 + gptAttentionCommon.cpp:649-780) -- on the same
inputs and the same int8 KV cache: attention output within fp16 noise, appended cache rows identical up to one
quantization step, nothing else in the cache touched."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(H, B, Smax, seed):
    from b200_whisper import _lib
    from b200_whisper.runtime.whisper_decoding import _QLinear
    lib = _lib.load()
    _lib.check(lib.b200_init(), "init")
    d = H * 64
    g = torch.Generator().manual_seed(seed)
    w = torch.randn((3 * d, d), generator=g) * d ** -0.5
    bias = torch.randn((3 * d,), generator=g) * 0.1
    bias[d:2 * d] = 0                                     # the key projection has no bias (weight.py:221-226)
    lin = _QLinear(w, bias, "cuda")
    gamma = (1.0 + 0.1 * torch.randn((d,), generator=g)).half().cuda()
    beta = (0.1 * torch.randn((d,), generator=g)).half().cuda()
    lin.fold_layernorm(lib, gamma, beta, torch.cuda.current_stream().cuda_stream)
    x = (torch.randn((B, d), generator=g) * 1.5 + 0.3).half().cuda()   # rows with an offset: exercises the LayerNorm statistics
    cache = torch.randint(-127, 128, (B, 2, H, Smax, 64), generator=g, dtype=torch.int8).cuda()
    return lib, lin, gamma, beta, x, cache


@pytest.mark.parametrize("H,B,past", [(20, 16, 37), (20, 16, 0), (20, 16, 200), (20, 16, 447), (20, 5, 9), (20, 1, 64),
                                      (2, 2, 12), (6, 3, 70), (12, 16, 33), (16, 4, 17)])
def test_fused_qkv_attention_equals_the_two_operators(H, B, past):
    from b200_whisper import _lib
    Smax, d = 448, H * 64
    lib, lin, gamma, beta, x, cache0 = _setup(H, B, Smax, seed=H * 1000 + B * 10 + past)
    assert lib.b200_qkv_mmha_decode_supported(B, H, 64) == 1
    st = torch.cuda.current_stream().cuda_stream
    # per-row lengths: the headline case steps all rows together, but the kernel takes a length per sequence
    seq = torch.full((B,), past, dtype=torch.int32, device="cuda")
    if B > 2 and past > 3:
        seq[1] = past - 3
    oq = torch.tensor([1.0 / 0.04], dtype=torch.float32, device="cuda")
    qo = torch.tensor([0.04], dtype=torch.float32, device="cuda")

    # reference: the two operators
    c_ref = cache0.clone()
    qkv = torch.empty((B, 3 * d), dtype=torch.float16, device="cuda")
    ws = torch.empty((max(lib.b200_woq_workspace_bytes(B, 3 * d, d), 1 << 20),), dtype=torch.uint8, device="cuda")
    _lib.check(lib.b200_woq_int8_gemm_ln_folded(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), lin.c1s.data_ptr(),
                                                lin.c2.data_ptr(), 1e-5, B, d, lin.weight.data_ptr(), lin.scales.data_ptr(),
                                                3 * d, lin.bias.data_ptr(), 0, None, qkv.data_ptr(), ws.data_ptr(), ws.numel(), st))
    o_ref = torch.empty((B, d), dtype=torch.float16, device="cuda")
    p = _lib.MmhaParams()
    p.qkv, p.qkv_bias, p.out = qkv.data_ptr(), None, o_ref.data_ptr()
    p.kv_cache, p.sequence_lengths, p.masked_tokens = c_ref.data_ptr(), seq.data_ptr(), None
    p.kv_scale_orig_quant, p.kv_scale_quant_orig = oq.data_ptr(), qo.data_ptr()
    p.batch_size, p.num_heads, p.head_size = B, H, 64
    p.max_seq_len, p.past_kv_length, p.int8_kv_cache, p.q_scaling = Smax, 0, 1, 1.0
    _lib.check(lib.b200_mmha_generation(ctypes.byref(p), st))

    # fused kernel
    c_new = cache0.clone()
    o_new = torch.empty((B, d), dtype=torch.float16, device="cuda")
    _lib.check(lib.b200_qkv_mmha_decode(x.data_ptr(), gamma.data_ptr(), lin.c1s.data_ptr(), lin.c2.data_ptr(), 1e-5,
                                        lin.weight.data_ptr(), lin.scales.data_ptr(), lin.bias.data_ptr(), c_new.data_ptr(),
                                        seq.data_ptr(), oq.data_ptr(), qo.data_ptr(), o_new.data_ptr(), B, H, 64, Smax, st))
    torch.cuda.synchronize()
    assert torch.isfinite(o_new.float()).all()
    scale = max(1.0, o_ref.float().abs().max().item())
    err = (o_new.float() - o_ref.float()).abs().max().item()
    assert err <= 4e-3 * scale, f"attention output differs by {err} (scale {scale})"
    # the appended rows: same quantization rule on (nearly) the same k / v -> at most one step apart, rarely
    dc = (c_new.int() - c_ref.int()).abs()
    assert dc.max().item() <= 1
    assert (dc != 0).float().mean().item() < 1e-3
    # nothing but row seq[b] of every (b, K|V, head) changed
    touched = (c_new != cache0)
    for b in range(B):
        touched[b, :, :, int(seq[b].item())] = False
    assert not touched.any()


def test_unsupported_shapes_are_refused():
    import b200_whisper
    lib = b200_whisper.load()
    assert lib.b200_qkv_mmha_decode_supported(17, 20, 64) == 0      # more than 16 rows
    assert lib.b200_qkv_mmha_decode_supported(16, 20, 128) == 0     # head size
    assert lib.b200_qkv_mmha_decode_supported(16, 28, 64) == 0      # 28 k-blocks / 4 = 7 per CTA: not instantiated
    assert lib.b200_qkv_mmha_decode_supported(16, 5, 64) == 0       # odd number of heads: no cluster split
    assert lib.b200_qkv_mmha_decode_supported(16, 20, 64) == 1


@pytest.mark.parametrize("dims_name,B", [("micro", 2), ("large-v2 width, 2 layers", 16)])
def test_decoder_with_the_fused_kernel_matches_the_operator_chain(dims_name, B):
    """The whole generation step with fuse_qkv_mmha on (one launch fewer per layer) against the default chain: same cache
    rows, same tokens wherever the decision is clear, logits within fp16 noise."""
    from b200_whisper.runtime import WhisperDecoding
    from oracle import whisper_oracle as wo
    dims = wo.MICRO if dims_name == "micro" else wo.ModelDimensions(80, 1500, 1280, 20, 2, 51865, 448, 1280, 20, 2)
    sd = wo.synthetic_state_dict(dims, seed=9, decoder_only=True)
    L = dims.n_text_layer
    torch.manual_seed(3)
    xa = torch.randn(B, 200 if dims_name != "micro" else dims.n_audio_ctx, dims.n_text_state).half().cuda()
    outs = []
    for fused in (True, False):
        dec = WhisperDecoding(dims, sd, B, [0.04] * L, [0.03] * L, n_audio_ctx=xa.shape[1])
        dec.fuse_qkv_mmha = fused
        dec.set_encoder_output(xa)
        dec.reset()
        toks = [dec.prefill([[3, 7, 11]] * B).clone()]
        logits = [dec.logits.clone()]
        for _ in range(5):
            dec._step_body()
            toks.append(dec.next_tokens.clone())
            logits.append(dec.logits.clone())
        torch.cuda.synchronize()
        outs.append((torch.stack(toks, 1), torch.stack(logits, 1), [c.clone() for c in dec.self_kv]))
    (t_new, l_new, c_new), (t_old, l_old, c_old) = outs
    scale = l_old.abs().max().item()
    same = torch.ones(t_new.shape, dtype=torch.bool, device=t_new.device)
    for b in range(B):
        d = (t_new[b] != t_old[b]).nonzero()
        if len(d):
            same[b, int(d[0]) + 1:] = False
            top2 = l_old[b, int(d[0])].topk(2).values
            assert (top2[0] - top2[1]).item() <= 1e-2 * scale, "tokens diverged on a clear decision"
    assert same[:, :2].all()
    assert ((l_new - l_old).abs().amax(-1) * same).max().item() <= 4e-3 * scale
    if bool(same.all()):
        for a, b in zip(c_new, c_old):
            assert (a.int() - b.int()).abs().max().item() <= 1
