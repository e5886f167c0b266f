"""b200_cross_attention_qproj (csrc/attention.cu, cross_attention_qproj_kernel: the cross-attention of the generation
step WITH its LayerNorm-folded q projection, one launch instead of two) against
  * the two operators it replaces -- b200_woq_int8_gemm_ln_folded (weight-only matmul plugin with the folded
    cross_attention_layernorm; reference weightOnlyQuantMatmulPlugin.cpp:162-222) followed by b200_cross_attention
    (reference graph: T/tensorrt_llm/layers/attention.py:308-323,385-406) -- on the same inputs and the same int8 cross-KV
    cache: output within fp16 noise;
  * a plain fp32 torch evaluation of LayerNorm -> Linear -> softmax(q k^T / 8) v on the int8-round-tripped K / V
    (oracle pin of the cross attention: T/examples/whisper/torch_model.py:88-103)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(H, B, S, seed):
    from b200_whisper import _lib
    from b200_whisper.runtime.whisper_decoding import _QLinear
    lib = _lib.load()
    _lib.check(lib.b200_init(), "init")
    d = H * 64
    g = torch.Generator().manual_seed(seed)
    w = torch.randn((d, d), generator=g) * d ** -0.5
    bias = torch.randn((d,), generator=g) * 0.1
    lin = _QLinear(w, bias, "cuda")
    gamma = (1.0 + 0.1 * torch.randn((d,), generator=g)).half().cuda()
    beta = (0.1 * torch.randn((d,), generator=g)).half().cuda()
    lin.fold_layernorm(lib, gamma, beta, torch.cuda.current_stream().cuda_stream)
    x = (torch.randn((B, d), generator=g) * 1.5 + 0.3).half().cuda()   # rows with an offset: exercises the statistics
    k = torch.randn((B, S, d), generator=g).half().cuda()
    v = torch.randn((B, S, d), generator=g).half().cuda()
    t = 4.5 / 127.0
    oq = torch.tensor([1.0 / t], dtype=torch.float32, device="cuda")
    qo = torch.tensor([t], dtype=torch.float32, device="cuda")
    cache = torch.empty((B, 2, H, S, 64), dtype=torch.int8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.b200_cross_kv_pack(k.data_ptr(), v.data_ptr(), cache.data_ptr(), oq.data_ptr(), B, S, H, 64, 1, st))
    return lib, lin, w, bias, gamma, beta, x, k, v, cache, oq, qo, t


def _two_operators(lib, lin, gamma, beta, x, cache, qo, B, H, S):
    from b200_whisper import _lib
    d = H * 64
    st = torch.cuda.current_stream().cuda_stream
    q = torch.empty((B, d), dtype=torch.float16, device="cuda")
    ws = torch.empty((max(lib.b200_woq_workspace_bytes(B, d, d), lib.b200_cross_attention_workspace_bytes(B, H, 64, S),
                          1 << 20),), dtype=torch.uint8, device="cuda")
    _lib.check(lib.b200_woq_int8_gemm_ln_folded(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), lin.c1s.data_ptr(),
                                                lin.c2.data_ptr(), 1e-5, B, d, lin.weight.data_ptr(), lin.scales.data_ptr(),
                                                d, lin.bias.data_ptr(), 0, None, q.data_ptr(), ws.data_ptr(), ws.numel(), st))
    out = torch.empty((B, d), dtype=torch.float16, device="cuda")
    _lib.check(lib.b200_cross_attention(q.data_ptr(), cache.data_ptr(), qo.data_ptr(), out.data_ptr(), B, 1, H, 64, S, 1,
                                        ws.data_ptr(), ws.numel(), st))
    return q, out


# (H, B, S): the headline shape, ragged last chunk, fewer heads / more rows per CTA, batch above 16
@pytest.mark.parametrize("H,B,S", [(20, 16, 1500), (20, 16, 200), (20, 9, 777), (20, 8, 64), (12, 16, 333), (6, 40, 130),
                                   (16, 24, 1500), (20, 40, 97)])
def test_fused_q_projection_equals_the_two_operators_and_torch(H, B, S):
    from b200_whisper import _lib
    lib, lin, w, bias, gamma, beta, x, k, v, cache, oq, qo, t = _setup(H, B, S, seed=H * 997 + B * 13 + S)
    d = H * 64
    assert lib.b200_cross_attention_qproj_supported(B, H, 64, S) == 1
    q_ref, o_ref = _two_operators(lib, lin, gamma, beta, x, cache, qo, B, H, S)
    st = torch.cuda.current_stream().cuda_stream
    o_new = torch.full((B, d), float("nan"), dtype=torch.float16, device="cuda")
    _lib.check(lib.b200_cross_attention_qproj(x.data_ptr(), gamma.data_ptr(), lin.c1s.data_ptr(), lin.c2.data_ptr(), 1e-5,
                                              lin.weight.data_ptr(), lin.scales.data_ptr(), lin.bias.data_ptr(),
                                              cache.data_ptr(), qo.data_ptr(), o_new.data_ptr(), B, H, 64, S, st))
    torch.cuda.synchronize()
    assert torch.isfinite(o_new.float()).all()
    scale = max(1.0, o_ref.float().abs().max().item())
    err = (o_new.float() - o_ref.float()).abs().max().item()
    assert err <= 4e-3 * scale, f"fused vs two operators: {err} (scale {scale})"

    # fp32 torch: LayerNorm -> Linear on the dequantized int8 weight -> attention over the int8-round-tripped K / V
    w_q = lin_dequant(lib, lin, d)
    h = torch.nn.functional.layer_norm(x.float(), (d,), gamma.float(), beta.float(), 1e-5)
    q32 = h @ w_q + bias.cuda().half().float()
    kq = torch.clamp(torch.round(k.float() / t), -128, 127) * t
    vq = torch.clamp(torch.round(v.float() / t), -128, 127) * t
    qh = q32.view(B, H, 1, 64)
    kh = kq.view(B, S, H, 64).permute(0, 2, 1, 3)
    vh = vq.view(B, S, H, 64).permute(0, 2, 1, 3)
    att = torch.softmax(qh @ kh.transpose(-1, -2) / 8.0, dim=-1) @ vh
    ref32 = att.reshape(B, d)
    err32 = (o_new.float() - ref32).abs().max().item()
    assert err32 <= 6e-3 * max(1.0, ref32.abs().max().item()), f"fused vs fp32 torch: {err32}"


def lin_dequant(lib, lin, d):
    """[K, N] fp32 dequantized weight of a _QLinear through the matmul itself (identity activations, as the reference's
    un-convert test does, tests/quantization/test_weight_only_quant_matmul.py:121-130)."""
    from b200_whisper import _lib
    eye = torch.eye(d, dtype=torch.float16, device="cuda")
    out = torch.empty((d, lin.n), dtype=torch.float16, device="cuda")
    ws = torch.empty((max(lib.b200_woq_workspace_bytes(d, lin.n, d), 1 << 20),), dtype=torch.uint8, device="cuda")
    _lib.check(lib.b200_woq_int8_gemm(eye.data_ptr(), d, d, lin.weight.data_ptr(), lin.scales.data_ptr(), lin.n,
                                      out.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    return out.float()


def test_a_row_does_not_depend_on_its_batch_mates():
    """Utterances are the sharding unit: row b's output bits must not depend on which CTA served it
    or on the other rows of the batch."""
    from b200_whisper import _lib
    H, S = 20, 1500
    lib, lin, w, bias, gamma, beta, x, k, v, cache, oq, qo, t = _setup(H, 16, S, seed=77)
    d = H * 64
    st = torch.cuda.current_stream().cuda_stream

    def run(xs, cs):
        B = xs.shape[0]
        o = torch.empty((B, d), dtype=torch.float16, device="cuda")
        _lib.check(lib.b200_cross_attention_qproj(xs.data_ptr(), gamma.data_ptr(), lin.c1s.data_ptr(), lin.c2.data_ptr(), 1e-5,
                                                  lin.weight.data_ptr(), lin.scales.data_ptr(), lin.bias.data_ptr(),
                                                  cs.data_ptr(), qo.data_ptr(), o.data_ptr(), B, H, 64, S, st))
        torch.cuda.synchronize()
        return o

    full = run(x, cache)
    again = run(x, cache)
    assert torch.equal(full, again)
    part = run(x[4:13].contiguous(), cache[4:13].contiguous())   # 9 rows: another row -> CTA assignment
    assert torch.equal(part, full[4:13])


def test_unsupported_shapes_are_refused():
    import b200_whisper
    lib = b200_whisper.load()
    assert lib.b200_cross_attention_qproj_supported(16, 20, 64, 1500) == 1
    assert lib.b200_cross_attention_qproj_supported(1, 20, 64, 1500) == 0     # fewer pairs than SMs: the plain kernels
    assert lib.b200_cross_attention_qproj_supported(4, 20, 64, 1500) == 0
    assert lib.b200_cross_attention_qproj_supported(64, 20, 64, 1500) == 0    # 10 rows per CTA: more than one finishing pass
    assert lib.b200_cross_attention_qproj_supported(16, 20, 128, 1500) == 0   # head size
    assert lib.b200_cross_attention_qproj_supported(16, 32, 64, 1500) == 0    # more than two k-blocks per warp


@pytest.mark.parametrize("dims_name,B", [("large-v2 width, 2 layers", 16)])
def test_decoder_with_the_fused_kernel_matches_the_operator_chain(dims_name, B):
    """The whole generation step with fuse_cross_q on (one launch fewer per layer) against the two-operator chain: same
    tokens wherever the decision is clear, logits within fp16 noise."""
    from b200_whisper.runtime import WhisperDecoding
    from oracle import whisper_oracle as wo
    dims = wo.ModelDimensions(80, 1500, 1280, 20, 2, 51865, 448, 1280, 20, 2)
    sd = wo.synthetic_state_dict(dims, seed=9, decoder_only=True)
    L = dims.n_text_layer
    torch.manual_seed(3)
    xa = torch.randn(B, 200, dims.n_text_state).half().cuda()
    outs = []
    for fused in (True, False):
        dec = WhisperDecoding(dims, sd, B, [0.04] * L, [0.03] * L, n_audio_ctx=xa.shape[1])
        dec.fuse_cross_q = fused
        dec.set_encoder_output(xa)
        dec.reset()
        toks = [dec.prefill([[3, 7, 11]] * B).clone()]
        logits = [dec.logits.clone()]
        n0 = dec.lib.b200_launch_count()
        for _ in range(5):
            dec._step_body()
            toks.append(dec.next_tokens.clone())
            logits.append(dec.logits.clone())
        torch.cuda.synchronize()
        outs.append((torch.stack(toks, 1), torch.stack(logits, 1), dec.lib.b200_launch_count() - n0))
    (t_new, l_new, n_new), (t_old, l_old, n_old) = outs
    assert n_old - n_new == 5 * L, "the fused path must save one launch per layer and step"
    scale = l_old.abs().max().item()
    same = torch.ones(t_new.shape, dtype=torch.bool, device=t_new.device)
    for b in range(B):
        dd = (t_new[b] != t_old[b]).nonzero()
        if len(dd):
            same[b, int(dd[0]) + 1:] = False
            top2 = l_old[b, int(dd[0])].topk(2).values
            assert (top2[0] - top2[1]).item() <= 1e-2 * scale, "tokens diverged on a clear decision"
    assert same[:, :2].all()
    assert ((l_new - l_old).abs().amax(-1) * same).max().item() <= 4e-3 * scale
