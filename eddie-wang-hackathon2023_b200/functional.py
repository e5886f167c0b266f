"""gpt_attention / cross_attention / conv1d / glue ops -- eager mirror of the reference's graph-building functions
(T/tensorrt_llm/functional.py:2202-2244 conv1d, :2738-2971 gpt_attention) on torch CUDA tensors, executing the
plugin contract of GPTAttentionPlugin::enqueue (T/cpp/tensorrt_llm/plugins/gptAttentionPlugin/gptAttentionPlugin.cpp:203-379).
"""
import ctypes

import torch

from . import _lib
from .quantization.functional import _workspace


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("CUDA tensors required (there is no CPU fallback)")


def gpt_attention(tensor, past_key_value, sequence_length, past_key_value_length, masked_tokens, input_lengths,
                  max_input_length, cache_indirection, num_heads, head_size, q_scaling, rotary_embedding_dim,
                  neox_rotary_style, multi_block_mode, multi_query_mode, kv_orig_quant_scale, kv_quant_orig_scale,
                  use_int8_kv_cache, use_fp8_kv_cache=False, mask_type=1, kv_cache_block_pointers=None,
                  host_input_lengths=None, host_request_types=None, qkv_bias=None):
    """Same arguments as the reference.  Returns (output, present_key_value); present aliases past (the cache is
    updated in place, T/tests/attention/test_gpt_attention.py:245-248).

    tensor                 [B, S, 3*H*Dh] fp16 (S == 1 in the generation phase)
    past_key_value         [B, 2, H, Smax, Dh] int8 (use_int8_kv_cache) or fp16; with kv_cache_block_pointers: the
                           block pool [blocks, 2, H, tokens_per_block, Dh]
    sequence_length        [B] int32 CUDA: tokens already cached per sequence in the generation phase
    past_key_value_length  [2] int32 *host* tensor: [past_len, is_context]  (gptAttentionPlugin.cpp:261-278)
    masked_tokens          [B, Smax] int32 CUDA, 1 = padding between the prompt and the generated tokens
    input_lengths          [B] int32 CUDA
    max_input_length       [max_in] (only its shape is used, gptAttentionPlugin.cpp:284)
    cache_indirection      [B, beam=1, Smax] (Smax is read from its last dim, gptAttentionPlugin.cpp:335)
    """
    if rotary_embedding_dim != 0 or neox_rotary_style:
        raise NotImplementedError("rotary embeddings are not on the Whisper hot path")
    if multi_query_mode or use_fp8_kv_cache or host_request_types is not None:
        raise NotImplementedError("multi-query / fp8 KV / in-flight batching are not on the Whisper hot path")
    if cache_indirection is not None and cache_indirection.dim() == 3 and cache_indirection.shape[1] != 1:
        raise NotImplementedError("beam search is not on the Whisper hot path (beam width 1)")
    if past_key_value_length.is_cuda:
        raise ValueError("past_key_value_length must be a host tensor [past_len, is_context] "
                         "(the plugin reads it on the host, gptAttentionPlugin.cpp:261-278)")
    if tensor.dtype != torch.float16:
        raise TypeError("gpt_attention: float16 activations only")
    _need_cuda(tensor, past_key_value)
    lib = _lib.load()
    B, S, three_hidden = tensor.shape
    hidden = num_heads * head_size
    assert three_hidden == 3 * hidden, "qkv last dim must be 3 * num_heads * head_size"
    past_len, is_context = int(past_key_value_length[0]), bool(int(past_key_value_length[1]))
    paged = kv_cache_block_pointers is not None
    if paged:
        # paged KV cache (gptAttentionPlugin.cpp:314-326): past_key_value is the block POOL
        # [blocks, 2, H, tokens_per_block, Dh]; kv_cache_block_pointers [B, beam = 1, 2, max_blocks_per_seq] int64 (or the
        # reference's int32 view with twice the last dim) holds device addresses of [H, tokens_per_block, Dh] blocks
        if cache_indirection is None:
            raise ValueError("paged KV cache: max_seq_len is read from cache_indirection's last dim")
        bp = kv_cache_block_pointers
        if bp.dtype == torch.int32:
            bp = bp.contiguous().view(torch.int64)
        assert bp.is_cuda and bp.dtype == torch.int64
        bp = bp.reshape(B, -1, 2, bp.shape[-1])
        if bp.shape[1] != 1:
            raise NotImplementedError("beam search is not on the Whisper hot path (beam width 1)")
        bp = bp[:, 0].contiguous()
        max_blocks, tokens_per_block = bp.shape[-1], past_key_value.shape[3]
        max_seq_len = cache_indirection.shape[-1]
        assert tuple(past_key_value.shape[1:]) == (2, num_heads, tokens_per_block, head_size)
    else:
        max_seq_len = cache_indirection.shape[-1] if cache_indirection is not None else past_key_value.shape[3]
        assert tuple(past_key_value.shape) == (B, 2, num_heads, max_seq_len, head_size)
    assert past_key_value.dtype == (torch.int8 if use_int8_kv_cache else torch.float16)
    x = tensor.contiguous()
    out = torch.empty((B, S, hidden), dtype=torch.float16, device=tensor.device)
    st = _lib.stream_ptr()
    if is_context and paged:
        rc = lib.b200_attention_context_paged(_lib.ptr(x), _lib.ptr(input_lengths), _lib.ptr(out), _lib.ptr(bp), max_blocks,
                                              tokens_per_block,
                                              _lib.ptr(kv_orig_quant_scale) if use_int8_kv_cache else None, B, S,
                                              num_heads, head_size, max_seq_len, int(use_int8_kv_cache),
                                              float(q_scaling), st)
        _lib.check(rc, "gpt_attention (context, paged)")
    elif is_context:
        rc = lib.b200_attention_context(_lib.ptr(x), _lib.ptr(input_lengths), _lib.ptr(out), _lib.ptr(past_key_value),
                                        _lib.ptr(kv_orig_quant_scale) if use_int8_kv_cache else None, B, S, num_heads,
                                        head_size, max_seq_len, int(use_int8_kv_cache), float(q_scaling), st)
        _lib.check(rc, "gpt_attention (context)")
    else:
        assert S == 1, "generation phase expects one token per sequence"
        p = _lib.MmhaParams()
        p.qkv = _lib.ptr(x)
        p.qkv_bias = _lib.ptr(qkv_bias)
        p.out = _lib.ptr(out)
        p.kv_cache = _lib.ptr(past_key_value)
        p.sequence_lengths = _lib.ptr(sequence_length)
        p.masked_tokens = _lib.ptr(masked_tokens)
        p.kv_scale_orig_quant = _lib.ptr(kv_orig_quant_scale) if use_int8_kv_cache else None
        p.kv_scale_quant_orig = _lib.ptr(kv_quant_orig_scale) if use_int8_kv_cache else None
        p.batch_size, p.num_heads, p.head_size = B, num_heads, head_size
        p.max_seq_len, p.past_kv_length = max_seq_len, past_len
        p.int8_kv_cache = int(use_int8_kv_cache)
        p.q_scaling = float(q_scaling)
        if paged:
            rc = lib.b200_mmha_generation_paged(ctypes.byref(p), _lib.ptr(bp), max_blocks, tokens_per_block, st)
        else:
            rc = lib.b200_mmha_generation(ctypes.byref(p), st)
        _lib.check(rc, "gpt_attention (generation)")
    return out, past_key_value


def cross_attention(q, cross_kv, kv_quant_orig_scale, num_heads, head_size, use_int8_kv_cache=True, out=None):
    """Cached cross-attention: q [R, H*Dh] fp16, cross_kv [B, 2, H, S_enc, Dh] int8|fp16 -> [R, H*Dh] fp16; R must be a
    multiple of B and query row r attends to sequence r // (R // B).
    Reference: Attention.forward(cross_attention=True) unfused path, T/tensorrt_llm/layers/attention.py:308-323,385-406."""
    _need_cuda(q, cross_kv)
    lib = _lib.load()
    R = q.shape[0]
    B = cross_kv.shape[0]
    S = cross_kv.shape[3]
    assert tuple(cross_kv.shape) == (B, 2, num_heads, S, head_size) and R % B == 0
    if out is None:
        out = torch.empty((R, num_heads * head_size), dtype=torch.float16, device=q.device)
    nbytes = lib.b200_cross_attention_workspace_bytes(R, num_heads, head_size, S)
    ws = _workspace(nbytes, q.device)
    rc = lib.b200_cross_attention(_lib.ptr(q.contiguous()), _lib.ptr(cross_kv), _lib.ptr(kv_quant_orig_scale),
                                  _lib.ptr(out), R, R // B, num_heads, head_size, S, int(use_int8_kv_cache),
                                  _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "cross_attention")
    return out


def cross_kv_pack(k, v, kv_orig_quant_scale, num_heads, head_size, use_int8_kv_cache=True):
    """K, V projections [B, S, H*Dh] fp16 -> cross-KV cache [B, 2, H, S, Dh] (CrossAttn_KV output layout,
    T/tensorrt_llm/models/whisper/model.py:509-521), int8-quantized with the self-attention cache rule."""
    _need_cuda(k, v)
    lib = _lib.load()
    B, S, _ = k.shape
    cache = torch.empty((B, 2, num_heads, S, head_size), dtype=torch.int8 if use_int8_kv_cache else torch.float16,
                        device=k.device)
    rc = lib.b200_cross_kv_pack(_lib.ptr(k.contiguous()), _lib.ptr(v.contiguous()), _lib.ptr(cache),
                                _lib.ptr(kv_orig_quant_scale), B, S, num_heads, head_size, int(use_int8_kv_cache),
                                _lib.stream_ptr())
    _lib.check(rc, "cross_kv_pack")
    return cache


def conv1d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, activation=None, impl="tc"):
    """input [B, Cin, T] fp16, weight [Cout, Cin, k] or the reference's [Cout, Cin, k, 1] -> [B, Cout, Tout] fp16.
    `activation` ('gelu') is a B200 extension fused into the epilogue (the reference applies gelu as a separate layer,
    T/tensorrt_llm/models/whisper/model.py:154-157).  impl: 'tc' = tcgen05 implicit GEMM (default), 'simt' = the
    CUDA-core direct convolution (no workspace)."""
    if dilation != 1 or groups != 1:
        raise NotImplementedError("dilation / groups are not used by the Whisper stem")
    _need_cuda(input, weight)
    lib = _lib.load()
    if weight.dim() == 4:
        weight = weight.squeeze(-1)
    B, cin, t_in = input.shape
    cout, cin_w, ksize = weight.shape
    assert cin_w == cin
    t_out = (t_in + 2 * padding - ksize) // stride + 1
    out = torch.empty((B, cout, t_out), dtype=torch.float16, device=input.device)
    act = {None: _lib.ACT_NONE, "gelu": _lib.ACT_GELU_ERF, "gelu_tanh": _lib.ACT_GELU_TANH}[activation]
    if impl == "simt":
        rc = lib.b200_conv1d_fp16(_lib.ptr(input.contiguous()), _lib.ptr(weight.contiguous()), _lib.ptr(bias),
                                  _lib.ptr(out), B, cin, cout, t_in, ksize, stride, padding, act, _lib.stream_ptr())
    else:
        ws = torch.empty((lib.b200_conv1d_workspace_bytes(B, cin, cout, t_in, ksize),), dtype=torch.uint8,
                         device=input.device)
        rc = lib.b200_conv1d_fp16_tc(_lib.ptr(input.contiguous()), _lib.ptr(weight.contiguous()), _lib.ptr(bias),
                                     _lib.ptr(out), B, cin, cout, t_in, ksize, stride, padding, act, _lib.ptr(ws),
                                     ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "conv1d")
    return out


def layer_norm(input, normalized_shape, weight, bias, eps=1e-5, out=None):
    _need_cuda(input)
    lib = _lib.load()
    cols = input.shape[-1]
    x = input.contiguous()
    rows = x.numel() // cols
    if out is None:
        out = torch.empty_like(x)
    rc = lib.b200_layernorm_fp16(_lib.ptr(x), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(out), rows, cols, float(eps),
                                 _lib.stream_ptr())
    _lib.check(rc, "layer_norm")
    return out


def embedding_with_position(tokens, positions, token_embedding, positional_embedding, out=None):
    """tokens, positions: int32 CUDA [R]; -> [R, d] fp16 = tok_emb[tokens] + pos_emb[positions]."""
    lib = _lib.load()
    rows = tokens.numel()
    vocab, cols = token_embedding.shape
    if out is None:
        out = torch.empty((rows, cols), dtype=torch.float16, device=token_embedding.device)
    rc = lib.b200_embed_tokens_fp16(_lib.ptr(tokens), _lib.ptr(positions), _lib.ptr(token_embedding),
                                    _lib.ptr(positional_embedding), _lib.ptr(out), rows, cols, vocab,
                                    positional_embedding.shape[0], _lib.stream_ptr())
    _lib.check(rc, "embedding")
    return out


def logits_argmax(x, token_embedding, want_logits=True, logits_out=None, tokens_out=None):
    """x [R, d] fp16, token_embedding [V, d] fp16 -> (logits fp32 [R, V] or None, argmax int32 [R])."""
    lib = _lib.load()
    rows, cols = x.shape
    vocab = token_embedding.shape[0]
    if logits_out is None:
        logits_out = torch.empty((rows, vocab), dtype=torch.float32, device=x.device)
    if tokens_out is None:
        tokens_out = torch.empty((rows,), dtype=torch.int32, device=x.device)
    rc = lib.b200_logits_argmax_fp16(_lib.ptr(x.contiguous()), _lib.ptr(token_embedding), _lib.ptr(logits_out),
                                     _lib.ptr(tokens_out), rows, cols, vocab, None, 0, _lib.stream_ptr())
    _lib.check(rc, "logits_argmax")
    return (logits_out if want_logits else None), tokens_out


def bidirectional_attention(qkv, num_heads, head_size=64, out=None):
    """qkv [B, S, 3*H*Dh] fp16 (q | k | v) -> softmax(q k^T / sqrt(Dh)) v as [B, S, H*Dh] (encoder self-attention)."""
    _need_cuda(qkv)
    lib = _lib.load()
    B, S, three_hidden = qkv.shape
    assert three_hidden == 3 * num_heads * head_size
    if out is None:
        out = torch.empty((B, S, num_heads * head_size), dtype=torch.float16, device=qkv.device)
    rc = lib.b200_attention_bidirectional_fp16(_lib.ptr(qkv.contiguous()), _lib.ptr(out), B, S, num_heads, head_size,
                                               _lib.stream_ptr())
    _lib.check(rc, "bidirectional_attention")
    return out


class WhisperLogitFilter:
    """Device-side state of the Whisper logit filters + greedy update for a batch of sequences (the reference keeps
    Python lists per sequence: T/examples/whisper/decoding.py:134-300).  suppress: iterable of token ids."""

    def __init__(self, batch, vocab, eot, no_timestamps, timestamp_begin, blank_token, suppress=(),
                 max_initial_timestamp_index=None, device="cuda"):
        self.batch, self.vocab = batch, vocab
        self.eot, self.no_timestamps, self.timestamp_begin, self.blank = eot, no_timestamps, timestamp_begin, blank_token
        self.max_initial = -1 if max_initial_timestamp_index is None else int(max_initial_timestamp_index)
        words = torch.zeros(((vocab + 31) // 32,), dtype=torch.int64)
        for t in suppress:
            words[t >> 5] |= 1 << (t & 31)
        self.bitmap = (words & 0xFFFFFFFF).to(torch.uint32).to(device) if len(tuple(suppress)) else None
        self.state = torch.zeros((batch, 4), dtype=torch.int32, device=device)
        self.sum_logprobs = torch.zeros((batch,), dtype=torch.float32, device=device)

    def reset(self):
        self.state.zero_()
        self.sum_logprobs.zero_()

    def __call__(self, logits, next_token):
        """logits [batch, vocab] fp32 CUDA (not modified); next_token int32 [batch] receives the chosen tokens."""
        lib = _lib.load()
        rc = lib.b200_whisper_filtered_argmax(
            _lib.ptr(logits), self.batch, self.vocab, _lib.ptr(self.bitmap), self.eot,
            -1 if self.no_timestamps is None else self.no_timestamps, self.timestamp_begin, self.blank, self.max_initial,
            _lib.ptr(self.state), _lib.ptr(next_token), _lib.ptr(self.sum_logprobs), _lib.stream_ptr())
        _lib.check(rc, "whisper_filtered_argmax")
        return next_token
