"""WhisperEncoder / WhisperDecoder / CrossAttn_KV under the reference's names, constructor arguments and parameter
names (T/tensorrt_llm/models/whisper/model.py:74-118 ResidualAttentionBlock, :124-172 WhisperEncoder, :201-300
WhisperDecoder, :469-555 CrossAttn_KV / KVLinearBlock), as eager torch modules over b200_whisper.layers and
b200_whisper.functional.  The reference classes BUILD a TensorRT graph; these EXECUTE the same graph, one library
kernel per operator, with the GPTAttention-plugin semantics (in-place int8 KV cache) and the int8 cross-KV cache.

They keep a reference user's call sites (module tree, `load_from_state_dict` filling the same attributes that
examples/whisper/weight.py:40-372 fills, forward arguments).  The fast path for serving is runtime.WhisperPipeline
(fused epilogues, folded LayerNorm, CUDA-graph loop); tests check both give the same tokens.

Configuration on the hot path only: fp16 activations, int8 weight-only Linear layers (quant_mode.is_weight_only()),
optionally int8 KV caches; batch sizes are free (the reference hard-codes batch 1, model.py:329-338)."""
import torch

from ... import functional
from ...layers import Attention, AttentionMaskType, Conv1d, LayerNorm, RaggedTensor
from ...quantization.layer import WeightOnlyQuantLinear, WeightOnlyQuantRowLinear
from ...quantization.mode import QuantMode


def _w8(quant_mode):
    if not quant_mode.is_int8_weight_only():
        raise ValueError("this library implements the int8 weight-only configuration (--use_weight_only "
                         "--weight_only_precision int8); there is no fp16 Linear kernel on the hot path")


class MLP(torch.nn.Module):
    """n_state -> 4 n_state -> GELU -> n_state (T/tensorrt_llm/layers/mlp.py as used at model.py:99-103)."""

    def __init__(self, hidden_size, ffn_hidden_size, hidden_act='gelu', bias=True, dtype=torch.float16,
                 quant_mode=QuantMode.use_weight_only()):
        super().__init__()
        if hidden_act != 'gelu':
            raise ValueError("Whisper uses GELU")
        self.fc = WeightOnlyQuantLinear(hidden_size, ffn_hidden_size, bias=bias, dtype=dtype, quant_mode=quant_mode)
        self.proj = WeightOnlyQuantRowLinear(ffn_hidden_size, hidden_size, bias=bias, dtype=dtype, quant_mode=quant_mode)

    def forward(self, x):
        from ...quantization.functional import weight_only_quant_matmul
        # erf GELU fused into the first matmul's epilogue (the oracle's nn.GELU, torch_model.py:120)
        h = weight_only_quant_matmul(x, self.fc.weight, self.fc.per_channel_scale, 1, bias=self.fc.bias, activation='gelu')
        return self.proj(h)


class ResidualAttentionBlock(torch.nn.Module):

    def __init__(self, n_state, n_head, n_ctx, dtype=torch.float16, cross_attention=False, quant_mode=QuantMode.use_weight_only(),
                 mask_type=AttentionMaskType.padding):
        super().__init__()
        _w8(quant_mode)
        i8 = quant_mode.has_int8_kv_cache()
        self.attn = Attention(n_state, n_head, n_ctx, attention_mask_type=mask_type, dtype=dtype, use_int8_kv_cache=i8,
                              quant_mode=quant_mode)
        self.attn_ln = LayerNorm(n_state, dtype=dtype)
        self.cross_attn = Attention(n_state, n_head, n_ctx, cross_attention=True, dtype=dtype, use_int8_kv_cache=i8,
                                    quant_mode=quant_mode) if cross_attention else None
        self.cross_attn_ln = LayerNorm(n_state, dtype=dtype) if cross_attention else None
        self.mlp = MLP(n_state, n_state * 4, dtype=dtype, quant_mode=quant_mode)
        self.mlp_ln = LayerNorm(n_state, dtype=dtype)

    def forward(self, hidden_states, mask=None, sequence_length=None, past_key_value_length=None, masked_tokens=None,
                cache_indirection=None, multi_kv_cache=None, cross_kv_cache=None, use_cache=False):
        x = hidden_states.data
        ragged = lambda t: RaggedTensor.from_row_lengths(t, hidden_states.row_lengths, hidden_states.max_row_length)  # noqa: E731
        a = self.attn(ragged(self.attn_ln(x)), attention_mask=mask, past_key_value=multi_kv_cache,
                      sequence_length=sequence_length, past_key_value_length=past_key_value_length,
                      masked_tokens=masked_tokens, use_cache=use_cache, cache_indirection=cache_indirection)
        present = None
        if use_cache:
            a, present = a
        x = x + a.data
        if self.cross_attn is not None:
            x = x + self.cross_attn(ragged(self.cross_attn_ln(x)), cross_key_value=cross_kv_cache).data
        x = x + self.mlp(self.mlp_ln(x))
        return (ragged(x), present) if use_cache else ragged(x)


def _load_linear(lin, sd, prefix, bias=True, zero_bias=False):
    b = sd.get(prefix + ".bias") if bias else None
    lin.load_from_linear_weight(sd[prefix + ".weight"], b)
    if zero_bias and lin.bias is not None:
        lin.bias.zero_()


def _load_ln(ln, sd, prefix):
    ln.weight.copy_(sd[prefix + ".weight"].to(ln.weight.dtype))
    ln.bias.copy_(sd[prefix + ".bias"].to(ln.bias.dtype))


@torch.no_grad()
def _load_block(blk, sd, p):
    w = torch.cat([sd[p + ".attn.query.weight"], sd[p + ".attn.key.weight"], sd[p + ".attn.value.weight"]], dim=0)
    qb = sd[p + ".attn.query.bias"]
    b = torch.cat([qb, torch.zeros_like(qb), sd[p + ".attn.value.bias"]], dim=0)  # key has no bias (weight.py:221-226)
    blk.attn.qkv.load_from_linear_weight(w, b)
    _load_linear(blk.attn.dense, sd, p + ".attn.out")
    _load_ln(blk.attn_ln, sd, p + ".attn_ln")
    if blk.cross_attn is not None:
        _load_linear(blk.cross_attn.q_linear, sd, p + ".cross_attn.query")
        _load_linear(blk.cross_attn.dense, sd, p + ".cross_attn.out")
        _load_ln(blk.cross_attn_ln, sd, p + ".cross_attn_ln")
    _load_linear(blk.mlp.fc, sd, p + ".mlp.0")
    _load_linear(blk.mlp.proj, sd, p + ".mlp.2")
    _load_ln(blk.mlp_ln, sd, p + ".mlp_ln")


class WhisperEncoder(torch.nn.Module):

    def __init__(self, n_mels, n_ctx, n_state, n_head, n_layer, dtype=torch.float16, mask_type=AttentionMaskType.padding,
                 quant_mode=QuantMode.use_weight_only()):
        super().__init__()
        _w8(quant_mode)
        self.n_ctx, self.n_state, self.n_head = n_ctx, n_state, n_head
        self.conv1 = Conv1d(n_mels, n_state, kernel_size=3, padding=1, dtype=dtype)
        self.conv2 = Conv1d(n_state, n_state, kernel_size=3, stride=2, padding=1, dtype=dtype)
        self.register_buffer("positional_embedding", torch.zeros((n_ctx, n_state), dtype=dtype))
        self.blocks = torch.nn.ModuleList([
            ResidualAttentionBlock(n_state, n_head, n_ctx, dtype, mask_type=mask_type, quant_mode=QuantMode.use_weight_only())
            for _ in range(n_layer)])  # no KV cache in the encoder
        self.ln_post = LayerNorm(n_state, dtype=dtype)
        self.dtype = dtype

    @torch.no_grad()
    def load_from_state_dict(self, sd):
        """OpenAI-style `model_state_dict` (encoder.* keys), as examples/whisper/weight.py:40-110 assigns them."""
        dev = self.positional_embedding.device
        self.conv1.weight.copy_(sd["encoder.conv1.weight"].to(dev, self.dtype).unsqueeze(-1))
        self.conv1.bias.copy_(sd["encoder.conv1.bias"].to(dev, self.dtype))
        self.conv2.weight.copy_(sd["encoder.conv2.weight"].to(dev, self.dtype).unsqueeze(-1))
        self.conv2.bias.copy_(sd["encoder.conv2.bias"].to(dev, self.dtype))
        self.positional_embedding.copy_(sd["encoder.positional_embedding"].to(dev, self.dtype))
        for i, blk in enumerate(self.blocks):
            _load_block(blk, sd, f"encoder.blocks.{i}")
        _load_ln(self.ln_post, sd, "encoder.ln_post")
        return self

    def forward(self, x):
        """x: RaggedTensor (or tensor) of log-mel [B, n_mels, 2 * n_ctx] -> [B, n_ctx, n_state] (model.py:152-171)."""
        mel = x.data if isinstance(x, RaggedTensor) else x
        B = mel.shape[0]
        h = self.conv1(mel.to(self.dtype), activation="gelu")
        h = self.conv2(h, activation="gelu")
        h = (h.permute(0, 2, 1) + self.positional_embedding).contiguous()
        for blk in self.blocks:
            qkv = blk.attn.qkv(blk.attn_ln(h))  # [B, n_ctx, 3 * n_state]
            ctx = functional.bidirectional_attention(qkv, self.n_head, self.n_state // self.n_head)
            h = h + blk.attn.dense(ctx.view(B, self.n_ctx, self.n_state))
            h = h + blk.mlp(blk.mlp_ln(h))
        return self.ln_post(h)


class WhisperDecoder(torch.nn.Module):

    def __init__(self, n_vocab, n_ctx, n_state, n_head, n_layer, dtype=torch.float16, quant_mode=QuantMode(0),
                 mask_type=AttentionMaskType.causal):
        super().__init__()
        _w8(quant_mode)
        self.n_vocab, self.n_ctx, self.n_state, self.n_head, self.n_layer = n_vocab, n_ctx, n_state, n_head, n_layer
        self.dtype, self.quant_mode = dtype, quant_mode
        self.register_buffer("token_embedding_weight", torch.zeros((n_vocab, n_state), dtype=dtype))
        self.register_buffer("positional_embedding", torch.zeros((n_ctx, n_state), dtype=dtype))
        self.blocks = torch.nn.ModuleList([
            ResidualAttentionBlock(n_state, n_head, n_ctx, dtype=dtype, cross_attention=True, quant_mode=quant_mode,
                                   mask_type=mask_type) for _ in range(n_layer)])
        self.ln = LayerNorm(n_state, dtype=dtype)
        self.kv_dtype = torch.int8 if quant_mode.has_int8_kv_cache() else dtype

    @torch.no_grad()
    def load_from_state_dict(self, sd, kv_scales=None, cross_kv_scales=None):
        """decoder.* keys of the checkpoint; kv_scales / cross_kv_scales: per-layer scale_y_quant_orig
        (examples/whisper/weight.py:236-243) when the KV caches are int8."""
        dev = self.token_embedding_weight.device
        self.token_embedding_weight.copy_(sd["decoder.token_embedding.weight"].to(dev, self.dtype))
        self.positional_embedding.copy_(sd["decoder.positional_embedding"].to(dev, self.dtype))
        for i, blk in enumerate(self.blocks):
            _load_block(blk, sd, f"decoder.blocks.{i}")
            if kv_scales is not None:
                blk.attn.kv_quant_orig_scale.fill_(float(kv_scales[i]))
                blk.attn.kv_orig_quant_scale.fill_(1.0 / float(kv_scales[i]))
            if cross_kv_scales is not None:
                blk.cross_attn.kv_quant_orig_scale.fill_(float(cross_kv_scales[i]))
                blk.cross_attn.kv_orig_quant_scale.fill_(1.0 / float(cross_kv_scales[i]))
        _load_ln(self.ln, sd, "decoder.ln")
        return self

    def forward(self, x, positional_embedding=None, mask=None, sequence_length=None, past_key_value_length=None,
                masked_tokens=None, cache_indirection=None, multi_kv_cache=None, cross_kv_cache=None, use_cache=False):
        """x: RaggedTensor of token ids [B, S]; positional_embedding [S, n_state] (the reference slices it on the host,
        examples/whisper/decoding.py:612-620) or None to slice by past_key_value_length[0]; multi_kv_cache / cross_kv_cache:
        per-layer self KV caches [B, 2, H, n_ctx, 64] (updated in place) and cross-KV caches from CrossAttn_KV.
        -> logits [B, S, n_vocab] fp32 (, presents)"""
        tokens = x.data
        B, S = tokens.shape
        if positional_embedding is None:
            off = int(past_key_value_length[0]) if past_key_value_length is not None and not int(past_key_value_length[1]) else 0
            positional_embedding = self.positional_embedding[off:off + S]
        h = (self.token_embedding_weight[tokens.long()] + positional_embedding).contiguous()
        hidden = RaggedTensor.from_row_lengths(h, x.row_lengths, x.max_row_length)
        presents = []
        for i, blk in enumerate(self.blocks):
            hidden = blk(hidden, mask=mask, sequence_length=sequence_length, past_key_value_length=past_key_value_length,
                         masked_tokens=masked_tokens, cache_indirection=cache_indirection, multi_kv_cache=multi_kv_cache[i],
                         cross_kv_cache=cross_kv_cache[i], use_cache=use_cache)
            if use_cache:
                presents.append(hidden[1])
                hidden = hidden[0]
        h = self.ln(hidden.data)
        logits, _ = functional.logits_argmax(h.view(B * S, self.n_state), self.token_embedding_weight, want_logits=True)
        logits = logits.view(B, S, self.n_vocab)
        return (logits, presents) if use_cache else logits


class KVLinearBlock(torch.nn.Module):
    """key (no bias) and value projections of one decoder layer's cross-attention (model.py:469-487)."""

    def __init__(self, n_state, dtype=torch.float16, quant_mode=QuantMode.use_weight_only()):
        super().__init__()
        self.key = WeightOnlyQuantLinear(n_state, n_state, bias=False, dtype=dtype, quant_mode=quant_mode)
        self.value = WeightOnlyQuantLinear(n_state, n_state, bias=True, dtype=dtype, quant_mode=quant_mode)

    def forward(self, xa):
        return self.key(xa), self.value(xa)


class CrossAttn_KV(torch.nn.Module):
    """The `cross_kv_cache_warping` model (model.py:489-555): per decoder layer K = key(xa), V = value(xa), laid out
    [B, 2, H, S_enc, 64] for the attention kernel.  With an int8 KV cache the projections are quantized on the way out
    (scale_y_quant_orig per layer), so the decoder streams half the bytes every step."""

    def __init__(self, n_state, n_head, n_layer, dtype=torch.float16, quant_mode=QuantMode.use_weight_only()):
        super().__init__()
        _w8(quant_mode)
        self.n_state, self.n_head, self.n_layer, self.dtype = n_state, n_head, n_layer, dtype
        self.int8 = quant_mode.has_int8_kv_cache()
        self.blocks = torch.nn.ModuleList([KVLinearBlock(n_state, dtype, quant_mode) for _ in range(n_layer)])
        self.register_buffer("kv_orig_quant_scale", torch.ones((n_layer, 1), dtype=torch.float32))

    @torch.no_grad()
    def load_from_state_dict(self, sd, cross_kv_scales=None):
        for i, blk in enumerate(self.blocks):
            p = f"decoder.blocks.{i}.cross_attn"
            blk.key.load_from_linear_weight(sd[p + ".key.weight"], None)
            # torch_model.py:62-63: the value projection HAS a bias (the reference's weight.py:372 drops it by assigning
            # a non-existent attribute; that bug is not reproduced)
            blk.value.load_from_linear_weight(sd[p + ".value.weight"], sd[p + ".value.bias"])
            if cross_kv_scales is not None:
                self.kv_orig_quant_scale[i].fill_(1.0 / float(cross_kv_scales[i]))
        return self

    def forward(self, xa):
        """xa [B, S_enc, n_state] fp16 -> list of per-layer cross-KV caches [B, 2, H, S_enc, 64]."""
        out = []
        for i, blk in enumerate(self.blocks):
            k, v = blk(xa)
            out.append(functional.cross_kv_pack(k, v, self.kv_orig_quant_scale[i], self.n_head, self.n_state // self.n_head,
                                                use_int8_kv_cache=self.int8))
        return out
