"""Attention under the reference's name, constructor and forward arguments (T/tensorrt_llm/layers/attention.py:48-415,
including the cross-attention mode the hackathon entry added, :308-323,385-406), as an eager module:

  * self-attention: qkv = WeightOnlyQuantLinear(x); functional.gpt_attention (the GPTAttention plugin contract: int8 KV
    cache updated in place, context phase when past_key_value_length[1] == 1, generation otherwise); dense.
  * cross-attention: q = q_linear(x); functional.cross_attention over the int8 cross-KV cache that CrossAttn_KV built
    (the reference recomputes softmax(q K^T) V with fp16 K / V, attention.py:385-406); dense.

Only the configuration on the Whisper hot path is implemented (int8 weight-only Linear layers, learned absolute
positions, no tensor parallelism, no multi-query, beam width 1); anything else raises."""
import enum
import math

import torch

from .. import functional
from ..quantization.layer import WeightOnlyQuantLinear, WeightOnlyQuantRowLinear
from ..quantization.mode import QuantMode


class AttentionMaskType(enum.IntEnum):  # T/tensorrt_llm/layers/attention.py:30-33
    padding = 0
    causal = 1
    bidirectional = 2


class PositionEmbeddingType(enum.Enum):  # T/tensorrt_llm/layers/attention.py:36-45
    learned_absolute = enum.auto()
    rope = enum.auto()
    alibi = enum.auto()


class RaggedTensor:
    """(data, row_lengths, max_row_length) -- T/tensorrt_llm/functional.py RaggedTensor, as plain torch tensors."""

    def __init__(self, data, row_lengths=None, max_row_length=None):
        self.data, self.row_lengths, self.max_row_length = data, row_lengths, max_row_length

    @staticmethod
    def from_row_lengths(data, row_lengths, max_row_length=None):
        return RaggedTensor(data, row_lengths, max_row_length)


class Attention(torch.nn.Module):

    def __init__(self, hidden_size, num_attention_heads, max_position_embeddings, num_layers=1, cross_attention=False,
                 apply_query_key_layer_scaling=False, attention_mask_type=AttentionMaskType.padding, bias=True,
                 dtype=torch.float16, position_embedding_type=PositionEmbeddingType.learned_absolute,
                 neox_rotary_style=False, use_int8_kv_cache=False, rotary_embedding_percentage=1.0, tp_group=None,
                 tp_size=1, multi_block_mode=False, multi_query_mode=False,
                 quant_mode=QuantMode.use_weight_only()):
        super().__init__()
        if tp_size != 1 or multi_query_mode or position_embedding_type != PositionEmbeddingType.learned_absolute:
            raise ValueError("only the Whisper configuration is on the hot path (no TP, no multi-query, learned positions)")
        self.attention_mask_type = attention_mask_type
        self.attention_head_size = hidden_size // num_attention_heads
        self.num_attention_heads = num_attention_heads
        self.hidden_size = hidden_size
        self.max_position_embeddings = max_position_embeddings
        self.num_layers = num_layers
        self.norm_factor = math.sqrt(self.attention_head_size)
        self.q_scaling = 1
        if apply_query_key_layer_scaling:
            self.norm_factor *= num_layers
            self.q_scaling *= num_layers
        self.rotary_embedding_dim = 0
        self.neox_rotary_style = neox_rotary_style
        self.multi_block_mode, self.multi_query_mode = multi_block_mode, multi_query_mode
        self.use_int8_kv_cache = use_int8_kv_cache
        self.cross_attention = cross_attention
        # scale_y_quant_orig and its inverse (examples/whisper/weight.py:236-243)
        self.register_buffer("kv_orig_quant_scale", torch.ones((1,), dtype=torch.float32))
        self.register_buffer("kv_quant_orig_scale", torch.ones((1,), dtype=torch.float32))
        if cross_attention:
            self.q_linear = WeightOnlyQuantLinear(hidden_size, hidden_size, bias=bias, dtype=dtype, quant_mode=quant_mode)
            self.qkv = None
        else:
            self.q_linear = None
            self.qkv = WeightOnlyQuantLinear(hidden_size, hidden_size * 3, bias=bias, dtype=dtype, quant_mode=quant_mode)
        self.dense = WeightOnlyQuantRowLinear(hidden_size, hidden_size, bias=bias, dtype=dtype, quant_mode=quant_mode)

    def forward(self, hidden_states, xa=None, attention_mask=None, past_key_value=None, sequence_length=None,
                past_key_value_length=None, cross_key_value=None, masked_tokens=None, use_cache=False,
                cache_indirection=None, kv_cache_block_pointers=None, inflight_batching_args=None,
                past_key_value_pointers=None):
        if inflight_batching_args is not None or past_key_value_pointers is not None:
            raise NotImplementedError("in-flight batching is not on the Whisper hot path")
        assert isinstance(hidden_states, RaggedTensor)
        input_lengths, max_input_length = hidden_states.row_lengths, hidden_states.max_row_length
        x = hidden_states.data
        if self.cross_attention:
            if cross_key_value is None:
                raise ValueError("cross attention reads the cross-KV cache built by CrossAttn_KV (cross_key_value)")
            B, S, _ = x.shape
            q = self.q_linear(x)
            ctx = functional.cross_attention(q.reshape(B * S, self.hidden_size), cross_key_value, self.kv_quant_orig_scale,
                                             self.num_attention_heads, self.attention_head_size,
                                             use_int8_kv_cache=cross_key_value.dtype == torch.int8)
            ctx = ctx.view(B, S, self.hidden_size)
            present = None
        else:
            qkv = self.qkv(x)
            ctx, present = functional.gpt_attention(
                qkv, past_key_value, sequence_length, past_key_value_length, masked_tokens, input_lengths, max_input_length,
                cache_indirection, self.num_attention_heads, self.attention_head_size, self.q_scaling,
                self.rotary_embedding_dim, self.neox_rotary_style, self.multi_block_mode, self.multi_query_mode,
                self.kv_orig_quant_scale, self.kv_quant_orig_scale, self.use_int8_kv_cache,
                kv_cache_block_pointers=kv_cache_block_pointers)
        out = RaggedTensor.from_row_lengths(self.dense(ctx), input_lengths, max_input_length)
        if use_cache:
            return out, present
        return out
