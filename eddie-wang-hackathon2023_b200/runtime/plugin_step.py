"""PluginDecoderStep -- one generation step of the decoder driven through the reference's actual boundary: every Linear
layer through `WeightOnlyQuantMatmulPlugin::enqueue` and the masked self-attention through `GPTAttentionPlugin::enqueue`
(the C++ IPluginV2DynamicExt classes, reached the way TensorRT reaches them: registry -> creator -> createPlugin ->
enqueue with PluginTensorDesc arrays, include/b200_plugin_harness.h), one enqueue per operator like the reference's
engine (weightOnlyQuantMatmulPlugin.cpp:162-222, gptAttentionPlugin.cpp:230-340).

What TensorRT itself contributes to the reference's graph -- LayerNorm, the bias / GELU / residual elementwise layers,
the embedding gather, the logits MatMul -- runs here through the library's standalone kernels (no fused epilogues, no
folded LayerNorm: the plugin sees exactly the tensors the reference's plugin sees).  The cached cross-attention is the
library's kernel through the C ABI (the reference builds it from unfused TensorRT layers, attention.py:385-406).

Purpose: (1) `bench.py`'s `e2e_plugin` figure -- what the marshalling of the reference's boundary costs next to the
fused CUDA-graph path -- and (2) a parity test: same tokens as runtime.WhisperDecoding."""
import numpy as np
import torch

from .. import _lib
from ..plugin import TrtPlugin

HALF = 1  # nvinfer1::DataType::kHALF


def _woq_fields():
    # quantization/functional.py:61-70 of the reference
    return [("type_id", np.array([HALF], np.int32)), ("weight_type_id", np.array(1, dtype=np.int32))]


def _attn_fields(num_heads, head_size):
    # functional.py:2925-2930 of the reference, the Whisper decoder's configuration
    f = [("num_heads", num_heads, np.int32), ("head_size", head_size, np.int32), ("unidirectional", 1, np.int32),
         ("q_scaling", 1.0, np.float32), ("rotary_embedding_dim", 0, np.int32), ("neox_rotary_style", 0, np.int8),
         ("context_fmha_type", 0, np.int8), ("multi_block_mode", 0, np.int8), ("multi_query_mode", 0, np.int8),
         ("int8_kv_cache", 1, np.int32), ("fp8_kv_cache", 0, np.int32), ("remove_input_padding", 0, np.int8),
         ("mask_type", 1, np.int32), ("paged_kv_cache", 0, np.int32), ("type_id", HALF, np.int32),
         ("in_flight_batching", 0, np.int32)]
    return [(k, np.array([v], dtype=t)) for k, v, t in f]


class PluginDecoderStep:
    """Shares weights, KV caches, token / length buffers with a runtime.WhisperDecoding (`dec`): after `dec.prefill()`
    either object can take the next step and they must agree."""

    def __init__(self, dec):
        self.dec = dec
        self.lib = dec.lib
        B, d, H, Dh, Smax = dec.B, dec.d, dec.H, dec.Dh, dec.Smax
        dev = dec.device
        self.matmul = TrtPlugin.create("WeightOnlyQuantMatmul", _woq_fields())
        self.attention = TrtPlugin.create("GPTAttention", _attn_fields(H, Dh))
        self._f16 = lambda *shape: torch.empty(shape, dtype=torch.float16, device=dev)  # noqa: E731
        self.x, self.h = self._f16(B, d), self._f16(B, d)
        self.qkv, self.ctx, self.q = self._f16(B, 3 * d), self._f16(B, d), self._f16(B, d)
        self.mm = {n: self._f16(B, n) for n in {d, 3 * d, dec.d_ff}}
        self.ff = self._f16(B, dec.d_ff)
        self.masked = torch.zeros((B, Smax), dtype=torch.int32, device=dev)
        self.input_lengths = torch.ones((B,), dtype=torch.int32, device=dev)
        self.max_in = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.cache_ind = torch.zeros((B, 1, Smax), dtype=torch.int32, device=dev)
        self.host_pkl = np.zeros((2,), np.int32)  # HOST tensor [past_len, is_context] (gptAttentionPlugin.cpp:261-278)
        self.enqueues = 0
        # descriptor lists are static per shape: build them once (TensorRT also hands the plugin prebuilt descriptors)
        self._mm_desc = {}
        cache_shape = (B, 2, H, Smax, Dh)
        self._attn_ins = [((B, 1, 3 * d), "float16"), (cache_shape, "int8"), ((B,), "int32"), ((2,), "int32"),
                          ((B, Smax), "int32"), ((B,), "int32"), ((1,), "int32"), ((B, 1, Smax), "int32"),
                          ((1,), "float32"), ((1,), "float32")]
        self._attn_outs = [((B, 1, d), "float16"), (cache_shape, "int8")]
        ws_bytes = max(self.matmul.workspace_size([((1, B, dec.d_ff), "float16"), ((dec.d_ff, dec.d_ff // 4), "float32"),
                                                   ((dec.d_ff,), "float16")], [((1, B, dec.d_ff), "float16")]), 1 << 20)
        self.ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)

    def close(self):
        self.matmul.destroy()
        self.attention.destroy()

    def _st(self):
        return torch.cuda.current_stream(self.dec.device).cuda_stream

    def _linear(self, x, lin, out, act=None, residual=None):
        """matmul plugin, then the separate elementwise layers of the reference's graph (bias add, GELU, residual add),
        each rounding to fp16 like a TensorRT fp16 layer"""
        B = self.dec.B
        key = (lin.k, lin.n)
        if key not in self._mm_desc:
            self._mm_desc[key] = ([((1, B, lin.k), "float16"), ((lin.k, lin.n // 4), "float32"), ((lin.n,), "float16")],
                                  [((1, B, lin.n), "float16")])
        ins, outs = self._mm_desc[key]
        y = self.mm[lin.n]
        rc = self.matmul.enqueue(ins, outs, [x.data_ptr(), lin.weight.data_ptr(), lin.scales.data_ptr()], [y.data_ptr()],
                                 self.ws.data_ptr(), self._st())
        _lib.check(rc, "WeightOnlyQuantMatmul enqueue")
        self.enqueues += 1
        if lin.bias is not None:
            y = y + lin.bias
        if act == "gelu":
            y = torch.nn.functional.gelu(y.float()).half()
        if residual is not None:
            y = residual + y
        out.copy_(y)
        return out

    def _ln(self, x, wb, out):
        self.dec._ln(x, wb, out, self.dec.B)
        return out

    def step(self):
        """One greedy step for the whole batch: consumes dec.tokens / dec.seq_len, leaves dec.next_tokens, advances."""
        dec = self.dec
        past = int(dec._host_len)
        if past >= dec.Smax:
            raise RuntimeError("the text context is full")
        self._enqueue_step(past)
        dec._host_len = past + 1
        return dec.next_tokens

    def _enqueue_step(self, past):
        """The device work of one step: every launch goes to the current stream and reads its lengths / tokens from device
        buffers, so the sequence can be captured in a CUDA graph (TensorRT captures an engine's enqueue the same way)."""
        dec, lib, B = self.dec, self.lib, self.dec.B
        st = self._st()
        x = self.x
        _lib.check(lib.b200_embed_tokens_fp16(dec.tokens.data_ptr(), dec.seq_len.data_ptr(), dec.tok_emb.data_ptr(),
                                              dec.pos_emb.data_ptr(), x.data_ptr(), B, dec.d, dec.V, dec.Smax, st), "embed")
        self.host_pkl[0], self.host_pkl[1] = past, 0
        for i, lay in enumerate(dec.layers):
            self._ln(x, lay["attn_ln"], self.h)
            self._linear(self.h, lay["qkv"], self.qkv)
            cache = dec.self_kv[i]
            rc = self.attention.enqueue(
                self._attn_ins, self._attn_outs,
                [self.qkv.data_ptr(), cache.data_ptr(), dec.seq_len.data_ptr(), self.host_pkl.ctypes.data,
                 self.masked.data_ptr(), self.input_lengths.data_ptr(), self.max_in.data_ptr(), self.cache_ind.data_ptr(),
                 lay["kv_oq"].data_ptr(), lay["kv_qo"].data_ptr()],
                [self.ctx.data_ptr(), cache.data_ptr()], None, st)
            _lib.check(rc, "GPTAttention enqueue")
            self.enqueues += 1
            self._linear(self.ctx, lay["attn_out"], x, residual=x)
            self._ln(x, lay["cross_ln"], self.h)
            self._linear(self.h, lay["cross_q"], self.q)
            rc = lib.b200_cross_attention(self.q.data_ptr(), dec.cross_kv[i].data_ptr(), lay["ckv_qo"].data_ptr(),
                                          self.ctx.data_ptr(), B, 1, dec.H, dec.Dh, dec.S_enc, 1, dec.ws.data_ptr(),
                                          dec.ws.numel(), st)
            _lib.check(rc, "cross_attention")
            self._linear(self.ctx, lay["cross_out"], x, residual=x)
            self._ln(x, lay["mlp_ln"], self.h)
            self._linear(self.h, lay["fc1"], self.ff, act="gelu")
            self._linear(self.ff, lay["fc2"], x, residual=x)
        dec._head(x, B, dec.logits, dec.next_tokens)
        dec.seq_len.add_(1)
        dec.tokens.copy_(dec.next_tokens)

    def capture(self):
        """Captures the step -- host token ids in, all plugin enqueues and glue layers, next ids out -- in ONE CUDA graph,
        the way a TensorRT execution context is captured (enqueue on a capturing stream): the plugins' enqueue performs
        no allocation and no synchronization, and the generation kernel takes its lengths from the device tensor
        `sequence_length` (the HOST past-length scalar is only range-checked at capture time)."""
        dec = self.dec
        snap, host_len = dec._snapshot(), int(dec._host_len)
        enq0 = self.enqueues
        s = torch.cuda.Stream(device=dec.device)
        s.wait_stream(torch.cuda.current_stream(dec.device))
        with torch.cuda.stream(s):
            self._enqueue_step(host_len)  # warm-up outside capture (function attributes, allocator pool)
        torch.cuda.current_stream(dec.device).wait_stream(s)
        torch.cuda.synchronize(dec.device)
        dec._restore(snap)
        enq = self.enqueues
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            dec.tokens.copy_(dec._pinned_in, non_blocking=True)
            self._enqueue_step(host_len)
            dec._pinned_out.copy_(dec.next_tokens, non_blocking=True)
        torch.cuda.synchronize(dec.device)
        dec._restore(snap)
        self.enqueues_per_step = self.enqueues - enq
        self.enqueues = enq0  # neither the warm-up nor the capture advanced the sequences
        self.graph = g
        return g

    def step_host_graph(self, tokens_host):
        """host token ids in -> host next-token ids out through the captured step (capture() first)"""
        dec = self.dec
        if int(dec._host_len) >= dec.Smax:
            raise RuntimeError("the text context is full")
        dec._host_len = int(dec._host_len) + 1
        dec._pinned_in.copy_(torch.as_tensor(tokens_host, dtype=torch.int32))
        self.graph.replay()
        self.enqueues += self.enqueues_per_step
        torch.cuda.current_stream(dec.device).synchronize()
        return dec._pinned_out

    def step_host(self, tokens_host):
        """host token ids in -> host next-token ids out (pinned buffers of the decoder)"""
        dec = self.dec
        dec._pinned_in.copy_(torch.as_tensor(tokens_host, dtype=torch.int32))
        dec.tokens.copy_(dec._pinned_in, non_blocking=True)
        self.step()
        dec._pinned_out.copy_(dec.next_tokens, non_blocking=True)
        torch.cuda.current_stream(dec.device).synchronize()
        return dec._pinned_out
