"""WhisperDecoding -- the quantized Whisper decoder (int8 weight-only + int8 self/cross KV cache) on one B200.

Host-side counterpart of the reference's decoder runtime, T/examples/whisper/decoding.py (WhisperDecoding.decode
:543-659, xa2cross_key_value :515-541, main_loop :785-821, GreedyDecoder :274-300) and of the graph the reference
builds for it (T/tensorrt_llm/models/whisper/model.py:74-118 ResidualAttentionBlock, :201-300 WhisperDecoder,
:469-555 CrossAttn_KV).  Differences that are the point of this repo:

  * the GPTAttention plugin semantics are real here (preallocated int8 cache, in-place append, host/device scalars);
    the reference's example builds the unfused graph by accident;
  * cross-attention reads an int8 cross-KV cache through a split-KV streaming kernel;
  * one decoder step is captured once in a CUDA graph and replayed; the greedy token feedback stays on the device
    (the reference pays a Python engine launch and a stream synchronize per token, decoding.py:646-650).

Everything numerical happens in libb200_whisper.so through the C ABI; torch provides device memory, streams and the
CUDA-graph capture.  Batch elements are independent utterances (the multi-GPU sharding unit).
"""
import ctypes
import os

import torch

from .. import _lib, ops
from ..quantization.mode import QuantMode


class _QLinear:
    """Preprocessed int8 weight [K, N], fp16 per-channel scales [N], optional fp16 bias [N]."""

    def __init__(self, weight_out_in, bias, device):
        w_kn = weight_out_in.detach().to(device=device, dtype=torch.float16).t().contiguous()
        self.k, self.n = w_kn.shape
        self.weight, self.scales = ops.symmetric_quantize_last_axis_of_batched_matrix(w_kn, torch.int8)
        self.bias = None if bias is None else bias.detach().to(device=device, dtype=torch.float16).contiguous()
        self.c1s = self.c2 = None

    def fold_layernorm(self, lib, gamma, beta, stream):
        """Per-column vectors that let the LayerNorm in front of this Linear be folded into the GEMM kernel
        (b200_woq_ln_fold_prepare)."""
        self.c1s = torch.empty((self.n,), dtype=torch.float32, device=self.weight.device)
        self.c2 = torch.empty_like(self.c1s)
        _lib.check(lib.b200_woq_ln_fold_prepare(self.weight.data_ptr(), self.scales.data_ptr(), gamma.data_ptr(),
                                                beta.data_ptr(), self.k, self.n, self.c1s.data_ptr(), self.c2.data_ptr(),
                                                stream), "ln_fold_prepare")


def _cat_qkv(sd, p, device):
    w = torch.cat([sd[p + ".query.weight"], sd[p + ".key.weight"], sd[p + ".value.weight"]], dim=0)
    qb = sd[p + ".query.bias"]
    b = torch.cat([qb, torch.zeros_like(qb), sd[p + ".value.bias"]], dim=0)  # key has no bias (weight.py:221-226)
    return _QLinear(w, b, device)


class WhisperDecoding:

    def __init__(self, dims, state_dict, batch_size, kv_scales, cross_kv_scales, device="cuda",
                 quant_mode=QuantMode.use_weight_only().set_int8_kv_cache(), n_audio_ctx=None):
        """dims: object with n_vocab, n_text_ctx, n_text_state, n_text_head, n_text_layer, n_audio_ctx.
        state_dict: OpenAI-style `model_state_dict` (decoder.* keys; fp32 or fp16 CPU/GPU tensors).
        kv_scales / cross_kv_scales: per-layer scale_y_quant_orig (= max|y|/127, weight.py:236-243)."""
        if not (quant_mode.is_int8_weight_only() and quant_mode.has_int8_kv_cache()):
            raise ValueError("WhisperDecoding implements the int8 weight-only + int8 KV cache configuration")
        self.lib = _lib.load()
        _lib.check(self.lib.b200_init(), "b200_init")
        # inside a decoder step the caches are never written by the kernel right before the one that reads them
        self.static_kv = os.environ.get("B200_STATIC_KV", "1") != "0"
        self.dims = dims
        self.B = batch_size
        self.device = torch.device(device)
        self.d = dims.n_text_state
        self.H = dims.n_text_head
        self.Dh = self.d // self.H
        self.L = dims.n_text_layer
        self.V = dims.n_vocab
        self.Smax = dims.n_text_ctx
        self.S_enc = n_audio_ctx or dims.n_audio_ctx
        dev = self.device
        sd = state_dict
        f16 = lambda t: t.detach().to(device=dev, dtype=torch.float16).contiguous()  # noqa: E731

        self.tok_emb = f16(sd["decoder.token_embedding.weight"])
        self.pos_emb = f16(sd["decoder.positional_embedding"])
        self.ln_w, self.ln_b = f16(sd["decoder.ln.weight"]), f16(sd["decoder.ln.bias"])
        self.layers = []
        for i in range(self.L):
            p = f"decoder.blocks.{i}"
            lay = {
                "attn_ln": (f16(sd[p + ".attn_ln.weight"]), f16(sd[p + ".attn_ln.bias"])),
                "qkv": _cat_qkv(sd, p + ".attn", dev),
                "attn_out": _QLinear(sd[p + ".attn.out.weight"], sd[p + ".attn.out.bias"], dev),
                "cross_ln": (f16(sd[p + ".cross_attn_ln.weight"]), f16(sd[p + ".cross_attn_ln.bias"])),
                "cross_q": _QLinear(sd[p + ".cross_attn.query.weight"], sd[p + ".cross_attn.query.bias"], dev),
                "cross_k": _QLinear(sd[p + ".cross_attn.key.weight"], None, dev),
                "cross_v": _QLinear(sd[p + ".cross_attn.value.weight"], sd[p + ".cross_attn.value.bias"], dev),
                "cross_out": _QLinear(sd[p + ".cross_attn.out.weight"], sd[p + ".cross_attn.out.bias"], dev),
                "mlp_ln": (f16(sd[p + ".mlp_ln.weight"]), f16(sd[p + ".mlp_ln.bias"])),
                "fc1": _QLinear(sd[p + ".mlp.0.weight"], sd[p + ".mlp.0.bias"], dev),
                "fc2": _QLinear(sd[p + ".mlp.2.weight"], sd[p + ".mlp.2.bias"], dev),
                # kv_orig_quant_scale = 1/t, kv_quant_orig_scale = t (weight.py:242-243)
                "kv_oq": torch.tensor([1.0 / kv_scales[i]], dtype=torch.float32, device=dev),
                "kv_qo": torch.tensor([kv_scales[i]], dtype=torch.float32, device=dev),
                "ckv_oq": torch.tensor([1.0 / cross_kv_scales[i]], dtype=torch.float32, device=dev),
                "ckv_qo": torch.tensor([cross_kv_scales[i]], dtype=torch.float32, device=dev),
            }
            self.layers.append(lay)
        # LayerNorm -> Linear pairs run as ONE kernel (B200_FUSE_LN=0: separate LayerNorm launch, for A/B timing)
        self.fuse_ln = os.environ.get("B200_FUSE_LN", "1") != "0"
        # qkv projection + self-attention of the generation step as ONE kernel (csrc/qkv_mmha.cu).  Opt-in
        # (B200_FUSE_QKV_MMHA=1 or .fuse_qkv_mmha = True): parity-green, but measured slower than the two tuned kernels it
        # replaces (8.5 vs 7.6 us per layer at batch 16, step 1.38 vs 1.29 ms)
        self.fuse_qkv_mmha = os.environ.get("B200_FUSE_QKV_MMHA", "0") == "1"
        # generation phase: the cross-attention kernel computes its own q projection (csrc/attention.cu,
        # cross_attention_qproj_kernel): one launch fewer per layer.  Opt-in (B200_FUSE_XQ=1 or .fuse_cross_q = True):
        # parity-green, but measured slower than the two kernels it replaces (step 1.306 vs 1.292 ms at batch 16: the
        # cross-attention is issue-bound, not stream-bound, so starting its stream earlier buys nothing and the projection
        # sits on its critical path)
        self.fuse_cross_q = os.environ.get("B200_FUSE_XQ", "0") == "1"
        if self.fuse_ln:
            st0 = torch.cuda.current_stream(dev).cuda_stream
            for lay in self.layers:
                for ln, lin in (("attn_ln", "qkv"), ("cross_ln", "cross_q"), ("mlp_ln", "fc1")):
                    lay[lin].fold_layernorm(self.lib, lay[ln][0], lay[ln][1], st0)

        B, d = self.B, self.d
        self.self_kv = [torch.zeros((B, 2, self.H, self.Smax, self.Dh), dtype=torch.int8, device=dev)
                        for _ in range(self.L)]
        self.cross_kv = [None] * self.L
        self.seq_len = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.tokens = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.next_tokens = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.logits = torch.empty((B, self.V), dtype=torch.float32, device=dev)
        self._bufs = {}
        max_rows = B * 8
        ws_bytes = max(self.lib.b200_woq_workspace_bytes(max(max_rows, B * self.S_enc), 4 * d, 4 * d),
                       self.lib.b200_cross_attention_workspace_bytes(max_rows, self.H, self.Dh, self.S_enc), 1 << 20)
        self.ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        self.graph = None
        self.graph_host = None
        # Independent utterances can be stepped as several interleaved chains of kernels (each on its own stream, all
        # inside the same CUDA graph): while one chain sits in the dependency latency between two small GEMMs, or
        # streams its cross-KV cache, the other chains' kernels fill the SMs.  Weights are read once per chain (the
        # second read is normally an L2 hit).  B200_CHAINS=1 restores the single chain.
        self.n_chains = int(os.environ.get("B200_CHAINS", "1"))
        if self.n_chains < 1 or B % self.n_chains != 0:
            self.n_chains = 1
        self._chain_streams = [torch.cuda.Stream(device=dev) for _ in range(self.n_chains - 1)] \
            if torch.cuda.is_available() else []
        self._chain_ws = [torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) for _ in range(self.n_chains - 1)]
        # side stream that pulls the next layer's cross-KV cache into L2 while the current layer's small kernels run
        self.prefetch_cross_kv = os.environ.get("B200_XKV_PREFETCH", "0") != "0"
        self._side = torch.cuda.Stream(device=dev) if torch.cuda.is_available() else None
        self._pinned_in = torch.zeros((B,), dtype=torch.int32).pin_memory() if torch.cuda.is_available() else None
        self._pinned_out = torch.zeros((B,), dtype=torch.int32).pin_memory() if torch.cuda.is_available() else None
        # The generation step as ONE persistent kernel (b200_decoder_step): weights and cross-KV stream through a
        # shared-memory ring ahead of the dependency chain.  Opt-in (B200_STEP_KERNEL=1 or .step_kernel = True): measured
        # on B200 it is parity-green but slower than the kernel-per-operator chain (1.78 vs 1.28 ms per batch-16 step;
        # DESIGN.md section 7), so the chain stays the default.
        self.d_ff = self.layers[0]["fc1"].n
        self.step_kernel_available = (self.fuse_ln and B <= 16
                                      and self.d <= 1280 and self.n_chains == 1
                                      and self.d_ff % (64 * ((self.d_ff + 1279) // 1280)) == 0)
        self.step_kernel = self.step_kernel_available and os.environ.get("B200_STEP_KERNEL", "0") == "1"
        self.step_ctas = int(os.environ.get("B200_STEP_CTAS", "0"))
        self._step_table = None
        self._step_scratch = None

    # ---- thin wrappers over the C ABI (pointers only; no torch math) --------------------------------------
    def _st(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _gemm(self, x, rows, lin, out, bias=True, act=_lib.ACT_NONE, residual=None, ws=None):
        ws = self.ws if ws is None else ws
        rc = self.lib.b200_woq_int8_gemm_fused(
            x.data_ptr(), rows, lin.k, lin.weight.data_ptr(), lin.scales.data_ptr(), lin.n,
            lin.bias.data_ptr() if (bias and lin.bias is not None) else None, act,
            residual.data_ptr() if residual is not None else None, out.data_ptr(), ws.data_ptr(), ws.numel(),
            self._st())
        _lib.check(rc, "woq gemm")

    def _gemm_ln(self, x, wb, rows, lin, out, act=_lib.ACT_NONE, ws=None):
        """out = act(LayerNorm(x; wb) @ W + bias) with the LayerNorm folded into the GEMM kernel (decode-sized row
        counts; larger ones take the two-launch route inside the same entry point)."""
        ws = self.ws if ws is None else ws
        rc = self.lib.b200_woq_int8_gemm_ln_folded(
            x.data_ptr(), wb[0].data_ptr(), wb[1].data_ptr(), lin.c1s.data_ptr(), lin.c2.data_ptr(), 1e-5, rows, lin.k,
            lin.weight.data_ptr(), lin.scales.data_ptr(), lin.n, lin.bias.data_ptr() if lin.bias is not None else None,
            act, None, out.data_ptr(), ws.data_ptr(), ws.numel(), self._st())
        _lib.check(rc, "woq gemm (folded LayerNorm)")

    def _ln(self, x, wb, out, rows):
        _lib.check(self.lib.b200_layernorm_fp16(x.data_ptr(), wb[0].data_ptr(), wb[1].data_ptr(), out.data_ptr(), rows,
                                                self.d, 1e-5, self._st()), "layernorm")

    def _buf(self, name, rows, cols, chain=0):
        key = (name, rows, cols, chain)
        if key not in self._bufs:
            self._bufs[key] = torch.empty((rows, cols), dtype=torch.float16, device=self.device)
        return self._bufs[key]

    # ---- cross-KV (the cross_kv_cache_warping model: CrossAttn_KV, model.py:469-555) ---------------------------
    def set_encoder_output(self, xa):
        """xa [B, S_enc, d] fp16: computes the int8 cross-KV cache of every layer (K has no bias, V has one,
        torch_model.py:62-63)."""
        B, S, d = xa.shape
        assert B == self.B and S == self.S_enc and d == self.d
        x = xa.to(device=self.device, dtype=torch.float16).contiguous().view(B * S, d)
        k = torch.empty((B * S, d), dtype=torch.float16, device=self.device)
        v = torch.empty_like(k)
        for i, lay in enumerate(self.layers):
            self._gemm(x, B * S, lay["cross_k"], k, bias=False)
            self._gemm(x, B * S, lay["cross_v"], v)
            if self.cross_kv[i] is None:
                self.cross_kv[i] = torch.empty((B, 2, self.H, S, self.Dh), dtype=torch.int8, device=self.device)
                self._step_table = None  # the step kernel's layer table holds the cache pointers
            rc = self.lib.b200_cross_kv_pack(k.data_ptr(), v.data_ptr(), self.cross_kv[i].data_ptr(),
                                             lay["ckv_oq"].data_ptr(), B, S, self.H, self.Dh, 1, self._st())
            _lib.check(rc, "cross_kv_pack")

    def set_cross_kv(self, caches):
        """Installs precomputed int8 cross-KV caches [B, 2, H, S_enc, Dh] (synthetic benchmarks)."""
        assert len(caches) == self.L
        self.cross_kv = [c.to(self.device) for c in caches]
        # captured graphs and the step kernel's layer table hold the old device pointers
        self.graph = self.graph_host = None
        self._step_table = None

    # ---- the persistent step kernel (b200_decoder_step) -----------------------------------------------------------
    def _build_step_table(self):
        """Device array of b200_decoder_layer (32 pointers per layer) + the zero-initialised scratch."""
        names = _lib.DecoderLayer._names
        rows = []
        for i, lay in enumerate(self.layers):
            ent = {
                "attn_ln_gamma": lay["attn_ln"][0], "qkv_w": lay["qkv"].weight, "qkv_scales": lay["qkv"].scales,
                "qkv_bias": lay["qkv"].bias, "qkv_c1s": lay["qkv"].c1s, "qkv_c2": lay["qkv"].c2,
                "attn_out_w": lay["attn_out"].weight, "attn_out_scales": lay["attn_out"].scales,
                "attn_out_bias": lay["attn_out"].bias,
                "cross_ln_gamma": lay["cross_ln"][0], "cross_q_w": lay["cross_q"].weight,
                "cross_q_scales": lay["cross_q"].scales, "cross_q_bias": lay["cross_q"].bias,
                "cross_q_c1s": lay["cross_q"].c1s, "cross_q_c2": lay["cross_q"].c2,
                "cross_out_w": lay["cross_out"].weight, "cross_out_scales": lay["cross_out"].scales,
                "cross_out_bias": lay["cross_out"].bias,
                "mlp_ln_gamma": lay["mlp_ln"][0], "fc1_w": lay["fc1"].weight, "fc1_scales": lay["fc1"].scales,
                "fc1_bias": lay["fc1"].bias, "fc1_c1s": lay["fc1"].c1s, "fc1_c2": lay["fc1"].c2,
                "fc2_w": lay["fc2"].weight, "fc2_scales": lay["fc2"].scales, "fc2_bias": lay["fc2"].bias,
                "self_kv": self.self_kv[i], "kv_scale_orig_quant": lay["kv_oq"], "kv_scale_quant_orig": lay["kv_qo"],
                "cross_kv": self.cross_kv[i], "cross_kv_scale_quant_orig": lay["ckv_qo"],
            }
            rows.append([0 if ent[n] is None else ent[n].data_ptr() for n in names])
        self._step_table = torch.tensor(rows, dtype=torch.int64).to(self.device)
        if self._step_scratch is None:
            nbytes = self.lib.b200_decoder_step_scratch_bytes(self.H, self.d_ff)
            self._step_scratch = torch.zeros((nbytes,), dtype=torch.uint8, device=self.device)

    def _step_kernel_call(self, x_out):
        if self._step_table is None:
            self._build_step_table()
        p = _lib.DecoderStepParams()
        p.layers = self._step_table.data_ptr()
        p.n_layers, p.batch_size, p.num_heads, p.d_ff = self.L, self.B, self.H, self.d_ff
        p.max_seq_len, p.enc_len, p.vocab, p.n_ctx = self.Smax, self.S_enc, self.V, self.Smax
        p.tokens, p.sequence_lengths = self.tokens.data_ptr(), self.seq_len.data_ptr()
        p.tok_emb, p.pos_emb = self.tok_emb.data_ptr(), self.pos_emb.data_ptr()
        p.x_out, p.scratch = x_out.data_ptr(), self._step_scratch.data_ptr()
        p.ln_eps, p.max_ctas = 1e-5, self.step_ctas
        _lib.check(self.lib.b200_decoder_step(ctypes.byref(p), self._st()), "decoder_step")

    def step_kernel_status(self):
        """0 unless a barrier or ring wait of the persistent step kernel timed out (synchronises the device)."""
        if self._step_scratch is None:
            return 0
        st = ctypes.c_int32(0)
        _lib.check(self.lib.b200_decoder_step_status(self._step_scratch.data_ptr(), ctypes.byref(st)), "status")
        return st.value

    def reset(self):
        self.seq_len.zero_()
        self._host_len = 0  # host-side copy of the (uniform) sequence length: bounds checks without a device read
        if getattr(self, "logit_filter", None) is not None:
            self.logit_filter.reset()

    # ---- one pass over the decoder stack for `rows` query rows ----------------------------------------------
    def _stack(self, x, rows, s_q, context, input_lengths=None, b0=0, nb=None, chain=0):
        """x [rows, d] are the rows of the batch elements [b0, b0 + nb) (nb * s_q == rows)."""
        d, H, Dh = self.d, self.H, self.Dh
        nb = self.B if nb is None else nb
        ws = self.ws if chain == 0 else self._chain_ws[chain - 1]
        h = self._buf("h", rows, d, chain)
        qkv = self._buf("qkv", rows, 3 * d, chain)
        ctx = self._buf("ctx", rows, d, chain)
        q = self._buf("q", rows, d, chain)
        u = self._buf("u", rows, 4 * d, chain)
        st = self._st()
        main = torch.cuda.current_stream(self.device)
        prefetch = self.prefetch_cross_kv and not context and self._side is not None and self.n_chains == 1
        if prefetch:
            self._side.wait_stream(main)
            self._prefetch(0)
        # generation phase: LayerNorm + qkv projection + self-attention as ONE kernel per layer (csrc/qkv_mmha.cu)
        fused_qkv = (not context and self.fuse_ln and self.fuse_qkv_mmha and s_q == 1
                     and self.lib.b200_qkv_mmha_decode_supported(nb, H, Dh) == 1)
        # generation phase: cross_q projection inside the cross-attention kernel (needs the static-cache promise: the
        # kernel streams the cross-KV cache before its dependency wait)
        fused_xq = (not context and self.fuse_ln and self.fuse_cross_q and self.static_kv and s_q == 1
                    and self.lib.b200_cross_attention_qproj_supported(nb, H, Dh, self.S_enc) == 1)
        for i, lay in enumerate(self.layers):
            kv_i = self.self_kv[i] if nb == self.B else self.self_kv[i][b0:b0 + nb]
            ckv_i = self.cross_kv[i] if nb == self.B else self.cross_kv[i][b0:b0 + nb]
            if fused_qkv:
                lin = lay["qkv"]
                rc = self.lib.b200_qkv_mmha_decode(
                    x.data_ptr(), lay["attn_ln"][0].data_ptr(), lin.c1s.data_ptr(), lin.c2.data_ptr(), 1e-5,
                    lin.weight.data_ptr(), lin.scales.data_ptr(), lin.bias.data_ptr() if lin.bias is not None else None,
                    kv_i.data_ptr(), self.seq_len[b0:b0 + nb].data_ptr(), lay["kv_oq"].data_ptr(), lay["kv_qo"].data_ptr(),
                    ctx.data_ptr(), nb, H, Dh, self.Smax, st)
                _lib.check(rc, "qkv_mmha_decode")
            elif self.fuse_ln:
                self._gemm_ln(x, lay["attn_ln"], rows, lay["qkv"], qkv, ws=ws)
            else:
                self._ln(x, lay["attn_ln"], h, rows)
                self._gemm(h, rows, lay["qkv"], qkv, ws=ws)
            if fused_qkv:
                pass
            elif context:
                rc = self.lib.b200_attention_context(
                    qkv.data_ptr(), input_lengths.data_ptr() if input_lengths is not None else None, ctx.data_ptr(),
                    kv_i.data_ptr(), lay["kv_oq"].data_ptr(), nb, s_q, H, Dh, self.Smax, 1, 1.0, st)
                _lib.check(rc, "attention_context")
            else:
                p = _lib.MmhaParams()
                p.qkv, p.qkv_bias, p.out = qkv.data_ptr(), None, ctx.data_ptr()
                p.kv_cache = kv_i.data_ptr()
                p.sequence_lengths = self.seq_len[b0:b0 + nb].data_ptr()
                p.masked_tokens = None
                p.kv_scale_orig_quant = lay["kv_oq"].data_ptr()
                p.kv_scale_quant_orig = lay["kv_qo"].data_ptr()
                p.batch_size, p.num_heads, p.head_size = nb, H, Dh
                p.max_seq_len, p.past_kv_length, p.int8_kv_cache, p.q_scaling = self.Smax, 0, 1, 1.0
                _lib.check(self.lib.b200_mmha_generation(ctypes.byref(p), st), "mmha_generation")
            self._gemm(ctx, rows, lay["attn_out"], x, residual=x, ws=ws)
            if fused_xq:
                if not getattr(self, "_measure_without_cross_attention", False):
                    lin = lay["cross_q"]
                    rc = self.lib.b200_cross_attention_qproj(
                        x.data_ptr(), lay["cross_ln"][0].data_ptr(), lin.c1s.data_ptr(), lin.c2.data_ptr(), 1e-5,
                        lin.weight.data_ptr(), lin.scales.data_ptr(), lin.bias.data_ptr() if lin.bias is not None else None,
                        ckv_i.data_ptr(), lay["ckv_qo"].data_ptr(), ctx.data_ptr(), nb, H, Dh, self.S_enc, st)
                    _lib.check(rc, "cross_attention_qproj")
            elif self.fuse_ln:
                self._gemm_ln(x, lay["cross_ln"], rows, lay["cross_q"], q, ws=ws)
            else:
                self._ln(x, lay["cross_ln"], h, rows)
                self._gemm(h, rows, lay["cross_q"], q, ws=ws)
            if fused_xq:
                pass
            elif not getattr(self, "_measure_without_cross_attention", False):  # bench.py: in-graph cost by difference
                rc = self.lib.b200_cross_attention(q.data_ptr(), ckv_i.data_ptr(), lay["ckv_qo"].data_ptr(),
                                                   ctx.data_ptr(), rows, s_q, H, Dh, self.S_enc, 1, ws.data_ptr(),
                                                   ws.numel(), st)
                _lib.check(rc, "cross_attention")
            if prefetch and i + 1 < self.L:
                # layer i's cross-KV is dead now: start pulling layer i+1's while the MLP and the next self-attention run
                self._side.wait_stream(main)
                self._prefetch(i + 1)
            self._gemm(ctx, rows, lay["cross_out"], x, residual=x, ws=ws)
            if self.fuse_ln:
                self._gemm_ln(x, lay["mlp_ln"], rows, lay["fc1"], u, act=_lib.ACT_GELU_ERF, ws=ws)
            else:
                self._ln(x, lay["mlp_ln"], h, rows)
                self._gemm(h, rows, lay["fc1"], u, act=_lib.ACT_GELU_ERF, ws=ws)
            self._gemm(u, rows, lay["fc2"], x, residual=x, ws=ws)
        if prefetch:
            main.wait_stream(self._side)
        return x

    def _prefetch(self, i):
        c = self.cross_kv[i]
        with torch.cuda.stream(self._side):
            _lib.check(self.lib.b200_l2_prefetch(c.data_ptr(), c.numel() * c.element_size(), self._side.cuda_stream),
                       "l2_prefetch")

    def enable_logit_filters(self, eot, no_timestamps, timestamp_begin, blank_token, suppress=(),
                             max_initial_timestamp_index=None):
        """Greedy decoding with the reference's logit filters (SuppressBlank, SuppressTokens, ApplyTimestampRules,
        decoding.py:134-217,332-348) evaluated on the device inside the captured step.  Call before prefill()."""
        from ..functional import WhisperLogitFilter
        self.logit_filter = WhisperLogitFilter(self.B, self.V, eot, no_timestamps, timestamp_begin, blank_token, suppress,
                                               max_initial_timestamp_index, device=self.device)
        self.graph = self.graph_host = None

    def _head(self, x_rows, rows, logits, next_tokens):
        h = self._buf("hf", rows, self.d)
        self._ln(x_rows, (self.ln_w, self.ln_b), h, rows)
        filt = getattr(self, "logit_filter", None)
        rc = self.lib.b200_logits_argmax_fp16(h.data_ptr(), self.tok_emb.data_ptr(), logits.data_ptr(),
                                              None if filt is not None else next_tokens.data_ptr(), rows, self.d,
                                              self.V, None, 0, self._st())
        _lib.check(rc, "logits_argmax")
        if filt is not None:
            filt(logits, next_tokens)

    # ---- public API ----------------------------------------------------------------------------------------
    def prefill(self, prompt_tokens):
        """Context phase: prompt_tokens [B, S] int (all sequences the same length, like the reference's
        sot/language/task prompt, decoding.py:314-319).  Returns the first generated tokens [B] (device int32)."""
        B = self.B
        prompt = torch.as_tensor(prompt_tokens, dtype=torch.int32, device=self.device).view(B, -1).contiguous()
        S = prompt.shape[1]
        if S > self.Smax:
            raise ValueError(f"prompt of {S} tokens exceeds n_text_ctx = {self.Smax}")
        rows = B * S
        pos = torch.arange(S, dtype=torch.int32, device=self.device).repeat(B)
        x = self._buf("x", rows, self.d)
        _lib.check(self.lib.b200_embed_tokens_fp16(prompt.data_ptr(), pos.data_ptr(), self.tok_emb.data_ptr(),
                                                   self.pos_emb.data_ptr(), x.data_ptr(), rows, self.d, self.V,
                                                   self.Smax, self._st()), "embed")
        self._stack(x, rows, S, context=True)
        last = x.view(B, S, self.d)[:, S - 1, :].contiguous()
        self._head(last, B, self.logits, self.next_tokens)
        self.seq_len.fill_(S)
        self._host_len = S
        self.tokens.copy_(self.next_tokens)
        return self.next_tokens

    def detect_language(self, sot, language_lo, language_hi, no_speech=None):
        """Forward pass over a lone start-of-transcript token (decoding.py:712-719), then on the device: the most
        probable language token and the softmax over the language tokens [language_lo, language_hi) (:721-725), and the
        no-speech probability = softmax over the whole vocabulary at `no_speech` (:762-766).
        Returns (language_tokens int32 [B], language_probs fp32 [B, n_languages], no_speech_probs fp32 [B] or None);
        the decoder state is reset afterwards, the logit-filter state is not touched."""
        filt = getattr(self, "logit_filter", None)
        self.logit_filter = None
        try:
            self.reset()
            self.prefill([[int(sot)]] * self.B)
        finally:
            self.logit_filter = filt
        self.reset()
        n = language_hi - language_lo
        lang = torch.empty((self.B,), dtype=torch.int32, device=self.device)
        probs = torch.empty((self.B, n), dtype=torch.float32, device=self.device)
        nsp = torch.empty((self.B,), dtype=torch.float32, device=self.device) if no_speech is not None else None
        rc = self.lib.b200_logits_range_softmax(self.logits.data_ptr(), self.B, self.V, language_lo, language_hi,
                                                -1 if no_speech is None else int(no_speech), lang.data_ptr(),
                                                probs.data_ptr(), None if nsp is None else nsp.data_ptr(), self._st())
        _lib.check(rc, "logits_range_softmax")
        return lang, probs, nsp

    def _step_body(self):
        B = self.B
        if self.step_kernel:
            x = self._buf("x", B, self.d)
            self._step_kernel_call(x)
            self._head(x, B, self.logits, self.next_tokens)
            self.seq_len.add_(1)
            self.tokens.copy_(self.next_tokens)
            return
        self.lib.b200_set_static_kv_hint(1 if self.static_kv else 0)
        x = self._buf("x", B, self.d)
        _lib.check(self.lib.b200_embed_tokens_fp16(self.tokens.data_ptr(), self.seq_len.data_ptr(),
                                                   self.tok_emb.data_ptr(), self.pos_emb.data_ptr(), x.data_ptr(), B,
                                                   self.d, self.V, self.Smax, self._st()), "embed")
        if self.n_chains == 1:
            self._stack(x, B, 1, context=False)
        else:
            nb = B // self.n_chains
            main = torch.cuda.current_stream(self.device)
            for c, s_c in enumerate(self._chain_streams, start=1):
                s_c.wait_stream(main)  # fork after the embedding
                with torch.cuda.stream(s_c):
                    self._stack(x[c * nb:(c + 1) * nb], nb, 1, context=False, b0=c * nb, nb=nb, chain=c)
            self._stack(x[:nb], nb, 1, context=False, b0=0, nb=nb, chain=0)
            for s_c in self._chain_streams:
                main.wait_stream(s_c)  # join before the (whole-batch) logits
        self._head(x, B, self.logits, self.next_tokens)
        self.lib.b200_set_static_kv_hint(0)
        self.seq_len.add_(1)
        self.tokens.copy_(self.next_tokens)

    def _snapshot(self):
        snap = [self.seq_len.clone(), self.tokens.clone()]
        filt = getattr(self, "logit_filter", None)
        if filt is not None:
            snap += [filt.state.clone(), filt.sum_logprobs.clone()]
        return snap

    def _restore(self, snap):
        self.seq_len.copy_(snap[0])
        self.tokens.copy_(snap[1])
        filt = getattr(self, "logit_filter", None)
        if filt is not None:
            filt.state.copy_(snap[2])
            filt.sum_logprobs.copy_(snap[3])

    def capture(self):
        """Captures one generation step in a CUDA graph (all shapes static; lengths, tokens and the logit-filter state
        live on the device)."""
        # warm-up outside capture: sets function attributes, allocates buffers
        snap = self._snapshot()
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self._step_body()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self._restore(snap)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step_body()
        torch.cuda.synchronize(self.device)
        self._restore(snap)
        self.graph = g
        # the same step with its host traffic inside the graph: pinned token ids -> device, step, next ids -> pinned.
        # step_host() then costs one graph launch and one stream synchronize.
        gh = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gh):
            self.tokens.copy_(self._pinned_in, non_blocking=True)
            self._step_body()
            self._pinned_out.copy_(self.next_tokens, non_blocking=True)
        torch.cuda.synchronize(self.device)
        self._restore(snap)
        self.graph_host = gh
        return g

    def rewind(self, length):
        """Sets every sequence back to `length` cached tokens (benchmarks: a long run of steps stays inside n_text_ctx
        and keeps a bounded self-attention length).  The cache keeps its bytes; only the lengths move."""
        if not 0 < length <= self.Smax:
            raise ValueError(f"length {length} outside (0, {self.Smax}]")
        self.seq_len.fill_(length)
        self._host_len = length

    def step(self):
        """One greedy generation step for the whole batch; consumes self.tokens, produces self.next_tokens.
        Raises once the text context is full: the reference stops at n_text_ctx (decoding.py:324,749); past it the
        kernels would only clamp (last KV slot overwritten, last position reused) and return garbage silently."""
        if getattr(self, "_host_len", 0) >= self.Smax:
            raise RuntimeError(f"the text context is full ({self.Smax} tokens): reset() or prefill() before stepping on")
        self._host_len = getattr(self, "_host_len", 0) + 1
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_body()
        return self.next_tokens

    def step_host(self, tokens_host=None):
        """End-to-end step with HOST buffers: pinned token ids in -> pinned next-token ids out."""
        if tokens_host is not None and getattr(self, "graph_host", None) is not None:
            if getattr(self, "_host_len", 0) >= self.Smax:
                raise RuntimeError(f"the text context is full ({self.Smax} tokens): reset() or prefill() before stepping on")
            self._host_len = getattr(self, "_host_len", 0) + 1
            self._pinned_in.copy_(torch.as_tensor(tokens_host, dtype=torch.int32))
            self.graph_host.replay()
            torch.cuda.current_stream(self.device).synchronize()
            return self._pinned_out
        if tokens_host is not None:
            self._pinned_in.copy_(torch.as_tensor(tokens_host, dtype=torch.int32))
            self.tokens.copy_(self._pinned_in, non_blocking=True)
        self.step()
        self._pinned_out.copy_(self.next_tokens, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._pinned_out

    def decode(self, prompt_tokens, n_new, use_graph=True):
        """Greedy decode (logit filters off): returns int32 [B, n_new] on the device."""
        S = len(prompt_tokens[0]) if not torch.is_tensor(prompt_tokens) else prompt_tokens.shape[-1]
        if S + n_new - 1 > self.Smax:
            raise ValueError(f"prompt ({S}) + {n_new} new tokens do not fit n_text_ctx = {self.Smax}")
        self.reset()
        out = torch.empty((self.B, n_new), dtype=torch.int32, device=self.device)
        out[:, 0] = self.prefill(prompt_tokens)
        if use_graph and self.graph is None and n_new > 1:
            self.capture()
        for t in range(1, n_new):
            if use_graph:
                out[:, t] = self.step()
            else:
                self._host_len += 1
                self._step_body()
                out[:, t] = self.next_tokens
        return out
