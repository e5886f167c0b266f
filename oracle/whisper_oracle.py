"""oracle/whisper_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU, torch fp32).

Restatement of the reference's PyTorch Whisper (T/examples/whisper/torch_model.py) as plain functions over an
OpenAI-style state dict, plus the "quantized semantics" the B200 path implements:

  * Linear weights replaced by their int8 weight-only dequantization  w16 = fp16(q * s16)
    (T/cpp/tensorrt_llm/kernels/weightOnlyMatrixVectorMultiplication.cu:44-53) -- "identically dequantized weights";
  * self-attention K/V cache int8 round trip for PAST tokens only; the current step uses the unquantized k, v
    (T/cpp/tensorrt_llm/kernels/decoderMaskedMultiheadAttention/decoderMaskedMultiheadAttentionTemplate.h:1503,1517,1920,1933;
    quant/dequant rules decoderMaskedMultiheadAttentionUtils.h:2276-2286,2357-2365);
  * cross-attention K/V int8 round trip for all encoder frames.

Function map (reference file:line):
  layer_norm            torch_model.py:25-27
  attention (qkv)       torch_model.py:88-103   (q*Dh^-.25, k*Dh^-.25, fp32 softmax)
  residual block        torch_model.py:125-138
  encoder               torch_model.py:152-171
  decoder               torch_model.py:196-218  (logits = x @ token_embedding^T)
  greedy loop           T/examples/whisper/decoding.py:743-783 (torch_main_loop) with GreedyDecoder.update :279-295,
                        logit filters off (raw argmax), as SURVEY.md 8d prescribes for kernel-parity runs.

Pinning: tests/golden/make_whisper_golden.py runs the REFERENCE torch_model.Whisper (imported from /root/reference)
on seeded synthetic weights and stores its logits/tokens; tests/test_oracle_whisper.py checks this restatement
against those fixtures (and against the reference directly when /root/reference is present).
"""
import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from . import woq


@dataclass
class ModelDimensions:  # same field order as torch_model.py:12-22
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int


LARGE_V2 = ModelDimensions(80, 1500, 1280, 20, 32, 51865, 448, 1280, 20, 32)
TINY = ModelDimensions(80, 1500, 384, 6, 4, 51865, 448, 384, 6, 4)
# small enough to keep golden fixtures tiny, same structure (head size 64, dims % 64 == 0)
MICRO = ModelDimensions(80, 96, 128, 2, 2, 1024, 64, 128, 2, 2)


def sinusoids(length, channels, max_timescale=10000):
    """torch_model.py:48-54"""
    log_timescale_increment = np.log(max_timescale) / (channels // 2 - 1)
    inv_timescales = torch.exp(-log_timescale_increment * torch.arange(channels // 2))
    scaled_time = torch.arange(length)[:, np.newaxis] * inv_timescales[np.newaxis, :]
    return torch.cat([torch.sin(scaled_time), torch.cos(scaled_time)], dim=1)


def synthetic_state_dict(dims: ModelDimensions, seed=0, dtype=torch.float32, decoder_only=False):
    """Seeded synthetic weights with the reference checkpoint's key names ({dims, model_state_dict} of large-v2.pt,
    T/examples/whisper/build.py:394).  Default-init Whisper emits one constant token and `positional_embedding` is
    torch.empty in the reference (torch_model.py:178), so every tensor is drawn explicitly: Linear N(0, 1/sqrt(in)),
    LayerNorm gamma ~ 1, token embedding N(0, 0.1), positions N(0, 0.5) -- activations stay O(1) through the pre-LN
    stack, the input token does not dominate the logits and the greedy tokens vary.
    """
    g = torch.Generator().manual_seed(seed)

    def randn(*shape, std=1.0):
        return (torch.randn(*shape, generator=g) * std).to(dtype)

    sd = {}

    def linear(prefix, n_out, n_in, bias=True, gain=1.0):
        sd[prefix + ".weight"] = randn(n_out, n_in, std=gain / math.sqrt(n_in))
        if bias:
            sd[prefix + ".bias"] = randn(n_out, std=0.02)

    def ln(prefix, n):
        sd[prefix + ".weight"] = (1.0 + 0.05 * torch.randn(n, generator=g)).to(dtype)
        sd[prefix + ".bias"] = randn(n, std=0.02)

    def block(prefix, n_state, cross):
        for a in (["attn", "cross_attn"] if cross else ["attn"]):
            linear(f"{prefix}.{a}.query", n_state, n_state)
            linear(f"{prefix}.{a}.key", n_state, n_state, bias=False)
            linear(f"{prefix}.{a}.value", n_state, n_state)
            linear(f"{prefix}.{a}.out", n_state, n_state, gain=1.0)
            ln(f"{prefix}.{a}_ln", n_state)
        linear(f"{prefix}.mlp.0", 4 * n_state, n_state)
        linear(f"{prefix}.mlp.2", n_state, 4 * n_state, gain=1.0)
        ln(f"{prefix}.mlp_ln", n_state)

    if not decoder_only:
        d = dims.n_audio_state
        sd["encoder.conv1.weight"] = randn(d, dims.n_mels, 3, std=1.0 / math.sqrt(3 * dims.n_mels))
        sd["encoder.conv1.bias"] = randn(d, std=0.02)
        sd["encoder.conv2.weight"] = randn(d, d, 3, std=1.0 / math.sqrt(3 * d))
        sd["encoder.conv2.bias"] = randn(d, std=0.02)
        sd["encoder.positional_embedding"] = sinusoids(dims.n_audio_ctx, d).to(dtype)
        for i in range(dims.n_audio_layer):
            block(f"encoder.blocks.{i}", d, cross=False)
        ln("encoder.ln_post", d)
    d = dims.n_text_state
    sd["decoder.token_embedding.weight"] = randn(dims.n_vocab, d, std=0.1)
    sd["decoder.positional_embedding"] = randn(dims.n_text_ctx, d, std=0.5)
    for i in range(dims.n_text_layer):
        block(f"decoder.blocks.{i}", d, cross=True)
    ln("decoder.ln", d)
    return sd


# ---- quantized-semantics helpers ----------------------------------------------------------------

def dequantized_linear_weight(w_out_in: torch.Tensor):
    """torch Linear weight [out, in] -> the effective fp16 weight the int8 weight-only kernels multiply by, as fp32
    [out, in].  Quantization follows examples/whisper/weight.py:76-80 (W^T [K, N] contiguous, fp16)."""
    w_kn = np.ascontiguousarray(w_out_in.detach().to(torch.float16).t().contiguous().numpy())
    raw, _, scales = woq.symmetric_quantize_int8(w_kn, np.float16)
    w16 = (raw.astype(np.float16) * scales[None, :]).astype(np.float16)  # fp16(fp16(q) * s16), one rounding
    return torch.from_numpy(w16.astype(np.float32)).t().contiguous()


def dequantized_linear_weight_cat(ws):
    """Fused qkv: the reference quantizes cat([q, k, v], dim=0)^T, i.e. per output column -- identical to quantizing
    each projection separately."""
    return [dequantized_linear_weight(w) for w in ws]


def kv_int8_roundtrip(x: torch.Tensor, scale_quant_orig: float):
    """x fp32 holding fp16-representable values -> dequant(quant(x)) as fp32 (int8 KV cache convention)."""
    x16 = x.to(torch.float16).numpy()
    q = woq.kv_quantize_int8(x16, np.float32(1.0) / np.float32(scale_quant_orig))
    return torch.from_numpy(woq.kv_dequantize_int8(q, scale_quant_orig).astype(np.float32)).reshape(x.shape)


def quantize_state_dict(sd, dims: ModelDimensions, decoder_only=True):
    """Returns a copy where every decoder (and optionally encoder) Linear weight is replaced by its int8 weight-only
    dequantization; everything else is rounded to fp16 (the GPU path stores fp16 parameters)."""
    out = {}
    for k, v in sd.items():
        is_linear_w = k.endswith(".weight") and v.dim() == 2 and "token_embedding" not in k
        in_scope = k.startswith("decoder.") or not decoder_only
        if is_linear_w and in_scope:
            out[k] = dequantized_linear_weight(v)
        else:
            out[k] = v.to(torch.float16).to(torch.float32)
    return out


# ---- model --------------------------------------------------------------------------------------

def layer_norm(x, sd, prefix):
    return F.layer_norm(x.float(), (x.shape[-1],), sd[prefix + ".weight"].float(), sd[prefix + ".bias"].float(), 1e-5)


def linear(x, sd, prefix):
    b = sd.get(prefix + ".bias")
    return F.linear(x, sd[prefix + ".weight"].float(), None if b is None else b.float())


def qkv_attention(q, k, v, n_head, mask=None):
    """torch_model.py:88-103"""
    n_batch, n_ctx, n_state = q.shape
    scale = (n_state // n_head) ** -0.25
    q = q.view(*q.shape[:2], n_head, -1).permute(0, 2, 1, 3) * scale
    k = k.view(*k.shape[:2], n_head, -1).permute(0, 2, 3, 1) * scale
    v = v.view(*v.shape[:2], n_head, -1).permute(0, 2, 1, 3)
    qk = q @ k
    if mask is not None:
        qk = qk + mask
    w = F.softmax(qk.float(), dim=-1)
    return (w @ v).permute(0, 2, 1, 3).flatten(start_dim=2)


def encoder_forward(sd, dims: ModelDimensions, mel):
    """torch_model.py:152-171.  mel [B, n_mels, T]."""
    x = F.gelu(F.conv1d(mel.float(), sd["encoder.conv1.weight"].float(), sd["encoder.conv1.bias"].float(), padding=1))
    x = F.gelu(F.conv1d(x, sd["encoder.conv2.weight"].float(), sd["encoder.conv2.bias"].float(), stride=2, padding=1))
    x = x.permute(0, 2, 1)
    x = x + sd["encoder.positional_embedding"].float()
    for i in range(dims.n_audio_layer):
        p = f"encoder.blocks.{i}"
        h = layer_norm(x, sd, p + ".attn_ln")
        a = qkv_attention(linear(h, sd, p + ".attn.query"), linear(h, sd, p + ".attn.key"),
                          linear(h, sd, p + ".attn.value"), dims.n_audio_head)
        x = x + linear(a, sd, p + ".attn.out")
        h = layer_norm(x, sd, p + ".mlp_ln")
        x = x + linear(F.gelu(linear(h, sd, p + ".mlp.0")), sd, p + ".mlp.2")
    return layer_norm(x, sd, "encoder.ln_post")


class DecoderState:
    """KV caches of the oracle decoder: self K/V per layer (fp32 tensors of unquantized values) and cross K/V."""

    def __init__(self, n_layer):
        self.k = [None] * n_layer
        self.v = [None] * n_layer
        self.ck = [None] * n_layer
        self.cv = [None] * n_layer
        self.offset = 0


def decoder_forward(sd, dims: ModelDimensions, tokens, xa, state: DecoderState = None, kv_scales=None,
                    cross_kv_scales=None, act_fp16=False):
    """torch_model.py:196-218 with the hook-based KV cache (torch_model.py:270-301) made explicit.

    tokens [B, T] (all prompt tokens on the first call, then one token per call).  Returns fp32 logits [B, T, V].
    kv_scales / cross_kv_scales: per-layer scale_quant_orig (t in weight.py:242-243) enabling the int8 KV semantics;
    None = unquantized (the plain reference model).
    act_fp16: round the activations that cross kernel boundaries on the GPU path to fp16.
    """
    if state is None:
        state = DecoderState(dims.n_text_layer)
    r16 = (lambda t: t.to(torch.float16).to(torch.float32)) if act_fp16 else (lambda t: t)
    B, T = tokens.shape
    off = state.offset
    x = sd["decoder.token_embedding.weight"].float()[tokens] + sd["decoder.positional_embedding"].float()[off:off + T]
    x = r16(x)
    n_head = dims.n_text_head
    for i in range(dims.n_text_layer):
        p = f"decoder.blocks.{i}"
        # --- self attention ---
        h = r16(layer_norm(x, sd, p + ".attn_ln"))
        q = r16(linear(h, sd, p + ".attn.query"))
        k = r16(linear(h, sd, p + ".attn.key"))
        v = r16(linear(h, sd, p + ".attn.value"))
        if state.k[i] is None:
            k_all, v_all = k, v  # context phase: unquantized K/V for the attention itself
            k_past = v_past = None
        else:
            k_past, v_past = state.k[i], state.v[i]
            if kv_scales is not None:
                k_past = kv_int8_roundtrip(k_past, kv_scales[i])
                v_past = kv_int8_roundtrip(v_past, kv_scales[i])
            k_all = torch.cat([k_past, k], dim=1)
            v_all = torch.cat([v_past, v], dim=1)
        L = k_all.shape[1]
        mask = torch.full((T, L), float("-inf")).triu_(L - T + 1)
        a = r16(qkv_attention(q, k_all, v_all, n_head, mask))
        state.k[i] = k if state.k[i] is None else torch.cat([state.k[i], k], dim=1)
        state.v[i] = v if state.v[i] is None else torch.cat([state.v[i], v], dim=1)
        x = r16(x + r16(linear(a, sd, p + ".attn.out")))
        # --- cross attention ---
        h = r16(layer_norm(x, sd, p + ".cross_attn_ln"))
        q = r16(linear(h, sd, p + ".cross_attn.query"))
        if state.ck[i] is None:
            ck = r16(linear(xa.float(), sd, p + ".cross_attn.key"))
            cv = r16(linear(xa.float(), sd, p + ".cross_attn.value"))
            if cross_kv_scales is not None:
                ck = kv_int8_roundtrip(ck, cross_kv_scales[i])
                cv = kv_int8_roundtrip(cv, cross_kv_scales[i])
            state.ck[i], state.cv[i] = ck, cv
        a = r16(qkv_attention(q, state.ck[i], state.cv[i], n_head))
        x = r16(x + r16(linear(a, sd, p + ".cross_attn.out")))
        # --- mlp ---
        h = r16(layer_norm(x, sd, p + ".mlp_ln"))
        u = r16(F.gelu(r16(linear(h, sd, p + ".mlp.0"))))
        x = r16(x + r16(linear(u, sd, p + ".mlp.2")))
    state.offset += T
    x = r16(layer_norm(x, sd, "decoder.ln"))
    logits = x @ sd["decoder.token_embedding.weight"].float().t()
    return logits, state


def greedy_decode(sd, dims, xa, prompt, n_new, kv_scales=None, cross_kv_scales=None, act_fp16=False):
    """Greedy loop with logit filters off: returns (tokens [B, n_new] int64, list of last-position logits)."""
    B = xa.shape[0]
    tokens = torch.tensor(prompt, dtype=torch.long).repeat(B, 1)
    state = None
    out, all_logits = [], []
    cur = tokens
    for _ in range(n_new):
        logits, state = decoder_forward(sd, dims, cur, xa, state, kv_scales, cross_kv_scales, act_fp16)
        last = logits[:, -1]
        nxt = last.argmax(dim=-1)
        all_logits.append(last)
        out.append(nxt)
        cur = nxt[:, None]
    return torch.stack(out, dim=1), all_logits


def calibrate_kv_scales(sd, dims, xa, prompt, n_steps=8):
    """scale_y_quant_orig = max|y| / 127 per layer (T/examples/whisper/utils/convert.py:78 via smoothquant.py:116-175),
    measured on the synthetic run itself: returns (self_scales, cross_scales) lists of python floats."""
    B = xa.shape[0]
    tokens = torch.tensor(prompt, dtype=torch.long).repeat(B, 1)
    state = None
    cur = tokens
    for _ in range(n_steps):
        logits, state = decoder_forward(sd, dims, cur, xa, state)
        cur = logits[:, -1].argmax(dim=-1)[:, None]
    self_s, cross_s = [], []
    for i in range(dims.n_text_layer):
        self_s.append(float(max(state.k[i].abs().max(), state.v[i].abs().max())) / 127.0)
        cross_s.append(float(max(state.ck[i].abs().max(), state.cv[i].abs().max())) / 127.0)
    return self_s, cross_s
