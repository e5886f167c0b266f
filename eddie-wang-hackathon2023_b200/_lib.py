"""ctypes binding of libb200_whisper.so (the C ABI declared in include/b200_whisper.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_PKG, "lib", "libb200_whisper.so")
_lib = None

OK = 0
DTYPE_F32, DTYPE_F16, DTYPE_I8, DTYPE_I32 = 0, 1, 2, 3
ACT_NONE, ACT_GELU_ERF, ACT_GELU_TANH = 0, 1, 2

_vp = ctypes.c_void_p
_i = ctypes.c_int
_sz = ctypes.c_size_t
_f = ctypes.c_float


class MmhaParams(ctypes.Structure):
    """b200_mmha_params (include/b200_whisper.h)."""
    _fields_ = [
        ("qkv", _vp), ("qkv_bias", _vp), ("out", _vp), ("kv_cache", _vp), ("sequence_lengths", _vp),
        ("masked_tokens", _vp), ("kv_scale_orig_quant", _vp), ("kv_scale_quant_orig", _vp),
        ("batch_size", ctypes.c_int32), ("num_heads", ctypes.c_int32), ("head_size", ctypes.c_int32),
        ("max_seq_len", ctypes.c_int32), ("past_kv_length", ctypes.c_int32), ("int8_kv_cache", ctypes.c_int32),
        ("q_scaling", ctypes.c_float),
    ]


class DecoderLayer(ctypes.Structure):
    """b200_decoder_layer (include/b200_whisper.h): 32 device pointers per decoder layer."""
    _names = [
        "attn_ln_gamma", "qkv_w", "qkv_scales", "qkv_bias", "qkv_c1s", "qkv_c2",
        "attn_out_w", "attn_out_scales", "attn_out_bias",
        "cross_ln_gamma", "cross_q_w", "cross_q_scales", "cross_q_bias", "cross_q_c1s", "cross_q_c2",
        "cross_out_w", "cross_out_scales", "cross_out_bias",
        "mlp_ln_gamma", "fc1_w", "fc1_scales", "fc1_bias", "fc1_c1s", "fc1_c2",
        "fc2_w", "fc2_scales", "fc2_bias",
        "self_kv", "kv_scale_orig_quant", "kv_scale_quant_orig",
        "cross_kv", "cross_kv_scale_quant_orig",
    ]
    _fields_ = [(n, _vp) for n in _names]


class DecoderStepParams(ctypes.Structure):
    """b200_decoder_step_params (include/b200_whisper.h)."""
    _fields_ = [
        ("layers", _vp),
        ("n_layers", ctypes.c_int32), ("batch_size", ctypes.c_int32), ("num_heads", ctypes.c_int32),
        ("d_ff", ctypes.c_int32), ("max_seq_len", ctypes.c_int32), ("enc_len", ctypes.c_int32),
        ("vocab", ctypes.c_int32), ("n_ctx", ctypes.c_int32),
        ("tokens", _vp), ("sequence_lengths", _vp), ("tok_emb", _vp), ("pos_emb", _vp), ("x_out", _vp),
        ("scratch", _vp), ("ln_eps", ctypes.c_float), ("max_ctas", ctypes.c_int32),
    ]


_SIGS = {
    "b200_last_error": (ctypes.c_char_p, []),
    "b200_abi_version": (_i, []),
    "b200_launch_count": (ctypes.c_ulonglong, []),
    "b200_init": (_i, []),
    "b200_set_pdl": (_i, [_i]),
    "b200_set_static_kv_hint": (_i, [_i]),
    "b200_l2_prefetch": (_i, [_vp, _sz, _vp]),
    "b200_symmetric_quantize_int8": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "b200_preprocess_weights_int8": (_i, [_vp, _i, _i, _vp, _vp]),
    "b200_symmetric_quantize_int8_host": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i]),
    "b200_preprocess_weights_int8_host": (_i, [_vp, _i, _i, _vp]),
    "b200_woq_workspace_bytes": (_sz, [_i, _i, _i]),
    "b200_woq_int8_gemm": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "b200_woq_int8_gemm_fused": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "b200_woq_int8_gemm_ln_fused": (_i, [_vp, _vp, _vp, _f, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "b200_woq_ln_fold_prepare": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "b200_woq_int8_gemm_ln_folded": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _sz,
                                          _vp]),
    "b200_woq_set_kernel_policy": (_i, [_i]),
    "b200_debug_woq_plan": (_i, [_i, _i, _i, _vp]),
    "b200_debug_tc_timing": (_i, [_vp]),
    "b200_debug_tc_timing_filter": (_i, [_i, _i]),
    "b200_debug_tc_timeline": (_i, [_vp, _i]),
    "b200_mmha_generation": (_i, [ctypes.POINTER(MmhaParams), _vp]),
    "b200_qkv_mmha_decode_supported": (_i, [_i, _i, _i]),
    "b200_cross_attention_qproj_supported": (_i, [_i, _i, _i, _i]),
    "b200_debug_xa_timeline": (_i, [_vp]),
    "b200_set_cross_attention_split": (_i, [_i]),
    "b200_cross_attention_qproj": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "b200_qkv_mmha_decode": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "b200_attention_context": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "b200_mmha_generation_paged": (_i, [ctypes.POINTER(MmhaParams), _vp, _i, _i, _vp]),
    "b200_attention_context_paged": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "b200_cross_attention_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "b200_cross_attention": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "b200_cross_kv_pack": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "b200_conv1d_fp16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "b200_transpose_add_pos_fp16": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "b200_attention_bidirectional_fp16": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "b200_whisper_filtered_argmax": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "b200_logits_range_softmax": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "b200_conv1d_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "b200_conv1d_fp16_tc": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "b200_log_mel_frames": (_i, [_i, _i]),
    "b200_log_mel_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "b200_log_mel_spectrogram": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i, _vp, _sz, _vp]),
    "b200_layernorm_fp16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "b200_embed_tokens_fp16": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "b200_logits_workspace_bytes": (_sz, [_i, _i]),
    "b200_logits_argmax_fp16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "b200_logits_set_kernel_policy": (_i, [_i]),
    "b200_decoder_step_scratch_bytes": (_sz, [_i, _i]),
    "b200_decoder_step": (_i, [ctypes.POINTER(DecoderStepParams), _vp]),
    "b200_decoder_step_status": (_i, [_vp, ctypes.POINTER(ctypes.c_int32)]),
    "b200_debug_decoder_step_timeline": (_i, [_vp]),
}


def lib_path():
    return _LIB_PATH


def load():
    """Loads the native library; raises if it has not been built (python -m b200_whisper._build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"{_LIB_PATH} not found: build it with `python __graft_entry__.py build` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for the hot path.")
    lib = ctypes.CDLL(_LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def declared_symbols():
    return sorted(_SIGS)


def check(rc, what=""):
    if rc != OK:
        msg = load().b200_last_error().decode(errors="replace")
        raise RuntimeError(f"{what or 'b200 call'} failed (code {rc}): {msg}")


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream
