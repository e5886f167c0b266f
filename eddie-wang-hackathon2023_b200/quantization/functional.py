"""weight_only_quant_matmul -- mirror of T/tensorrt_llm/quantization/functional.py:51-74.

In the reference this adds a `WeightOnlyQuantMatmul` plugin layer to a TensorRT network; here it runs the same
plugin contract eagerly on torch CUDA tensors: shape inference and dispatch follow
WeightOnlyQuantMatmulPlugin::getOutputDimensions / enqueue
(T/cpp/tensorrt_llm/plugins/weightOnlyQuantMatmulPlugin/weightOnlyQuantMatmulPlugin.cpp:73-110,162-222).
"""
import torch

from .. import _lib

_ws_cache = {}


def _workspace(nbytes, device):
    """Per-device scratch (TensorRT hands the plugin a workspace; eager callers get a cached one)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _as_int8_kn(weights, k):
    """The reference stores int8 weights as a float32 tensor [K, N/4] because TensorRT plugins could not take int8
    inputs (weightOnlyQuantMatmulPlugin.cpp:120-129, examples/whisper/weight.py:79-80).  Accept both views."""
    if weights.dtype == torch.float32:
        if weights.dim() != 2 or weights.shape[0] != k:
            raise ValueError(f"weights must be [K={k}, N/4] float32 (int8 bytes viewed as float)")
        return weights.view(torch.int8)
    if weights.dtype == torch.int8:
        if weights.dim() != 2 or weights.shape[0] != k:
            raise ValueError(f"weights must be [K={k}, N] int8")
        return weights
    raise TypeError("weights must be int8 [K, N] or its float32 view [K, N/4]")


def weight_only_quant_matmul(input, weights, scales, weightTypeId, bias=None, activation=None, residual=None,
                             out=None):
    """input [..., K] fp16 x preprocessed int8 weights -> [..., N] fp16.

    weightTypeId: 1 = int8 weight-only (2 = int4 is not on the B200 hot path).  `bias`, `activation`
    ('gelu' | 'gelu_tanh') and `residual` are B200 extensions fused into the kernel epilogue; the reference adds the
    bias with a separate layer (quantization/layer.py:311-312).
    """
    if weightTypeId != 1:
        raise TypeError("Weight Only Quant MatMul: only weightTypeId=1 (int8) is supported on the B200 hot path")
    if input.dtype != torch.float16 or scales.dtype != torch.float16:
        raise TypeError("Weight Only Quant MatMul is only supported with float16 activations and scales")
    if not input.is_cuda:
        raise RuntimeError("weight_only_quant_matmul needs CUDA tensors (there is no CPU fallback)")
    lib = _lib.load()
    k = input.shape[-1]
    w = _as_int8_kn(weights, k)
    n = w.shape[1]
    if scales.numel() != n:
        raise ValueError(f"scales must have N={n} elements")
    x = input.contiguous()
    m = x.numel() // k
    if out is None:
        out = torch.empty(*input.shape[:-1], n, dtype=torch.float16, device=input.device)
    act = {None: _lib.ACT_NONE, "gelu": _lib.ACT_GELU_ERF, "gelu_tanh": _lib.ACT_GELU_TANH}[activation]
    ws_bytes = lib.b200_woq_workspace_bytes(max(m, 1), n, k)
    ws = _workspace(ws_bytes, input.device)
    rc = lib.b200_woq_int8_gemm_fused(_lib.ptr(x), m, k, _lib.ptr(w), _lib.ptr(scales), n, _lib.ptr(bias), act,
                                      _lib.ptr(residual), _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "weight_only_quant_matmul")
    return out
