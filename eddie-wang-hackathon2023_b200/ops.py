"""Weight preparation ops -- mirror of the torch custom ops the reference registers from libth_common.so
(T/cpp/tensorrt_llm/thop/weightOnlyQuantOp.cpp:143-236,343-371), as called by T/examples/whisper/weight.py:76-80:

    processed_int8, scales = torch.ops.fastertransformer.symmetric_quantize_last_axis_of_batched_matrix(w_kn_cpu, torch.int8)

The reference runs single-threaded host loops; here the same bit-exact result comes from CUDA kernels
(csrc/quantize.cu).  CPU tensors in -> CPU tensors out like the reference; CUDA tensors in -> CUDA tensors out.
"""
import torch

from . import _lib


def _check_weight(weight, quant_type):
    if quant_type != torch.int8:
        raise ValueError("only torch.int8 weight-only quantization is on the B200 hot path (int4 is out of scope)")
    if weight.dim() != 2:
        raise ValueError("Invalid dim. The dim of weight should be 2 ([K, N]); batched (3-D) weights are unsupported")
    if weight.numel() == 0:
        raise ValueError("weight should not be empty tensor")
    if weight.dtype not in (torch.float16, torch.float32):
        raise ValueError("Invalid datatype. Weight must be FP16 or FP32")
    if not weight.is_contiguous():
        raise ValueError("weight must be contiguous")


def _symmetric_quantize(weight, quant_type, return_unprocessed):
    _check_weight(weight, quant_type)
    lib = _lib.load()
    K, N = weight.shape
    wd = _lib.DTYPE_F16 if weight.dtype == torch.float16 else _lib.DTYPE_F32
    dev = weight.device
    proc = torch.empty((K, N), dtype=torch.int8, device=dev)
    raw = torch.empty((K, N), dtype=torch.int8, device=dev) if return_unprocessed else None
    scales = torch.empty((N,), dtype=weight.dtype, device=dev)  # reference: scales have the weight's dtype
    if dev.type == "cuda":
        rc = lib.b200_symmetric_quantize_int8(_lib.ptr(weight), wd, K, N, _lib.ptr(proc), _lib.ptr(raw),
                                              _lib.ptr(scales), wd, _lib.stream_ptr())
    else:
        rc = lib.b200_symmetric_quantize_int8_host(_lib.ptr(weight), wd, K, N, _lib.ptr(proc), _lib.ptr(raw),
                                                   _lib.ptr(scales), wd)
    _lib.check(rc, "symmetric_quantize")
    if return_unprocessed:
        return [raw, proc, scales]
    return [proc, scales]


def symmetric_quantize_last_axis_of_batched_matrix(weight, quant_type=torch.int8):
    """-> [processed int8 [K, N], scales [N]]   (weightOnlyQuantOp.cpp:224-227)"""
    return _symmetric_quantize(weight, quant_type, False)


def _symmetric_quantize_last_axis_of_batched_matrix(weight, quant_type=torch.int8):
    """-> [raw int8 [K, N], processed int8 [K, N], scales [N]]   (weightOnlyQuantOp.cpp:229-235)"""
    return _symmetric_quantize(weight, quant_type, True)


def preprocess_weights_for_mixed_gemm(row_major_quantized_weight, quant_type=torch.int8):
    """raw int8 [K, N] -> processed layout (weightOnlyQuantOp.cpp:112-141)."""
    w = row_major_quantized_weight
    if quant_type != torch.int8 or w.dtype != torch.int8:
        raise ValueError("Quantized tensor must be int8 dtype")
    if w.dim() != 2 or not w.is_contiguous():
        raise ValueError("Invalid dim. The dim of weight should be 2 and contiguous")
    lib = _lib.load()
    K, N = w.shape
    proc = torch.empty_like(w)
    if w.device.type == "cuda":
        rc = lib.b200_preprocess_weights_int8(_lib.ptr(w), K, N, _lib.ptr(proc), _lib.stream_ptr())
    else:
        rc = lib.b200_preprocess_weights_int8_host(_lib.ptr(w), K, N, _lib.ptr(proc))
    _lib.check(rc, "preprocess_weights_for_mixed_gemm")
    return proc
