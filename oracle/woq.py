"""oracle/woq.py -- TEST INFRASTRUCTURE.  ctypes bindings for oracle/woq_oracle.c (liboracle.so) and,
when present, the reference's own quantizer compiled into oracle/_ref/libref_quant.so.

Every function cites the reference code it restates in woq_oracle.c's header.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def build(force=False):
    """Compile liboracle.so (gcc) and, if /root/reference exists, _ref/libref_quant.so."""
    if force or not os.path.exists(os.path.join(_HERE, "liboracle.so")) or (
            os.path.getmtime(os.path.join(_HERE, "liboracle.so")) < os.path.getmtime(os.path.join(_HERE, "woq_oracle.c"))):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/tensorrt_llm_july-release-v1") and (
            force or not os.path.exists(os.path.join(_HERE, "_ref", "libref_quant.so"))):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference") and not os.path.exists(os.path.join(_HERE, "_ref", "libref_gpu.so")):
        # the reference's GEMV + MMHA kernels for sm_100a (about 3 minutes, once; GPU tests use them as a second pin)
        try:
            subprocess.call(["make", "-C", _HERE, "refgpu"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                            timeout=900)
        except subprocess.TimeoutExpired:
            pass


def lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = ctypes.CDLL(os.path.join(_HERE, "liboracle.so"))
    return _LIB


def ref_lib():
    """The reference's cutlass_preprocessors.cpp behind extern "C" (None if not built)."""
    global _REF
    if _REF is None:
        p = os.path.join(_HERE, "_ref", "libref_quant.so")
        if not os.path.exists(p):
            try:
                build()
            except Exception:
                pass
        if os.path.exists(p):
            _REF = ctypes.CDLL(p)
    return _REF


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def symmetric_quantize_int8(w: np.ndarray, scale_dtype=np.float16):
    """w: [K, N] float16 or float32.  Returns (raw int8 [K,N], processed int8 [K,N], scales [N])."""
    assert w.ndim == 2 and w.dtype in (np.float16, np.float32)
    w = np.ascontiguousarray(w)
    K, N = w.shape
    raw = np.empty((K, N), np.int8)
    proc = np.empty((K, N), np.int8)
    scales = np.empty((N,), scale_dtype)
    rc = lib().oracle_symmetric_quantize_int8(_p(w), int(w.dtype == np.float16), K, N, _p(raw), _p(proc), _p(scales),
                                              int(scale_dtype == np.float16))
    if rc != 0:
        raise ValueError(f"oracle_symmetric_quantize_int8 rc={rc} (K%64, N%64 required)")
    return raw, proc, scales


def preprocess_weights_int8(raw: np.ndarray, closed_form=False):
    raw = np.ascontiguousarray(raw, dtype=np.int8)
    K, N = raw.shape
    proc = np.empty((K, N), np.int8)
    if closed_form:
        lib().oracle_preprocess_closed_form_int8(_p(proc), _p(raw), K, N)
    else:
        rc = lib().oracle_preprocess_weights_int8(_p(proc), _p(raw), K, N)
        if rc != 0:
            raise ValueError(f"oracle_preprocess_weights_int8 rc={rc}")
    return proc


def unprocess_int8(proc: np.ndarray):
    proc = np.ascontiguousarray(proc, dtype=np.int8)
    K, N = proc.shape
    raw = np.empty((K, N), np.int8)
    lib().oracle_unprocess_int8(_p(raw), _p(proc), K, N)
    return raw


def woq_matmul(a: np.ndarray, raw: np.ndarray, scales: np.ndarray, mode="cutlass"):
    """a [M,K] fp16, raw [K,N] int8, scales [N] fp16 -> [M,N] fp16 under the reference's arithmetic."""
    a = np.ascontiguousarray(a, dtype=np.float16)
    raw = np.ascontiguousarray(raw, dtype=np.int8)
    scales = np.ascontiguousarray(scales, dtype=np.float16)
    M, K = a.shape
    N = raw.shape[1]
    c = np.empty((M, N), np.float16)
    lib().oracle_woq_matmul(_p(a), M, K, _p(raw), _p(scales), N, _p(c), {"cutlass": 0, "gemv": 1, "ideal": 2, "gemv_exact": 3}[mode])
    return c


def kv_quantize_int8(x: np.ndarray, scale_orig_quant: float):
    x = np.ascontiguousarray(x, dtype=np.float16)
    out = np.empty(x.shape, np.int8)
    lib().oracle_kv_quantize_int8(_p(x), ctypes.c_size_t(x.size), ctypes.c_float(scale_orig_quant), _p(out))
    return out


def kv_dequantize_int8(q: np.ndarray, scale_quant_orig: float):
    q = np.ascontiguousarray(q, dtype=np.int8)
    out = np.empty(q.shape, np.float16)
    lib().oracle_kv_dequantize_int8(_p(q), ctypes.c_size_t(q.size), ctypes.c_float(scale_quant_orig), _p(out))
    return out


# ---- the reference itself (oracle/_ref) --------------------------------------------------------

def ref_symmetric_quantize_int8(w: np.ndarray):
    """Runs the REFERENCE's symmetric_quantize<half,half> / <half,float>. Returns (raw, proc, scales fp16)."""
    r = ref_lib()
    if r is None:
        raise RuntimeError("oracle/_ref/libref_quant.so not built")
    w = np.ascontiguousarray(w)
    K, N = w.shape
    raw = np.empty((K, N), np.int8)
    proc = np.empty((K, N), np.int8)
    scales = np.empty((N,), np.float16)
    fn = r.ref_symmetric_quantize_f16 if w.dtype == np.float16 else r.ref_symmetric_quantize_f32w_f16s
    rc = fn(_p(w), K, N, _p(raw), _p(proc), _p(scales))
    if rc != 0:
        raise ValueError("reference symmetric_quantize threw")
    return raw, proc, scales


def ref_preprocess_weights_int8(raw: np.ndarray):
    r = ref_lib()
    raw = np.ascontiguousarray(raw, dtype=np.int8)
    K, N = raw.shape
    proc = np.empty((K, N), np.int8)
    if r.ref_preprocess_weights(_p(raw), K, N, _p(proc)) != 0:
        raise ValueError("reference preprocess_weights_for_mixed_gemm threw")
    return proc


# ---- the reference's own CUDA kernels on the GPU (oracle/_ref/libref_gpu.so, `make -C oracle refgpu`) ----------

_REFGPU = None


def ref_gpu_lib():
    """The reference GEMV and MMHA kernels compiled for sm_100a behind extern "C" (None if not built)."""
    global _REFGPU
    if _REFGPU is None:
        p = os.path.join(_HERE, "_ref", "libref_gpu.so")
        if os.path.exists(p):
            _REFGPU = ctypes.CDLL(p)
            _REFGPU.ref_gpu_gemv.restype = ctypes.c_int
            _REFGPU.ref_gpu_gemv.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
            _REFGPU.ref_gpu_mmha.restype = ctypes.c_int
            _REFGPU.ref_gpu_mmha.argtypes = [ctypes.c_void_p] * 8 + [ctypes.c_int] * 6 + [ctypes.c_float, ctypes.c_void_p]
    return _REFGPU
