"""CPU tests: the oracle's quantizer/layout restatement against (a) the committed golden vectors produced by the
reference's own cutlass_preprocessors.cpp and (b) that binary itself when oracle/_ref is present.
Reference spec sources: T/tests/quantization/test_weight_only_quant_matmul.py:112-130, _utils.py:15-22,
T/cpp/tensorrt_llm/kernels/cutlass_kernels/cutlass_preprocessors.cpp:154-157 (permutation KAT)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import woq

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gen_weight(k, n, dtype=torch.float16, seed=0):
    torch.manual_seed(seed)
    return (torch.rand((k, n), dtype=dtype) * 2 - 1.0).numpy()


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_golden_full_vectors():
    g = np.load(os.path.join(GOLD, "quant_golden.npz"))
    keys = sorted({k.rsplit("_", 1)[0] for k in g.files})
    assert len(keys) == 8
    for key in keys:
        w = g[key + "_w"]
        raw, proc, scales = woq.symmetric_quantize_int8(w)
        assert np.array_equal(raw, g[key + "_raw"]), key
        assert np.array_equal(proc, g[key + "_proc"]), key
        assert np.array_equal(scales.view(np.uint16), g[key + "_scales"].view(np.uint16)), key
        # closed form and inverse agree with the reference bytes too
        assert np.array_equal(woq.preprocess_weights_int8(raw, closed_form=True), g[key + "_proc"])
        assert np.array_equal(woq.unprocess_int8(g[key + "_proc"]), raw)


@pytest.mark.parametrize("k,n", [(4096, 1024), (512, 4096), (1280, 3840), (5120, 1280), (384, 1152)])
def test_golden_digests(k, n):
    with open(os.path.join(GOLD, "quant_digests.json")) as f:
        d = json.load(f)[f"k{k}_n{n}_f16"]
    raw, proc, scales = woq.symmetric_quantize_int8(gen_weight(k, n))
    assert digest(raw) == d["raw"]
    assert digest(proc) == d["proc"]
    assert digest(scales) == d["scales"]


def test_row_permutation_kat():
    # cutlass_preprocessors.cpp:154-157: "0 1 8 9 2 3 10 11 4 5 12 13 6 7 14 15"
    K, N = 64, 64
    raw = np.repeat(np.arange(K, dtype=np.int8)[:, None], N, axis=1)
    out = np.empty_like(raw)
    woq.lib().oracle_permute_B_rows_int8(out.ctypes.data, raw.ctypes.data, K, N)
    assert out[:16, 0].tolist() == [0, 1, 8, 9, 2, 3, 10, 11, 4, 5, 12, 13, 6, 7, 14, 15]
    assert out[16:32, 5].tolist() == [16 + v for v in [0, 1, 8, 9, 2, 3, 10, 11, 4, 5, 12, 13, 6, 7, 14, 15]]


def test_bias_interleave_kat():
    # cutlass_preprocessors.cpp:392-398: [e3 e2 e1 e0] -> [e3 e1 e2 e0], values + 128
    buf = np.array([0, 1, 2, 3, -128, 127, -1, 5], np.int8)
    woq.lib().oracle_add_bias_and_interleave_int8_inplace(buf.ctypes.data, buf.size)
    assert buf.view(np.uint8).tolist() == [128, 130, 129, 131, 0, 127, 255, 133]


def test_shape_checks():
    with pytest.raises(ValueError):
        woq.preprocess_weights_int8(np.zeros((64, 32), np.int8))  # N % 64 (cutlass_preprocessors.cpp:498-500)
    with pytest.raises(ValueError):
        woq.preprocess_weights_int8(np.zeros((24, 64), np.int8))  # K % 16 / tile (:187-192)


def test_extremes_and_rounding():
    # half-away-from-zero rounding and clamp to [-128, 127] (cutlass_preprocessors.cpp:683-687)
    K, N = 64, 64
    w = np.zeros((K, N), np.float32)
    w[0, :] = 128.0          # amax -> scale 1.0
    w[1, :] = 0.5            # round(0.5) = 1 (away from zero), not 0 (half-even)
    w[2, :] = -0.5           # -> -1
    w[3, :] = 2.5            # -> 3
    w[4, :] = -128.0         # -> -128
    raw, _, scales = woq.symmetric_quantize_int8(w)
    assert scales.astype(np.float32)[0] == 1.0
    assert raw[0, 0] == 127 and raw[1, 0] == 1 and raw[2, 0] == -1 and raw[3, 0] == 3 and raw[4, 0] == -128


@pytest.mark.skipif(woq.ref_lib() is None, reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("k,n,dt", [(64, 64, np.float16), (256, 192, np.float32), (1280, 1280, np.float16)])
def test_against_reference_binary(k, n, dt):
    rng = np.random.default_rng(k * 7 + n)
    w = (rng.standard_normal((k, n)) * 0.02).astype(dt)
    r0, p0, s0 = woq.symmetric_quantize_int8(w)
    r1, p1, s1 = woq.ref_symmetric_quantize_int8(w)
    assert np.array_equal(r0, r1) and np.array_equal(p0, p1)
    assert np.array_equal(s0.view(np.uint16), s1.view(np.uint16))


def test_kv_quant_rounding():
    # cvt.rni.sat.s8.f32: round-half-even + saturate (decoderMaskedMultiheadAttentionUtils.h:2276-2286;
    # expected value of T/cpp/tests/runtime/transposeKVKernelTest.cpp:79-84)
    x = np.array([0.5, 1.5, 2.5, -0.5, -1.5, 300.0, -300.0, 126.5, 127.5], np.float16)
    q = woq.kv_quantize_int8(x, 1.0)
    assert q.tolist() == [0, 2, 2, 0, -2, 127, -128, 126, 127]
    d = woq.kv_dequantize_int8(np.array([-128, -1, 0, 1, 127], np.int8), 0.0123)
    assert np.array_equal(d, (np.float32(0.0123) * np.array([-128, -1, 0, 1, 127], np.float32)).astype(np.float16))
