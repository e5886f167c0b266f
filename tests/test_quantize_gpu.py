"""GPU parity: the CUDA quantizer / layout kernels against the oracle (bit-exact) and the reference-generated golden
vectors.  Spec: T/tests/quantization/test_weight_only_quant_matmul.py:121-130 (layout), cutlass_preprocessors.cpp:615-721."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import woq

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gen_weight(k, n, dtype=torch.float16, seed=0):
    torch.manual_seed(seed)
    return torch.rand((k, n), dtype=dtype) * 2 - 1.0


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("k,n", [(64, 64), (128, 64), (64, 128), (192, 256)])
@pytest.mark.parametrize("tag,dt", [("f16", torch.float16), ("f32", torch.float32)])
def test_golden_full_vectors(k, n, tag, dt):
    import b200_whisper as bw
    g = np.load(os.path.join(GOLD, "quant_golden.npz"))
    key = f"k{k}_n{n}_{tag}"
    w = torch.from_numpy(g[key + "_w"]).cuda()
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(w, torch.int8)
    assert np.array_equal(raw.cpu().numpy(), g[key + "_raw"])
    assert np.array_equal(proc.cpu().numpy(), g[key + "_proc"])
    # reference <half,float> returns fp16 scales; the torch op returns scales in the weight dtype: compare in fp16
    assert np.array_equal(scales.cpu().to(torch.float16).numpy().view(np.uint16), g[key + "_scales"].view(np.uint16))


@pytest.mark.parametrize("k,n", [(4096, 1024), (512, 4096), (1280, 3840), (1280, 1280), (1280, 5120), (5120, 1280),
                                 (12288, 6144)])
def test_golden_digests_device_and_host_api(k, n):
    import b200_whisper as bw
    with open(os.path.join(GOLD, "quant_digests.json")) as f:
        d = json.load(f)[f"k{k}_n{n}_f16"]
    w = gen_weight(k, n)
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(w.cuda(), torch.int8)
    assert digest(raw.cpu().numpy()) == d["raw"]
    assert digest(proc.cpu().numpy()) == d["proc"]
    assert digest(scales.cpu().numpy()) == d["scales"]
    if k * n <= 1280 * 5120:
        # CPU tensors in -> CPU tensors out, like torch.ops.fastertransformer.* (host-pointer entry points)
        proc_h, scales_h = bw.ops.symmetric_quantize_last_axis_of_batched_matrix(w, torch.int8)
        assert not proc_h.is_cuda and digest(proc_h.numpy()) == d["proc"] and digest(scales_h.numpy()) == d["scales"]


def test_against_oracle_random_normal_and_preprocess_only():
    import b200_whisper as bw
    rng = np.random.default_rng(7)
    for (k, n) in [(384, 1152), (1536, 384), (256, 64)]:
        w = (rng.standard_normal((k, n)) * 0.02).astype(np.float16)
        r0, p0, s0 = woq.symmetric_quantize_int8(w)
        r1, p1, s1 = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(torch.from_numpy(w).cuda(), torch.int8)
        assert np.array_equal(r0, r1.cpu().numpy()) and np.array_equal(p0, p1.cpu().numpy())
        assert np.array_equal(s0.view(np.uint16), s1.cpu().numpy().view(np.uint16))
        p2 = bw.ops.preprocess_weights_for_mixed_gemm(torch.from_numpy(r0).cuda(), torch.int8)
        assert np.array_equal(p0, p2.cpu().numpy())
        # round trip property at any size: un-processing the processed bytes gives the raw matrix back
        assert np.array_equal(woq.unprocess_int8(p2.cpu().numpy()), r0)


def test_extremes():
    import b200_whisper as bw
    K, N = 64, 64
    w = torch.zeros((K, N), dtype=torch.float32)
    w[0, :] = 128.0
    w[1, :] = 0.5
    w[2, :] = -0.5
    w[3, :] = 2.5
    w[4, :] = -128.0
    raw, _, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(w.cuda(), torch.int8)
    raw = raw.cpu()
    assert float(scales[0]) == 1.0
    assert [int(raw[i, 0]) for i in range(5)] == [127, 1, -1, 3, -128]


def test_shape_errors():
    import b200_whisper as bw
    with pytest.raises(RuntimeError):
        bw.ops.symmetric_quantize_last_axis_of_batched_matrix(torch.zeros((64, 32), dtype=torch.float16).cuda())
    with pytest.raises(RuntimeError):
        bw.ops.symmetric_quantize_last_axis_of_batched_matrix(torch.zeros((24, 64), dtype=torch.float16).cuda())
    with pytest.raises(ValueError):
        bw.ops.symmetric_quantize_last_axis_of_batched_matrix(torch.zeros((0, 64), dtype=torch.float16).cuda())
