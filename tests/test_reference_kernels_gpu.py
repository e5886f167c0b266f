"""Second pin: the REFERENCE's own CUDA kernels, compiled for sm_100a from /root/reference into
oracle/_ref/libref_gpu.so (oracle/Makefile `refgpu`, oracle/ref_gpu_harness.cu), run on the same B200:

  * weight_only_gemv_launcher (T/cpp/tensorrt_llm/kernels/weightOnlyMatrixVectorMultiplication.cu:136-205,371-378)
      - must agree BIT-EXACTLY with oracle/woq_oracle.c's "gemv_exact" restatement (this pins the oracle),
      - and bounds this repo's GEMV / tcgen05 GEMM at M = 1 on the reference's preprocessed layout;
  * masked_multihead_attention_kernel, Dh = 64, fp16, int8 / fp16 KV cache
    (T/cpp/tensorrt_llm/kernels/decoderMaskedMultiheadAttention/decoderMaskedMultiheadAttentionTemplate.h:1195-2017)
      - cache append bytes identical, attention output within the reference test's own 2e-3 tolerance
        (T/tests/attention/test_gpt_attention.py:828-831).

The library travels to the GPU box as a built artefact; the tests skip when it was not built."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import woq

pytestmark = pytest.mark.gpu


def _ref():
    lib = woq.ref_gpu_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libref_gpu.so not built (make -C oracle refgpu needs /root/reference)")
    return lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("k,n", [(1280, 1280), (1280, 3840), (5120, 1280), (4096, 1024)])
def test_reference_gemv_pins_oracle_and_kernels(k, n):
    import b200_whisper as bw
    from b200_whisper import _lib
    ref = _ref()
    lib = _lib.load()
    torch.manual_seed(k + n)
    x = (torch.rand((1, k)) * 2 - 1).half()
    w = ((torch.rand((k, n)) * 2 - 1)).half()
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(w.cuda(), torch.int8)
    xg = x.cuda()
    out_ref = torch.empty((1, n), dtype=torch.float16, device="cuda")
    rc = ref.ref_gpu_gemv(xg.data_ptr(), proc.data_ptr(), scales.data_ptr(), None, out_ref.data_ptr(), k, n, _stream())
    torch.cuda.synchronize()
    assert rc == 0
    # (1) the oracle's restatement of the GEMV arithmetic is the kernel, bit for bit
    exp = woq.woq_matmul(x.numpy(), raw.cpu().numpy(), scales.cpu().numpy(), mode="gemv_exact")
    assert np.array_equal(out_ref.cpu().numpy().view(np.uint16), exp.view(np.uint16)), "oracle gemv != reference kernel"
    # (2) this repo's kernels at M = 1: SIMT GEMV and tcgen05 GEMM, against the reference kernel
    ws = torch.empty((lib.b200_woq_workspace_bytes(1, n, k),), dtype=torch.uint8, device="cuda")
    for policy in (1, 2):
        out = torch.empty((1, n), dtype=torch.float16, device="cuda")
        _lib.check(lib.b200_woq_set_kernel_policy(policy))
        try:
            _lib.check(lib.b200_woq_int8_gemm(xg.data_ptr(), 1, k, proc.data_ptr(), scales.data_ptr(), n, out.data_ptr(),
                                              ws.data_ptr(), ws.numel(), _stream()))
            torch.cuda.synchronize()
        finally:
            lib.b200_woq_set_kernel_policy(0)
        r, o = out_ref.float(), out.float()
        # the reference rounds every product to fp16 before accumulating; ours keeps more precision: bound the gap by
        # the reference tests' own column tolerance (T/tests/quantization/_utils.py:66-88)
        atol = r.abs().max().item() * (1.0 / 128) * 1.5
        assert (r - o).abs().max().item() <= atol
        assert (r - o).abs().mean().item() <= 2e-3 * r.abs().mean().item() + 1e-3


@pytest.mark.parametrize("int8", [True, False])
@pytest.mark.parametrize("B,past,masked", [(3, 5, False), (3, 37, False), (3, 200, False), (16, 37, False),
                                           (16, 447, False),  # headline batch, last slot of Smax = 448
                                           (3, 200, True), (16, 100, True)])  # padding keys via masked_tokens
def test_reference_mmha_matches(int8, B, past, masked):
    from b200_whisper import _lib
    ref = _ref()
    lib = _lib.load()
    torch.manual_seed(past + B)
    H, D, Smax = 20, 64, 448
    hidden = H * D
    dev = "cuda"
    qkv = torch.randn((B, 3 * hidden), device=dev).half()
    t = 4.0 / 127.0
    oq = torch.tensor([1.0 / t], dtype=torch.float32, device=dev)
    qo = torch.tensor([t], dtype=torch.float32, device=dev)
    if int8:
        cache0 = torch.randint(-127, 128, (B, 2, H, Smax, D), device=dev, dtype=torch.int8)
    else:
        cache0 = torch.randn((B, 2, H, Smax, D), device=dev).half()
    seq = torch.full((B,), past, dtype=torch.int32, device=dev)
    zeros = torch.zeros((B,), dtype=torch.int32, device=dev)
    mask = None
    if masked:
        # nonzero = key excluded (padding between prompt and generated tokens, Template.h:1678-1680,1730); never the
        # whole row, never the slot being written
        mask = (torch.rand((B, Smax), device=dev) < 0.25).to(torch.int32)
        mask[:, 0] = 0
        mask[:, past:] = 0
    mask_ptr = None if mask is None else mask.data_ptr()

    c_ref, o_ref = cache0.clone(), torch.empty((B, hidden), dtype=torch.float16, device=dev)
    rc = ref.ref_gpu_mmha(qkv.data_ptr(), o_ref.data_ptr(), c_ref.data_ptr(), seq.data_ptr(), mask_ptr, zeros.data_ptr(),
                          oq.data_ptr(), qo.data_ptr(), B, H, Smax, past, past, 1 if int8 else 0, 1.0, _stream())
    torch.cuda.synchronize()
    assert rc == 0

    c_our, o_our = cache0.clone(), torch.empty((B, hidden), dtype=torch.float16, device=dev)
    p = _lib.MmhaParams()
    p.qkv, p.qkv_bias, p.out = qkv.data_ptr(), None, o_our.data_ptr()
    p.kv_cache = c_our.data_ptr()
    p.sequence_lengths = seq.data_ptr()
    p.masked_tokens = mask_ptr
    p.kv_scale_orig_quant = oq.data_ptr()
    p.kv_scale_quant_orig = qo.data_ptr()
    p.batch_size, p.num_heads, p.head_size = B, H, D
    p.max_seq_len, p.past_kv_length, p.int8_kv_cache, p.q_scaling = Smax, past, 1 if int8 else 0, 1.0
    _lib.check(lib.b200_mmha_generation(ctypes.byref(p), _stream()), "mmha_generation")
    torch.cuda.synchronize()

    # the appended K/V row and every untouched row: identical bytes
    assert torch.equal(c_ref.view(torch.uint8), c_our.view(torch.uint8)), "KV cache contents differ from the reference kernel"
    err = (o_ref.float() - o_our.float()).abs().max().item()
    assert err <= 2e-3 * max(1.0, o_ref.float().abs().max().item()), f"MMHA output differs from the reference kernel: {err}"
