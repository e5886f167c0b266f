"""The persistent large-M weight-only GEMM (csrc/woq_gemm_tc.cu, woq_gemm_large_kernel: int8 weights expanded once per call,
fp16 x fp16 SS-UMMA pipeline, two TMEM accumulators so the epilogue of a tile overlaps the main loop of the next) -- the
path b200_woq_int8_gemm* takes from 4096 rows up (encoder: M = 1500 x batch).  Reference behaviour: the fpA_intB GEMM of
T/cpp/tensorrt_llm/kernels/cutlass_kernels/fpA_intB_gemm/fpA_intB_gemm_template.h:358-435 behind
WeightOnlyQuantMatmulPlugin::enqueue (weightOnlyQuantMatmulPlugin.cpp:162-222) and the separate fp16 bias / GELU /
residual layers of quantization/layer.py:311-312.

Checked against (1) the SAME operator evaluated in row blocks below the threshold (the per-tile kernel with its in-kernel
dequantisation): same arithmetic, so the results must agree to the bit; (2) the reference tests' column tolerance and the
tighter |A| @ |W| bound against an fp64 evaluation; (3) ragged last m-tile, in-place residual, determinism."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def gen(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g, dtype=torch.float32).half() * 2 - 1.0


def _fused(lib, x, proc, scales, n, bias, act, residual, out, ws):
    from b200_whisper import _lib
    m, k = x.shape
    _lib.check(lib.b200_woq_int8_gemm_fused(x.data_ptr(), m, k, proc.data_ptr(), scales.data_ptr(), n,
                                            bias.data_ptr() if bias is not None else None, act,
                                            residual.data_ptr() if residual is not None else None, out.data_ptr(),
                                            ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "fused gemm")


# (m, k, n): the four Whisper large-v2 shapes at encoder-like row counts; ragged last m-tiles (m % 128 != 0)
@pytest.mark.parametrize("m,k,n,act,use_res", [(4100, 1280, 1280, "none", True), (6000, 1280, 3840, "none", False),
                                               (4500, 1280, 5120, "gelu", False), (4224, 5120, 1280, "none", True),
                                               (24000, 1280, 1280, "none", True), (4097, 384, 1536, "gelu", True)])
def test_large_m_path_equals_the_tiled_kernel_and_the_reference_tolerance(m, k, n, act, use_res):
    import b200_whisper as bw
    from b200_whisper import _lib
    lib = _lib.load()
    x = (gen((m, k), 1) * 3).cuda()
    weight = gen((k, n), 2) * 0.05
    bias = gen((n,), 3).cuda()
    resid = gen((m, n), 4).cuda() if use_res else None
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(weight.cuda(), torch.int8)
    A = {"none": _lib.ACT_NONE, "gelu": _lib.ACT_GELU_ERF}[act]
    need = lib.b200_woq_workspace_bytes(m, n, k)
    assert need >= n * k * 2, "the workspace query must cover the expanded fp16 weights of the large-M path"
    ws = torch.empty((need,), dtype=torch.uint8, device="cuda")
    out = torch.full((m, n), float("nan"), dtype=torch.float16, device="cuda")
    n0 = lib.b200_launch_count()
    _fused(lib, x, proc, scales, n, bias, A, resid, out, ws)
    torch.cuda.synchronize()
    assert lib.b200_launch_count() - n0 == 2, "large-M path = weight expansion + one persistent GEMM launch"
    assert torch.isfinite(out.float()).all()

    # (1) the same operator in row blocks below the threshold
    ref = torch.empty_like(out)
    blk = 2048
    for r0 in range(0, m, blk):
        r1 = min(m, r0 + blk)
        _fused(lib, x[r0:r1], proc, scales, n, bias, A, resid[r0:r1] if use_res else None, ref[r0:r1], ws)
    torch.cuda.synchronize()
    if act == "none" and k <= 1280:
        assert torch.equal(out, ref), f"max diff {(out.float() - ref.float()).abs().max().item()}"
    elif act == "none":
        # deep K: the per-tile kernel splits K over CTAs at 2048 rows (another summation order): a couple of fp16 ulps apart at most (product and residual add both round)
        d = (out.float() - ref.float()).abs()
        assert (d <= ref.float().abs() * 2.0 ** -9 + 2e-3).all(), f"max diff {d.max().item()}"
    else:
        # the two epilogues share finish_output_tile: same bits expected here too, but GELU may amplify a last-bit difference
        assert (out.float() - ref.float()).abs().max().item() <= 1e-3

    # (2) fp64 evaluation on the dequantized weights, |A| @ |W| error bound (plain product only)
    if act == "none":
        w16 = (raw.to(torch.float16) * scales[None, :].to(torch.float16)).double()
        rows = torch.arange(0, m, max(1, m // 512), device="cuda")   # a sample of rows keeps the fp64 product small
        ideal = x[rows].double() @ w16 + bias.double()[None, :]
        if use_res:
            ideal = ideal + resid[rows].double()
        bound = 2e-3 * (x[rows].abs().double() @ w16.abs()) + ideal.abs() * 2.0 ** -9 + 2e-3
        err = (out[rows].double() - ideal).abs()
        assert (err / bound).max().item() <= 1.0

    # (3) determinism
    again = torch.empty_like(out)
    _fused(lib, x, proc, scales, n, bias, A, resid, again, ws)
    torch.cuda.synchronize()
    assert torch.equal(out, again)


def test_large_m_residual_in_place():
    """x += Linear(h) with the output buffer as the residual, as the encoder runtime updates its residual stream."""
    import b200_whisper as bw
    from b200_whisper import _lib
    lib = _lib.load()
    m, k, n = 5000, 1280, 1280
    h = gen((m, k), 11).cuda()
    weight = gen((k, n), 12) * 0.05
    bias = gen((n,), 13).cuda()
    x0 = gen((m, n), 14).cuda()
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(weight.cuda(), torch.int8)
    ws = torch.empty((lib.b200_woq_workspace_bytes(m, n, k),), dtype=torch.uint8, device="cuda")
    separate = torch.empty_like(x0)
    _fused(lib, h, proc, scales, n, bias, _lib.ACT_NONE, x0, separate, ws)
    x = x0.clone()
    _fused(lib, h, proc, scales, n, bias, _lib.ACT_NONE, x, x, ws)
    torch.cuda.synchronize()
    assert torch.equal(x, separate)


def test_small_workspace_falls_back_to_the_tiled_kernel():
    """A caller that sized its workspace for the per-tile kernel only still gets the right answer (no expansion buffer:
    the per-tile kernel runs)."""
    import b200_whisper as bw
    from b200_whisper import _lib
    lib = _lib.load()
    m, k, n = 4200, 5120, 1280
    x = gen((m, k), 21).cuda()
    weight = gen((k, n), 22) * 0.05
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(weight.cuda(), torch.int8)
    big = torch.empty((lib.b200_woq_workspace_bytes(m, n, k),), dtype=torch.uint8, device="cuda")
    small = torch.empty((n * k * 2 - 4096,), dtype=torch.uint8, device="cuda")
    a, b = (torch.empty((m, n), dtype=torch.float16, device="cuda") for _ in range(2))
    _fused(lib, x, proc, scales, n, None, _lib.ACT_NONE, None, a, big)
    n0 = lib.b200_launch_count()
    _fused(lib, x, proc, scales, n, None, _lib.ACT_NONE, None, b, small)
    torch.cuda.synchronize()
    assert lib.b200_launch_count() - n0 == 1
    assert torch.equal(a, b)


def test_misaligned_workspace_falls_back_to_the_tiled_kernel():
    """The expanded weights are a TMA source (128-byte aligned box rows): a workspace pointer that is not keeps the per-tile
    kernel instead of failing."""
    import b200_whisper as bw
    from b200_whisper import _lib
    lib = _lib.load()
    m, k, n = 4200, 1280, 1280
    x = gen((m, k), 31).cuda()
    weight = gen((k, n), 32) * 0.05
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(weight.cuda(), torch.int8)
    big = torch.empty((lib.b200_woq_workspace_bytes(m, n, k) + 256,), dtype=torch.uint8, device="cuda")
    a, b = (torch.empty((m, n), dtype=torch.float16, device="cuda") for _ in range(2))
    _fused(lib, x, proc, scales, n, None, _lib.ACT_NONE, None, a, big)
    n0 = lib.b200_launch_count()
    _fused(lib, x, proc, scales, n, None, _lib.ACT_NONE, None, b, big[16:])
    torch.cuda.synchronize()
    assert lib.b200_launch_count() - n0 == 1
    assert torch.equal(a, b)
