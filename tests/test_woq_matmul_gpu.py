"""GPU parity of weight_only_quant_matmul (SIMT GEMV and tcgen05 GEMM) through the C ABI.

Procedures follow the reference's own tests, T/tests/quantization/test_weight_only_quant_matmul.py:84-130 and
_utils.py:15-88: seed 0, activations U(-1,1)*200, weights U(-1,1); reference = fp32 x @ (int8 * scale); tolerance per
column atol = 1.5 * max(col) / 128.  On top of that a much tighter check against the oracle's "identically dequantized
weights" arithmetic: |err| <= 2e-3 * (|A| @ |W16|) + fp16 rounding of the result.
"""
import numpy as np
import pytest
import torch

from oracle import woq

pytestmark = pytest.mark.gpu

POLICIES = {"auto": 0, "simt": 1, "tc": 2}


def gen(shape, seed=0):
    torch.manual_seed(seed)
    return torch.rand(shape, dtype=torch.float16) * 2 - 1.0


def run_plugin(mat1, proc, scales, policy, **kw):
    import b200_whisper as bw
    from b200_whisper import _lib
    from b200_whisper.quantization.functional import weight_only_quant_matmul
    lib = _lib.load()
    _lib.check(lib.b200_woq_set_kernel_policy(POLICIES[policy]))
    try:
        # the reference passes the int8 bytes viewed as float32 [K, N/4]
        out = weight_only_quant_matmul(mat1.cuda(), proc.cuda().view(torch.float32), scales.cuda(), 1, **kw)
        torch.cuda.synchronize()
    finally:
        lib.b200_woq_set_kernel_policy(0)
    return out.cpu()


def colwise_near(ref, act):
    # T/tests/quantization/_utils.py:66-88
    ref = ref.float().numpy()
    act = act.float().numpy()
    if ref.shape[0] > 1:
        atol = np.maximum(ref.max(axis=0), 0) * (1.0 / 128) * 1.5
        bad = np.abs(ref - act) > atol[None, :] + 1e-7 * np.abs(act)
        assert not bad.any(), f"{bad.sum()} elements outside the reference tolerance"
    else:
        atol = ref.max() * (1.0 / 128) * 1.5
        np.testing.assert_allclose(ref, act, atol=atol)


def tight_check(mat1, raw, scales, out, what):
    a = mat1.float()
    w16 = (raw.to(torch.float16) * scales[None, :].to(torch.float16)).float()  # fp16(q * s16)
    ideal = a.double() @ w16.double()
    bound = 2e-3 * (a.abs().double() @ w16.abs().double()) + ideal.abs() * 2.0 ** -10 + 1e-6
    err = (out.double() - ideal).abs()
    worst = (err / bound).max().item()
    assert worst <= 1.0, f"{what}: worst err/bound = {worst:.3f} (max abs err {err.max().item():.4g})"


@pytest.mark.parametrize("policy", ["simt", "tc"])
@pytest.mark.parametrize("n,k", [(1024, 4096), (4096, 512)])
def test_conversion(policy, n, k):
    """Identity activation 'un-converts' the preprocessed weights: pins layout + dequant end to end."""
    import b200_whisper as bw
    weight = gen((k, n))
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(weight, torch.int8)
    eye = torch.eye(k, dtype=torch.float16)
    if policy == "simt":
        eye = eye[:64]  # the SIMT path is for small M; 64 rows of the identity are enough to pin the layout
    act = run_plugin(eye, proc, scales, policy)
    expect = (raw.to(torch.float16) * scales[None, :]).to(torch.float16)[: eye.shape[0]]
    assert torch.equal(act, expect), f"max diff {(act.float() - expect.float()).abs().max()}"
    colwise_near(weight[: eye.shape[0]], act)


@pytest.mark.parametrize("policy,m,n,k", [("simt", 1, 1024, 4096), ("tc", 1, 1024, 4096), ("tc", 128, 6144, 12288),
                                          ("auto", 1, 1024, 4096), ("auto", 128, 6144, 12288)])
def test_matmul_reference_shapes(policy, m, n, k):
    import b200_whisper as bw
    mat1 = gen((m, k)) * 200.0
    weight = gen((k, n))
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(weight, torch.int8)
    out = run_plugin(mat1, proc, scales, policy)
    ref = (mat1.float() @ raw.float()) * scales.float()[None, :]  # woq_gt_matmul, _utils.py:37-63
    colwise_near(ref.to(torch.float16), out)
    tight_check(mat1, raw, scales, out, f"{policy} m={m}")


WHISPER_SHAPES = [(1280, 3840), (1280, 1280), (1280, 5120), (5120, 1280), (384, 1152), (1536, 384)]


@pytest.mark.parametrize("k,n", WHISPER_SHAPES)
@pytest.mark.parametrize("m", [1, 2, 3, 4])
def test_simt_whisper_shapes(k, n, m):
    import b200_whisper as bw
    mat1 = gen((m, k), seed=1)
    weight = gen((k, n), seed=2) * 0.05
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(weight, torch.int8)
    out = run_plugin(mat1, proc, scales, "simt")
    tight_check(mat1, raw, scales, out, f"simt m={m} k={k} n={n}")
    if m == 1 and k == 1280 and n == 1280:
        # the oracle's C restatement of the two reference accumulation orders brackets the result too
        o_gemv = woq.woq_matmul(mat1.numpy(), raw.numpy(), scales.numpy(), "gemv").astype(np.float32)
        assert np.abs(out.float().numpy() - o_gemv).max() <= 0.02 * np.abs(o_gemv).max()


@pytest.mark.parametrize("k,n", WHISPER_SHAPES + [(1280, 51904)])
@pytest.mark.parametrize("m", [1, 5, 16, 17, 64, 100, 256, 300])
def test_tc_whisper_shapes(k, n, m):
    import b200_whisper as bw
    if n > 10000 and m not in (1, 16, 256):
        pytest.skip("logits-sized N only at a few M")
    mat1 = gen((m, k), seed=3)
    weight = gen((k, n), seed=4) * 0.05
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(weight.cuda(), torch.int8)
    out = run_plugin(mat1, proc, scales, "tc")
    tight_check(mat1, raw.cpu(), scales.cpu(), out, f"tc m={m} k={k} n={n}")


@pytest.mark.parametrize("policy,m", [("simt", 1), ("simt", 4), ("tc", 16), ("tc", 48), ("tc", 100), ("tc", 300),
                                      ("tc", 1500)])
def test_fused_epilogue(policy, m):
    """bias, GELU and residual fused in the epilogue == the reference's separate fp16 layers (layer.py:311-312)."""
    import b200_whisper as bw
    k, n = 1280, 5120
    mat1 = gen((m, k), seed=5)
    weight = gen((k, n), seed=6) * 0.05
    bias = gen((n,), seed=7)
    resid = gen((m, n), seed=8)
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(weight, torch.int8)
    plain = run_plugin(mat1, proc, scales, policy)
    fused = run_plugin(mat1, proc, scales, policy, bias=bias.cuda(), activation="gelu", residual=resid.cuda())
    step = (plain.float() + bias.float()[None, :]).half()
    step = torch.nn.functional.gelu(step.float()).half()
    step = (step.float() + resid.float()).half()
    assert (fused.float() - step.float()).abs().max() <= 2e-3 * step.float().abs().max() + 1e-3
    # deterministic: same inputs, same bits
    again = run_plugin(mat1, proc, scales, policy, bias=bias.cuda(), activation="gelu", residual=resid.cuda())
    assert torch.equal(fused, again)


@pytest.mark.parametrize("m", [16, 200, 1500])
def test_residual_in_place(m):
    """x += Linear(h): the output buffer IS the residual (how the encoder and decoder runtimes update the residual stream).
    Large tiles fetch a batch of residual rows before they store any, so the in-place form must still read every element
    before it is overwritten."""
    import b200_whisper as bw
    from b200_whisper import _lib
    lib = _lib.load()
    k, n = 1280, 1280
    h = gen((m, k), seed=31).cuda()
    weight = gen((k, n), seed=32) * 0.05
    bias = gen((n,), seed=33).cuda()
    x0 = gen((m, n), seed=34).cuda()
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(weight.cuda(), torch.int8)
    ws = torch.empty((max(lib.b200_woq_workspace_bytes(m, n, k), 1 << 20),), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run(residual, out):
        _lib.check(lib.b200_woq_int8_gemm_fused(h.data_ptr(), m, k, proc.data_ptr(), scales.data_ptr(), n, bias.data_ptr(),
                                                _lib.ACT_NONE, residual.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                                ws.numel(), st), "fused gemm")
        torch.cuda.synchronize()

    separate = torch.empty_like(x0)
    run(x0, separate)
    x = x0.clone()
    run(x, x)
    assert torch.equal(x, separate)
    plain = run_plugin(h.cpu(), proc.cpu(), scales.cpu(), "tc", bias=bias)
    want = (plain.float() + x0.cpu().float()).half()
    assert (x.cpu().float() - want.float()).abs().max().item() <= 2e-3 * want.float().abs().max().item() + 1e-3


def test_linearity_property_large():
    """Size-independent property at a full large-v2 shape: f(a + b) ~= f(a) + f(b) and f(2a) == 2 f(a) exactly."""
    import b200_whisper as bw
    k, n, m = 5120, 1280, 16
    a = gen((m, k), seed=9)
    weight = gen((k, n), seed=10) * 0.05
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(weight, torch.int8)
    fa = run_plugin(a, proc, scales, "tc")
    f2a = run_plugin(a * 2, proc, scales, "tc")
    assert torch.equal(f2a, fa * 2)  # scaling by 2 is exact in fp16/fp32 arithmetic


def test_errors_and_empty():
    from b200_whisper import _lib
    lib = _lib.load()
    x = torch.zeros((1, 96), dtype=torch.float16, device="cuda")
    w = torch.zeros((96, 64), dtype=torch.int8, device="cuda")
    s = torch.ones((64,), dtype=torch.float16, device="cuda")
    o = torch.zeros((1, 64), dtype=torch.float16, device="cuda")
    rc = lib.b200_woq_int8_gemm(x.data_ptr(), 1, 96, w.data_ptr(), s.data_ptr(), 64, o.data_ptr(), None, 0, None)
    assert rc == 1 and b"multiple of 64" in lib.b200_last_error()
    rc = lib.b200_woq_int8_gemm(x.data_ptr(), 0, 64, w.data_ptr(), s.data_ptr(), 64, o.data_ptr(), None, 0, None)
    assert rc == 0  # empty batch is a no-op
    rc = lib.b200_woq_int8_gemm(None, 1, 64, w.data_ptr(), s.data_ptr(), 64, o.data_ptr(), None, 0, None)
    assert rc == 1


@pytest.mark.parametrize("policy,m,k,n", [("tc", 16, 1280, 3840), ("tc", 16, 1280, 5120), ("tc", 5, 384, 1152),
                                          ("tc", 48, 1280, 1280), ("tc", 100, 1280, 1280), ("tc", 300, 1280, 1280),
                                          ("simt", 3, 1280, 1280)])
def test_layernorm_fused_gemm(policy, m, k, n):
    """b200_woq_int8_gemm_ln_fused == LayerNorm kernel followed by the plain GEMM (same fp16 rounding of LN(x))."""
    import b200_whisper as bw
    from b200_whisper import _lib
    from b200_whisper.functional import layer_norm
    lib = _lib.load()
    torch.manual_seed(m + n)
    x = (torch.randn((m, k)) * 2 + 0.3).half().cuda()
    gamma = (1 + 0.1 * torch.randn(k)).half().cuda()
    beta = (0.1 * torch.randn(k)).half().cuda()
    weight = gen((k, n), seed=11) * 0.05
    bias = gen((n,), seed=12).cuda()
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(weight.cuda(), torch.int8)
    h = layer_norm(x, (k,), gamma, beta, 1e-5)
    ref = run_plugin(h.cpu(), proc.cpu(), scales.cpu(), policy, bias=bias, activation="gelu")
    out = torch.empty((m, n), dtype=torch.float16, device="cuda")
    ws = torch.empty((lib.b200_woq_workspace_bytes(m, n, k) + m * k * 2,), dtype=torch.uint8, device="cuda")
    _lib.check(lib.b200_woq_set_kernel_policy(POLICIES[policy]))
    try:
        rc = lib.b200_woq_int8_gemm_ln_fused(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-5, m, k, proc.data_ptr(),
                                             scales.data_ptr(), n, bias.data_ptr(), _lib.ACT_GELU_ERF, None, out.data_ptr(),
                                             ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "ln fused gemm")
        torch.cuda.synchronize()
    finally:
        lib.b200_woq_set_kernel_policy(0)
    # LN(x) is rounded to fp16 the same way in both paths; accumulation order is identical -> tiny differences only
    assert (out.cpu().float() - ref.float()).abs().max().item() <= 2e-3 * ref.float().abs().max().item() + 1e-3


@pytest.mark.parametrize("m,k,n,res", [(16, 1280, 3840, False), (16, 1280, 5120, False), (16, 1280, 1280, True),
                                       (5, 384, 1152, False), (32, 1280, 1280, False), (9, 1536, 256, True), (9, 2048, 256, True),
                                       (48, 1280, 1280, False), (3, 1280, 1280, False)])
def test_layernorm_folded_gemm(m, k, n, res):
    """b200_woq_int8_gemm_ln_folded (LayerNorm folded into the tcgen05 kernel: gamma in the dequant, statistics in the
    epilogue) against LayerNorm -> reference-order GEMM.  x has a large per-row offset so a mean term that does not
    cancel would show.  M = 48 and M = 3 take the two-launch fallback of the same entry point."""
    import b200_whisper as bw
    from b200_whisper import _lib
    from b200_whisper.functional import layer_norm
    lib = _lib.load()
    torch.manual_seed(m + n + k)
    x = (torch.randn((m, k)) * 2 + 3.0 * torch.randn((m, 1))).half().cuda()
    gamma = (1 + 0.2 * torch.randn(k)).half().cuda()
    beta = (0.2 * torch.randn(k)).half().cuda()
    weight = gen((k, n), seed=21) * 0.05
    bias = gen((n,), seed=22).cuda()
    residual = gen((m, n), seed=23).cuda() if res else None
    raw, proc, scales = bw.ops._symmetric_quantize_last_axis_of_batched_matrix(weight.cuda(), torch.int8)
    c1s = torch.empty((n,), dtype=torch.float32, device="cuda")
    c2 = torch.empty((n,), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.b200_woq_ln_fold_prepare(proc.data_ptr(), scales.data_ptr(), gamma.data_ptr(), beta.data_ptr(), k, n,
                                            c1s.data_ptr(), c2.data_ptr(), st), "ln fold prepare")
    # the prepared vectors against their definition, from the raw (unprocessed) int8 weights
    wg = (raw.float() * gamma.float()[:, None]).half().float()
    exp_c1 = wg.sum(0) * scales.float()
    exp_c2 = (raw.float() * beta.float()[:, None]).sum(0) * scales.float()
    assert (c1s - exp_c1).abs().max().item() <= 1e-4 * exp_c1.abs().max().item() + 1e-5
    assert (c2 - exp_c2).abs().max().item() <= 1e-4 * exp_c2.abs().max().item() + 1e-5
    h = layer_norm(x, (k,), gamma, beta, 1e-5)
    ref = run_plugin(h.cpu(), proc.cpu(), scales.cpu(), "tc", bias=bias)
    if res:
        ref = (ref.float() + residual.cpu().float()).half()
    out = torch.empty((m, n), dtype=torch.float16, device="cuda")
    ws = torch.empty((lib.b200_woq_workspace_bytes(m, n, k) + m * k * 2,), dtype=torch.uint8, device="cuda")
    rc = lib.b200_woq_int8_gemm_ln_folded(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), c1s.data_ptr(), c2.data_ptr(),
                                          1e-5, m, k, proc.data_ptr(), scales.data_ptr(), n, bias.data_ptr(),
                                          _lib.ACT_NONE, residual.data_ptr() if res else None, out.data_ptr(),
                                          ws.data_ptr(), ws.numel(), st)
    _lib.check(rc, "ln folded gemm")
    torch.cuda.synchronize()
    # tolerance: fp16 rounding of LN(x) (reference order) vs fp16 rounding of Wint*gamma (folded order)
    assert (out.cpu().float() - ref.float()).abs().max().item() <= 3e-3 * ref.float().abs().max().item() + 2e-3


@pytest.mark.parametrize("mode", ["cluster", "global"])
def test_splitk_modes_agree(mode, monkeypatch):
    """Cluster (DSMEM) and global-slab split-K reductions are both deterministic and agree to fp32 rounding."""
    import subprocess
    import sys
    code = r"""
import torch, sys
sys.path.insert(0, '.')
import b200_whisper as bw
from b200_whisper.quantization.functional import weight_only_quant_matmul
torch.manual_seed(0)
k, n, m = 5120, 1280, 16
a = (torch.rand((m, k)) * 2 - 1).half().cuda()
w = ((torch.rand((k, n)) * 2 - 1) * 0.05).half().cuda()
proc, scales = bw.ops.symmetric_quantize_last_axis_of_batched_matrix(w, torch.int8)
o1 = weight_only_quant_matmul(a, proc, scales, 1).clone()
o2 = weight_only_quant_matmul(a, proc, scales, 1).clone()
torch.cuda.synchronize()
assert torch.equal(o1, o2)
w16 = None
print(float(o1.float().abs().sum()))
"""
    import os
    env = dict(os.environ, B200_SPLITK=mode)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=os.path.dirname(os.path.dirname(__file__)))
    assert r.returncode == 0, r.stderr[-2000:]
    test_splitk_modes_agree.results = getattr(test_splitk_modes_agree, "results", {})
    test_splitk_modes_agree.results[mode] = float(r.stdout.strip().splitlines()[-1])
    if len(test_splitk_modes_agree.results) == 2:
        a, b = test_splitk_modes_agree.results.values()
        assert abs(a - b) <= 1e-3 * abs(a)
