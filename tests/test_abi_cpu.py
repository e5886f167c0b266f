"""CPU tests: the C-ABI library builds for sm_100a, loads, exports every symbol include/b200_whisper.h declares, and
fails loudly (no fallback) without a GPU.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__
    so = os.path.join(ROOT, "eddie-wang-hackathon2023_b200", "lib", "libb200_whisper.so")
    if not os.path.exists(so):
        __graft_entry__.build()
    import b200_whisper
    return b200_whisper.load()


def header_functions():
    text = open(os.path.join(ROOT, "include", "b200_whisper.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200_whisper.h but not exported"
    from b200_whisper import _lib
    assert set(_lib.declared_symbols()) <= set(names) | {"b200_init"}


def test_sass_is_blackwell_native():
    so = os.path.join(ROOT, "eddie-wang-hackathon2023_b200", "lib", "libb200_whisper.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS"


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a GPU-less machine")
def test_no_cpu_fallback(lib):
    import b200_whisper as bw
    with pytest.raises(RuntimeError, match="no CPU fallback|no sm_100"):
        bw.ops.symmetric_quantize_last_axis_of_batched_matrix(torch.zeros((64, 64), dtype=torch.float16))
    x = torch.zeros((1, 64), dtype=torch.float16)
    rc = lib.b200_woq_int8_gemm(x.data_ptr(), 1, 64, x.data_ptr(), x.data_ptr(), 64, x.data_ptr(), None, 0, None)
    assert rc == 3 and b"no CPU fallback" in lib.b200_last_error()


def test_argument_validation_without_gpu(lib):
    # argument errors are reported before any device work
    rc = lib.b200_woq_int8_gemm(None, 1, 64, None, None, 64, None, None, 0, None)
    assert rc == 1
    assert lib.b200_woq_workspace_bytes(16, 1280, 1280) > 0
    assert lib.b200_cross_attention_workspace_bytes(1, 4, 64, 1500) > 0  # a handful of (row, head) pairs: split across CTAs
    # the entry points added around the three operators validate their arguments the same way (1 = invalid argument,
    # 2 = unsupported configuration), without a device
    assert lib.b200_woq_ln_fold_prepare(None, None, None, None, 1280, 1280, None, None, None) == 1
    assert lib.b200_woq_int8_gemm_ln_folded(None, None, None, None, None, 1e-5, 16, 1280, None, None, 1280, None, 0,
                                            None, None, None, 0, None) == 1
    assert lib.b200_conv1d_workspace_bytes(16, 80, 1280, 3000, 3) == ((3 * 1280 * 128 * 2 + 1023) // 1024) * 1024 \
        + 16 * 3000 * 128 * 2
    assert lib.b200_conv1d_fp16_tc(None, None, None, None, 1, 80, 1280, 3000, 3, 1, 1, 0, None, 0, None) == 1
    assert lib.b200_attention_bidirectional_fp16(None, None, 1, 1500, 20, 64, None) == 1
    assert lib.b200_whisper_filtered_argmax(None, 1, 51865, None, 50257, 50363, 50364, 220, -1, None, None, None, None) == 1
    assert lib.b200_transpose_add_pos_fp16(None, None, None, 1, 1280, 1500, None) == 1
    # paged KV cache entry points (kvCacheUtils.h:34-112): the block geometry is validated before any device work
    from b200_whisper._lib import MmhaParams
    import ctypes as ct
    mp = MmhaParams()
    t8 = torch.zeros(8)
    mp.qkv = mp.out = t8.data_ptr()
    mp.batch_size, mp.num_heads, mp.head_size, mp.max_seq_len, mp.past_kv_length, mp.q_scaling = 1, 4, 64, 448, 3, 1.0
    assert lib.b200_mmha_generation_paged(ct.byref(mp), None, 7, 64, None) == 1                      # no table
    assert lib.b200_mmha_generation_paged(ct.byref(mp), t8.data_ptr(), 7, 48, None) == 1            # not a power of 2
    assert b"power of 2" in lib.b200_last_error()
    assert lib.b200_mmha_generation_paged(ct.byref(mp), t8.data_ptr(), 6, 64, None) == 1            # 384 < 448 tokens
    assert lib.b200_attention_context_paged(t8.data_ptr(), None, t8.data_ptr(), t8.data_ptr(), 7, 64, None, 1, 4, 4, 32, 448,
                                            0, 1.0, None) == 2                                      # head size 32
    x = torch.zeros(8)
    assert lib.b200_logits_range_softmax(None, 1, 51865, 50259, 50358, 50362, None, None, None, None) == 1
    assert lib.b200_logits_range_softmax(x.data_ptr(), 1, 8, 5, 3, 0, x.data_ptr(), None, None, None) == 1   # empty range
    assert lib.b200_logits_range_softmax(x.data_ptr(), 1, 8, 2, 6, 9, None, None, x.data_ptr(), None) == 1   # probe outside
    assert lib.b200_logits_range_softmax(x.data_ptr(), 1, 8, 2, 6, 0, None, None, None, None) == 1           # no output


def test_quant_mode_flags():
    # T/tests/quantization/test_mode.py
    from b200_whisper import QuantMode
    qm = QuantMode.ACTIVATIONS | QuantMode.INT8_WEIGHTS
    assert qm._all(QuantMode.ACTIVATIONS | QuantMode.INT8_WEIGHTS) and not qm._all(QuantMode.ACTIVATIONS)
    assert qm._all(QuantMode.ACTIVATIONS, mask=QuantMode.ACTIVATIONS)
    assert qm._any(QuantMode.ACTIVATIONS) and not qm._any(QuantMode.PER_TOKEN)
    assert QuantMode.COUNT.value == 1 << 7
    assert QuantMode.from_description(True, False, False, False) == QuantMode.INT8_WEIGHTS
    assert QuantMode.use_weight_only() == QuantMode.INT8_WEIGHTS
    assert QuantMode.use_weight_only(True) == QuantMode.INT4_WEIGHTS
    assert QuantMode.from_description(True, True, False, False) == QuantMode.ACTIVATIONS | QuantMode.INT8_WEIGHTS
    assert QuantMode.use_smooth_quant(True, True) == (QuantMode.ACTIVATIONS | QuantMode.INT8_WEIGHTS
                                                      | QuantMode.PER_TOKEN | QuantMode.PER_CHANNEL)
    with pytest.raises(ValueError):
        QuantMode.from_description(False, True)
    with pytest.raises(ValueError):
        QuantMode.from_description(True, False, per_token=True)
    m = QuantMode.use_weight_only().set_int8_kv_cache()
    assert m.is_int8_weight_only() and m.is_weight_only() and m.has_int8_kv_cache() and not m.has_fp8_kv_cache()
    assert m.has_any_quant() and not QuantMode(0).has_any_quant()


def test_tc_launch_plan_for_the_decoder_shapes(lib):
    """The split-K planner (host logic, 148 SMs assumed without a GPU): the decode shapes of large-v2 get cluster splits
    that keep each CTA's k range inside the weight ring, 256-row tiles are not split (and only used when they fill the
    GPU), the M = 128 cliff stays fixed."""
    import ctypes
    if os.environ.get("B200_SPLITK", "cluster") != "cluster":
        pytest.skip("planner defaults are tested in cluster mode")

    def plan(m, n, k):
        out = (ctypes.c_int * 5)()
        assert lib.b200_debug_woq_plan(m, n, k, out) == 0
        return tuple(out)

    # (MT, m_tiles, n_tiles, splits, cluster)
    assert plan(16, 3840, 1280) == (16, 1, 30, 4, 1)
    assert plan(16, 1280, 1280) == (16, 1, 10, 8, 1)
    assert plan(16, 5120, 1280) == (16, 1, 40, 4, 1)   # 160 CTAs: two per SM on a few SMs
    assert plan(16, 1280, 5120) == (16, 1, 10, 8, 1)   # 10 k-blocks per CTA = the whole weight ring
    assert plan(1, 3840, 1280)[0] == 16 and plan(1, 3840, 1280)[4] == 1
    # 128-row tiles split K only when K is deep (round 2: the 64 KB DSMEM exchange of a split cost more than it saved at
    # K = 1280 -- 1280 -> 5120 at M = 128 was slower than at M = 256, profiles/r02_gemm_sweep_midM.txt)
    assert plan(128, 3840, 1280) == (128, 1, 30, 1, 0) and plan(128, 5120, 1280) == (128, 1, 40, 1, 0)
    mt, mtiles, ntiles, splits, cluster = plan(128, 1280, 5120)
    assert mt == 128 and cluster == 1 and 2 <= splits <= 8
    # fewer 256-row tiles than SMs: 128-row tiles instead (twice the CTAs, cluster split-K still available for deep K)
    assert plan(256, 3840, 1280) == (128, 2, 30, 1, 0) and plan(256, 1280, 5120) == (128, 2, 10, 4, 1)
    assert plan(1500, 1280, 5120) == (128, 12, 10, 1, 0) and plan(1500, 3840, 1280)[:4] == (256, 6, 30, 1)
    assert plan(24000, 3840, 1280)[:4] == (256, 94, 30, 1)
    for m, n, k in [(16, 3840, 1280), (16, 1280, 5120), (32, 1280, 1280), (32, 3840, 1280), (4, 1280, 1280)]:
        mt, _, nt, s, c = plan(m, n, k)
        assert c == 1 and (k // 64 + s - 1) // s <= {16: 10, 32: 7}[mt]  # decode tiles: k range resident in the ring
