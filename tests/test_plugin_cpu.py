"""CPU tests of the TensorRT plugin CLASSES (no compute): registry identity, creator fields, serialization byte
layout, shape inference and format combinations -- the boundary contract of SURVEY.md 8b, checked against the
reference sources cited in the plugin headers."""
import struct

import numpy as np
import pytest

from b200_whisper.plugin import TrtPlugin, get_plugin_creator

HALF, FLOAT = 1, 0


def woq_fields(type_id=HALF, weight_type_id=1):
    # same list as quantization/functional.py:61-70 of the reference
    return [("type_id", np.array([type_id], np.int32)), ("weight_type_id", np.array(weight_type_id, dtype=np.int32))]


def attn_fields(**over):
    f = dict(num_heads=20, head_size=64, unidirectional=1, q_scaling=1.0, rotary_embedding_dim=0, neox_rotary_style=0,
             context_fmha_type=0, multi_block_mode=0, multi_query_mode=0, int8_kv_cache=1, fp8_kv_cache=0,
             remove_input_padding=0, mask_type=1, paged_kv_cache=0, type_id=HALF, in_flight_batching=0)
    f.update(over)
    dt = dict(q_scaling=np.float32, neox_rotary_style=np.int8, context_fmha_type=np.int8, multi_block_mode=np.int8,
              multi_query_mode=np.int8, remove_input_padding=np.int8)
    # order of functional.py:2925-2930
    order = ["num_heads", "head_size", "unidirectional", "q_scaling", "rotary_embedding_dim", "neox_rotary_style",
             "context_fmha_type", "multi_block_mode", "multi_query_mode", "int8_kv_cache", "fp8_kv_cache",
             "remove_input_padding", "mask_type", "paged_kv_cache", "type_id", "in_flight_batching"]
    return [(k, np.array([f[k]], dtype=dt.get(k, np.int32))) for k in order]


def test_registry_identity_and_idempotent_init():
    for name in ("WeightOnlyQuantMatmul", "GPTAttention"):
        c1 = get_plugin_creator(name, "1", "tensorrt_llm")
        c2 = get_plugin_creator(name, "1", "tensorrt_llm")  # second initLibNvInferPlugins call: same creator
        assert c1 and c1 == c2
    assert not get_plugin_creator("WeightOnlyQuantMatmul", "2", "tensorrt_llm")
    assert not get_plugin_creator("WeightOnlyQuantMatmul", "1", "")
    assert not get_plugin_creator("Gemm", "1", "tensorrt_llm")  # out of scope, deliberately absent


def test_creator_field_names():
    assert TrtPlugin.field_names("WeightOnlyQuantMatmul") == ["type_id", "weight_type_id"]
    assert TrtPlugin.field_names("GPTAttention") == [
        "num_heads", "head_size", "unidirectional", "q_scaling", "rotary_embedding_dim", "neox_rotary_style",
        "context_fmha_type", "multi_block_mode", "multi_query_mode", "int8_kv_cache", "fp8_kv_cache",
        "remove_input_padding", "mask_type", "paged_kv_cache", "type_id", "in_flight_batching"]


def test_woq_plugin_contract():
    p = TrtPlugin.create("WeightOnlyQuantMatmul", woq_fields())
    assert (p.plugin_type, p.plugin_version, p.plugin_namespace, p.num_outputs) == ("WeightOnlyQuantMatmul", "1", "tensorrt_llm", 1)
    # serialization: DataType (4 B) || int weightTypeId (4 B), little endian  (weightOnlyQuantMatmulPlugin.cpp:256-267)
    blob = p.serialize()
    assert blob == struct.pack("<ii", HALF, 1)
    q = TrtPlugin.deserialize("WeightOnlyQuantMatmul", blob)
    assert q.serialize() == blob and q.clone().serialize() == blob
    with pytest.raises(RuntimeError):
        TrtPlugin.deserialize("WeightOnlyQuantMatmul", blob + b"\0")  # exact length is asserted
    # shape inference: [.., K] x weight [K, N/4] (int8 bytes viewed as float32) -> [.., N]
    ins = [((3, 5, 1280), "float16"), ((1280, 960), "float32"), ((3840,), "float16")]
    assert p.output_dims(0, ins) == (3, 5, 3840)
    assert p.output_dtype(0, ["float16", "float32", "float16"]) == HALF
    io = ins + [((3, 5, 3840), "float16")]
    assert all(p.supports_format(pos, io, 3) for pos in range(4))
    bad = list(io)
    bad[1] = ((1280, 960), "int8")  # the weight slot must be declared kFLOAT (the reference's int8-as-float hack)
    assert not p.supports_format(1, bad, 3)
    bad = list(io)
    bad[0] = ((3, 5, 1280), "float32")
    assert not p.supports_format(0, bad, 3)
    assert p.workspace_size(ins, [io[3]]) > 0
    for pl in (p, q):
        pl.destroy()
    # unsupported configurations are rejected at creation like the reference's PLUGIN_ASSERT(false) (-> nullptr)
    with pytest.raises(RuntimeError):
        TrtPlugin.create("WeightOnlyQuantMatmul", woq_fields(type_id=FLOAT))
    with pytest.raises(RuntimeError):
        TrtPlugin.create("WeightOnlyQuantMatmul", woq_fields(weight_type_id=2))  # int4: out of scope on B200
    with pytest.raises(RuntimeError):
        TrtPlugin.create("WeightOnlyQuantMatmul", woq_fields()[:1])  # missing field


def test_gpt_attention_plugin_contract():
    p = TrtPlugin.create("GPTAttention", attn_fields())
    assert (p.plugin_type, p.plugin_version, p.plugin_namespace, p.num_outputs) == ("GPTAttention", "1", "tensorrt_llm", 2)
    blob = p.serialize()
    assert len(blob) == 38
    # int numHeads, int headSize, int unidirectional, float qScaling, int rotaryDim, 8 x bool, int maskType,
    # bool pagedKV, DataType type (gptAttentionCommon.cpp:862-890) || bool inFlightBatching (gptAttentionPlugin.cpp:443-455)
    expect = struct.pack("<iiifi8?i?i?", 20, 64, 1, 1.0, 0, False, False, False, False, False, True, False, False, 1,
                         False, HALF, False)
    assert blob == expect
    q = TrtPlugin.deserialize("GPTAttention", blob)
    assert q.serialize() == blob and q.clone().serialize() == blob
    with pytest.raises(RuntimeError):
        TrtPlugin.deserialize("GPTAttention", blob[:37])
    B, S, Smax = 16, 1, 448
    ins = [((B, S, 3 * 1280), "float16"), ((B, 2, 20, Smax, 64), "int8"), ((B,), "int32"), ((2,), "int32"),
           ((B, Smax), "int32"), ((B,), "int32"), ((4,), "int32"), ((B, 1, Smax), "int32"), ((1,), "float32"),
           ((1,), "float32")]
    assert p.output_dims(0, ins) == (B, S, 1280)
    assert p.output_dims(1, ins) == (B, 2, 20, Smax, 64)
    outs = [((B, S, 1280), "float16"), ((B, 2, 20, Smax, 64), "int8")]
    io = ins + outs
    assert all(p.supports_format(pos, io, len(ins)) for pos in range(len(io)))
    bad = list(io)
    bad[1] = ((B, 2, 20, Smax, 64), "float16")  # int8 KV cache on: the cache must be kINT8
    assert not p.supports_format(1, bad, len(ins))
    bad = list(io)
    bad[8] = ((1,), "float16")  # KV scales are kFLOAT
    assert not p.supports_format(8, bad, len(ins))
    assert p.output_dtype(1, ["float16", "int8"] + ["int32"] * 6 + ["float32"] * 2) == 2
    # paged KV cache: input 1 is the block pool, the block pointers arrive as int32 pairs in slot 10 (8 without int8 KV)
    pg = TrtPlugin.create("GPTAttention", attn_fields(paged_kv_cache=1))
    assert pg.serialize()[32] == 1
    pins = list(ins)
    pins[1] = ((64, 2, 20, 64, 64), "int8")
    pins.append(((B, 1, 2, 2 * 7), "int32"))
    pio = pins + [((B, S, 1280), "float16"), ((64, 2, 20, 64, 64), "int8")]
    assert all(pg.supports_format(pos, pio, len(pins)) for pos in range(len(pio)))
    bad = list(pio)
    bad[10] = ((B, 1, 2, 2 * 7), "float16")
    assert not pg.supports_format(10, bad, len(pins))
    assert pg.output_dims(1, pins) == (64, 2, 20, 64, 64)
    # context FMHA flags round-trip through the byte layout (type 2 -> enable + force fp32 acc)
    r = TrtPlugin.create("GPTAttention", attn_fields(context_fmha_type=2, int8_kv_cache=0))
    b2 = r.serialize()
    assert b2[21:23] == b"\x01\x01" and b2[25] == 0
    with pytest.raises(RuntimeError):
        TrtPlugin.create("GPTAttention", attn_fields(in_flight_batching=1))  # requires remove_input_padding (ctor assert)
    with pytest.raises(RuntimeError):
        TrtPlugin.create("GPTAttention", attn_fields()[:-1])  # missing in_flight_batching


def test_gpt_attention_unsupported_config_fails_loudly_at_enqueue():
    # constructible + serializable (engines round-trip), but it must not compute something else
    p = TrtPlugin.create("GPTAttention", attn_fields(rotary_embedding_dim=32))
    ins = [((1, 1, 192), "float16"), ((1, 2, 1, 8, 64), "int8"), ((1,), "int32"), ((2,), "int32"), ((1, 8), "int32"),
           ((1,), "int32"), ((1,), "int32"), ((1, 1, 8), "int32"), ((1,), "float32"), ((1,), "float32")]
    outs = [((1, 1, 64), "float16"), ((1, 2, 1, 8, 64), "int8")]
    host = np.array([0, 1], np.int32)
    rc = p.enqueue(ins, outs, [0, 0, 0, host.ctypes.data, 0, 0, 0, 0, 0, 0], [0, 0], None, None)
    assert rc == 2  # B200_ERR_UNSUPPORTED
