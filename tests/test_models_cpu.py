"""The reference's call sites kept: module classes under its names (T/tensorrt_llm/layers/attention.py:48-415,
layers/conv.py:52-94, models/whisper/model.py:124-172,201-300,469-555) and the build.py / run.py flag sets
(T/examples/whisper/build.py:42-143, run.py:25-31).  CPU: names, constructor arguments, parameter names, argument
parsing -- nothing computes without a GPU."""
import importlib.util
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(f"examples_whisper_{name}", os.path.join(ROOT, "examples", "whisper", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_build_flags_of_the_reference_parse():
    from b200_whisper.quantization import QuantMode
    build = _load("build")
    # the README's command line of the hackathon entry
    args = build.parse_arguments(["--use_weight_only", "--weight_only_precision", "int8", "--int8_kv_cache",
                                  "--use_gpt_attention_plugin", "--model_dir", "large-v2.pt", "--quantize_dir", "quantize/1-gpu",
                                  "--output_dir", "out", "--max_batch_size", "16", "--use_gemm_plugin", "float16"])
    assert args.use_weight_only and args.int8_kv_cache and args.weight_only_precision == "int8"
    assert args.use_gpt_attention_plugin == "float16"       # flag without a value -> the model dtype (build.py:123-130)
    assert args.use_gemm_plugin == "float16" and args.use_layernorm_plugin is False
    assert args.quant_mode == QuantMode.use_weight_only().set_int8_kv_cache()
    assert args.quant_mode.is_int8_weight_only() and args.quant_mode.has_int8_kv_cache()
    build.check_supported(args)
    # defaults of the reference
    d = build.parse_arguments([])
    assert (d.world_size, d.model_dir, d.quantize_dir, d.dtype, d.max_batch_size, d.max_input_len, d.max_output_len,
            d.max_beam_width, d.output_dir) == (1, "large-v2.pt", "quantize/1-gpu", "float16", 256, 200, 200, 1, "whisper_outputs")
    assert d.use_gpt_attention_plugin is False and d.quant_mode == QuantMode(0)
    with pytest.raises(ValueError):
        build.check_supported(d)                            # no fp16 Linear path: weight-only int8 is the product
    with pytest.raises(ValueError):
        build.check_supported(build.parse_arguments(["--use_weight_only", "--weight_only_precision", "int4"]))
    with pytest.raises(ValueError):
        build.check_supported(build.parse_arguments(["--use_weight_only", "--world_size", "2"]))
    with pytest.raises(SystemExit):
        build.parse_arguments(["--dtype", "int8"])
    assert build.get_engine_name(build.MODEL_DECODER_NAME, "float16", 1, 0) == "whisper_decoder_float16_tp1_rank0.engine"
    assert build.MODEL_CROSSATTN_NAME == "whsiper_crossattn"


def test_run_flags_of_the_reference_parse():
    run = _load("run")
    a = run.parse_arguments([])
    assert (a.log_level, a.engine_dir, a.input_file) == ("error", "whisper_outputs", "test.m4a")
    a = run.parse_arguments(["--engine_dir", "e", "--input_file", "x.wav", "--log_level", "info"])
    assert (a.engine_dir, a.input_file) == ("e", "x.wav")
    with pytest.raises(RuntimeError):
        run.load_audio("test.m4a")                          # ffmpeg containers are refused loudly, not silently zeroed


def test_load_audio_wav_roundtrip(tmp_path):
    import wave

    import numpy as np
    run = _load("run")
    x = (np.sin(np.arange(1600) * 0.05) * 12000).astype(np.int16)
    p = str(tmp_path / "a.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(1), w.setsampwidth(2), w.setframerate(16000)
        w.writeframes(x.tobytes())
    got = run.load_audio(p)
    assert got.dtype == np.float32 and np.array_equal(got, x.astype(np.float32) / 32768.0)
    np.save(str(tmp_path / "a.npy"), got)
    assert np.array_equal(run.load_audio(str(tmp_path / "a.npy")), got)


def test_module_classes_keep_the_reference_names_and_parameters():
    from b200_whisper.layers import Attention, AttentionMaskType, Conv1d, LayerNorm
    from b200_whisper.models import CrossAttn_KV, WhisperDecoder, WhisperEncoder
    from b200_whisper.quantization import QuantMode
    from b200_whisper.quantization.layer import WeightOnlyQuantLinear, WeightOnlyQuantRowLinear
    qm = QuantMode.use_weight_only().set_int8_kv_cache()
    att = Attention(128, 2, 64, use_int8_kv_cache=True, quant_mode=qm)
    assert isinstance(att.qkv, WeightOnlyQuantLinear) and isinstance(att.dense, WeightOnlyQuantRowLinear)
    assert att.qkv.out_features == 384 and att.q_linear is None and att.attention_head_size == 64
    assert {"kv_orig_quant_scale", "kv_quant_orig_scale"} <= set(dict(att.named_buffers()))
    x_att = Attention(128, 2, 64, cross_attention=True, quant_mode=qm)
    assert x_att.qkv is None and x_att.q_linear.out_features == 128
    with pytest.raises(ValueError):
        Attention(128, 2, 64, tp_size=2)
    conv = Conv1d(80, 128, kernel_size=3, stride=2, padding=1)
    assert tuple(conv.weight.shape) == (128, 80, 3, 1) and tuple(conv.bias.shape) == (128,)   # conv.py:75-80: 4-D weight
    ln = LayerNorm(128)
    assert tuple(ln.weight.shape) == (128,)

    enc = WhisperEncoder(80, 96, 128, 2, 2)
    dec = WhisperDecoder(1024, 64, 128, 2, 2, quant_mode=qm)
    ckv = CrossAttn_KV(128, 2, 2, quant_mode=qm)
    names = set(enc.state_dict())
    # the attribute paths examples/whisper/weight.py:40-110 assigns
    for n in ("conv1.weight", "conv2.bias", "positional_embedding", "blocks.0.attn.qkv.weight", "blocks.0.attn.qkv.per_channel_scale",
              "blocks.1.attn.dense.bias", "blocks.0.attn_ln.weight", "blocks.0.mlp.fc.weight", "blocks.0.mlp.proj.weight",
              "blocks.0.mlp_ln.bias", "ln_post.weight"):
        assert n in names, n
    names = set(dec.state_dict())
    for n in ("token_embedding_weight", "positional_embedding", "blocks.0.cross_attn.q_linear.weight", "blocks.0.cross_attn.dense.weight",
              "blocks.1.cross_attn_ln.weight", "blocks.0.attn.kv_quant_orig_scale", "ln.bias"):
        assert n in names, n
    assert dec.kv_dtype == torch.int8 and dec.blocks[0].attn.attention_mask_type == AttentionMaskType.causal
    names = set(ckv.state_dict())
    assert {"blocks.0.key.weight", "blocks.1.value.bias", "kv_orig_quant_scale"} <= names and "blocks.0.key.bias" not in names
    with pytest.raises(ValueError):
        WhisperDecoder(1024, 64, 128, 2, 2, quant_mode=QuantMode(0))    # no fp16 Linear path on the hot path
