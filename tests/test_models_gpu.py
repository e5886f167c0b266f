"""build.py -> engine directory -> run.py's WhisperEncoding / WhisperDecoding (the module classes under the reference's
names, one library kernel per operator) against runtime.WhisperPipeline (fused epilogues, folded LayerNorm, CUDA-graph
loop) and the oracle, on the MICRO model: the reference's call sites and the serving path give the same tokens."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
from log_mel_cases import speech_like  # noqa: E402

from oracle import whisper_oracle as wo  # noqa: E402

pytestmark = pytest.mark.gpu


def _load(name):
    d = os.path.join(ROOT, "examples", "whisper")
    if d not in sys.path:
        sys.path.insert(0, d)
    spec = importlib.util.spec_from_file_location(name, os.path.join(d, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def test_build_then_run_equals_pipeline_and_oracle(tmp_path):
    from b200_whisper.runtime import WhisperPipeline, save_checkpoint, write_kv_scales
    build, run = _load("build"), _load("run")
    dims = wo.MICRO
    B, n_new, prompt = 2, 6, [3, 7, 11]
    n = 2 * dims.n_audio_ctx * 160
    audio = np.stack([speech_like(n, 71), speech_like(n, 72, amp=0.3)])
    sd = wo.synthetic_state_dict(dims, seed=1)
    sdq = wo.quantize_state_dict(sd, dims, decoder_only=False)
    kv_s, ckv_s = [0.05] * dims.n_text_layer, [0.04] * dims.n_text_layer
    save_checkpoint(str(tmp_path / "micro.pt"), dims, sd)
    write_kv_scales(str(tmp_path / "q"), kv_s, ckv_s)
    kv_s = [float(np.float32(s)) for s in kv_s]
    ckv_s = [float(np.float32(s)) for s in ckv_s]

    # the reference's command line
    args = build.run_build(["--model_dir", str(tmp_path / "micro.pt"), "--quantize_dir", str(tmp_path / "q"),
                            "--output_dir", str(tmp_path / "engines"), "--use_weight_only", "--weight_only_precision", "int8",
                            "--int8_kv_cache", "--use_gpt_attention_plugin"])
    for f in ("whisper_encoder_float16_tp1_rank0.engine", "whisper_decoder_float16_tp1_rank0.engine",
              "whsiper_crossattn_float16_tp1_rank0.engine", "encoder_config.json", "decoder_config.json", "crossattn_config.json"):
        assert os.path.exists(os.path.join(args.output_dir, f)), f

    enc = run.WhisperEncoding(args.output_dir)
    dec = run.WhisperDecoding(args.output_dir)
    pipe = WhisperPipeline.from_files(str(tmp_path / "micro.pt"), str(tmp_path / "q"), batch_size=B)
    mel = pipe.log_mel(audio)
    xa_mod = enc.get_audio_features(mel)
    xa_pipe = pipe.get_audio_features(mel)
    with torch.no_grad():
        xa_ref = wo.encoder_forward(sdq, dims, mel.float().cpu())
    scale = max(1.0, xa_ref.abs().max().item())
    assert (xa_mod.float().cpu() - xa_ref).abs().max().item() <= 3e-2 * scale
    assert (xa_mod.float() - xa_pipe.float()).abs().max().item() <= 3e-2 * scale

    # same encoder output into both decoders and the oracle
    tokens_mod, _, _ = dec.main_loop(xa_pipe, prompt=prompt, n_new=n_new, use_filters=False)
    pipe.decoder.set_encoder_output(xa_pipe)
    tokens_pipe = pipe.decoder.decode([prompt] * B, n_new).long().cpu()
    with torch.no_grad():
        ref_tokens, _ = wo.greedy_decode(sdq, dims, xa_pipe.float().cpu(), prompt, n_new, kv_s, ckv_s, act_fp16=True)
    assert tokens_mod.tolist() == ref_tokens.tolist()
    assert tokens_mod.tolist() == tokens_pipe.tolist()


def test_run_generate_from_a_wav_file(tmp_path, capsys):
    """`python run.py --engine_dir ... --input_file x.wav` end to end (language detection, filters, post-processing)"""
    import wave

    from b200_whisper.runtime import save_checkpoint
    build, run = _load("build"), _load("run")
    dims = wo.ModelDimensions(80, 96, 128, 2, 2, 2048, 64, 128, 2, 2)   # 440 ranks + the 1608 special tokens
    sd = wo.synthetic_state_dict(dims, seed=2)
    save_checkpoint(str(tmp_path / "m.pt"), dims, sd)
    build.run_build(["--model_dir", str(tmp_path / "m.pt"), "--output_dir", str(tmp_path / "e"), "--use_weight_only",
                     "--use_gpt_attention_plugin", "float16"])     # fp16 KV caches: no quantize_dir needed
    x = (speech_like(2 * dims.n_audio_ctx * 160, 81, amp=0.5) * 32767).astype(np.int16)
    p = str(tmp_path / "x.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(1), w.setsampwidth(2), w.setframerate(16000)
        w.writeframes(x.tobytes())
    res = run.generate(engine_dir=str(tmp_path / "e"), input_file=p, max_new_tokens=12)
    out = capsys.readouterr().out
    assert "transcribe time" in out
    ids = res["tokens"]
    assert 0 < len(ids) <= 12 and res["text"] is None and np.isfinite(res["sum_logprob"])
    assert 441 + 1 <= res["language"] < 441 + 1 + 99           # a language token was detected and used in the prompt
    assert 0.0 <= res["no_speech_prob"] <= 1.0
    assert ids[0] >= 542 + 2                                   # ApplyTimestampRules: decoding starts with a timestamp
