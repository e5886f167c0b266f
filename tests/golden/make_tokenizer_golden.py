"""Generates tests/golden/tokenizer_golden.json from the REFERENCE tokenizer: the `Tokenizer` dataclass of
T/examples/whisper/tokenizer.py:125-265 over the encoding that T/examples/whisper/decoding.py:423-450 builds from
assets/multilingual.tiktoken (decoding.py imported from /root/reference with its TensorRT imports stubbed).
Stored: every special id, sot sequences, non_speech_tokens, the default suppression list and a few text round trips.
Run:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_tokenizer_golden.py
"""
import json
import os
import sys
import types

W = "/root/reference/tensorrt_llm_july-release-v1/examples/whisper"
for name in ("tensorrt_llm", "tensorrt_llm.runtime", "tensorrt_llm.runtime.session", "tensorrt_llm.logger",
             "tensorrt_llm._utils", "build"):
    m = types.ModuleType(name)
    for attr in ("ModelConfig", "SamplingConfig", "Session", "TensorInfo", "str_dtype_to_torch", "str_dtype_to_trt",
                 "trt_dtype_to_torch", "get_engine_name"):
        setattr(m, attr, object)
    sys.modules[name] = m
sys.modules["tensorrt_llm"].runtime = sys.modules["tensorrt_llm.runtime"]
sys.modules["tensorrt_llm"].logger = sys.modules["tensorrt_llm.logger"]
sys.path.insert(0, W)
sys.dont_write_bytecode = True
import decoding as ref_decoding  # noqa: E402
import tokenizer as ref_tokenizer  # noqa: E402

TEXTS = ["Hello, world!", " the quick brown fox", "Привет мир", "你好，世界", " [MUSIC] ♪♪ (applause)", "naïve café — 3.14"]


def main():
    enc = ref_decoding.WhisperDecoding.get_encoding.__wrapped__(None, "multilingual")
    out = {"n_vocab": enc.n_vocab, "cases": {}}
    for lang, task in (("en", "transcribe"), ("zh", "translate"), ("su", "transcribe")):
        t = ref_tokenizer.Tokenizer(encoding=enc, language=lang, task=task)
        out["cases"][f"{lang}/{task}"] = {
            "sot_sequence": list(t.sot_sequence),
            "sot_sequence_including_notimestamps": list(t.sot_sequence_including_notimestamps),
            "language_token": t.language_token,
        }
    t = ref_tokenizer.Tokenizer(encoding=enc, language="en", task="transcribe")
    out["ids"] = {k: getattr(t, k) for k in ("eot", "sot", "translate", "transcribe", "sot_lm", "sot_prev", "no_speech",
                                             "no_timestamps", "timestamp_begin")}
    out["all_language_tokens"] = list(t.all_language_tokens)
    out["all_language_codes"] = list(t.all_language_codes)
    out["special_tokens"] = {k: v for k, v in sorted(t.special_tokens.items(), key=lambda kv: kv[1])}
    out["non_speech_tokens"] = list(t.non_speech_tokens)
    # _get_suppress_tokens (decoding.py:394-421) with the default "-1"
    holder = types.SimpleNamespace(options=types.SimpleNamespace(suppress_tokens="-1"), tokenizer=t)
    out["suppress_default"] = list(ref_decoding.WhisperDecoding._get_suppress_tokens(holder))
    out["texts"] = [{"text": s, "ids": t.encode(s)} for s in TEXTS]
    ts = [50257, 50364, 2425, 11, 1002, 0, 50414]
    out["decode"] = {"ids": ts, "plain": t.decode(ts), "with_timestamps": t.decode_with_timestamps(ts)}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tokenizer_golden.json"), "w") as f:
        json.dump(out, f, ensure_ascii=True, indent=0)
    print("n_vocab", enc.n_vocab, "non-speech", len(out["non_speech_tokens"]), "suppress", len(out["suppress_default"]))


if __name__ == "__main__":
    main()
