"""Token-id mirror (b200_whisper.tokenizer) against golden values produced by the reference's own Tokenizer and encoding
construction (T/examples/whisper/tokenizer.py:125-265, decoding.py:394-486; tests/golden/make_tokenizer_golden.py).
The vocabulary-free part always runs; text conversion and the vocabulary-derived lists run when the reference's
multilingual.tiktoken is present (build container)."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
VOCAB = "/root/reference/tensorrt_llm_july-release-v1/examples/whisper/assets/multilingual.tiktoken"


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "tokenizer_golden.json")) as f:
        return json.load(f)


def test_special_ids_without_a_vocabulary(golden):
    from b200_whisper.tokenizer import LANGUAGE_CODES, Tokenizer, get_tokenizer, special_tokens
    assert special_tokens() == golden["special_tokens"]
    # the reference lists languages in the iteration order of a set of strings (tokenizer.py:135-137,215-226): compare
    # as sets; the id of every language token is pinned by special_tokens above
    assert sorted(LANGUAGE_CODES) == sorted(golden["all_language_codes"]) and len(LANGUAGE_CODES) == 99
    t = get_tokenizer(True)
    assert (t.language, t.task) == ("en", "transcribe")          # multilingual defaults, decoding.py:472-475
    for k, v in golden["ids"].items():
        assert getattr(t, k) == v, k
    assert t.n_vocab == golden["n_vocab"] == 51865
    assert sorted(t.all_language_tokens) == sorted(golden["all_language_tokens"]) == list(range(50259, 50358))
    for key, want in golden["cases"].items():
        lang, task = key.split("/")
        tk = get_tokenizer(True, language=lang.upper(), task=task)
        assert list(tk.sot_sequence) == want["sot_sequence"]
        assert list(tk.sot_sequence_including_notimestamps) == want["sot_sequence_including_notimestamps"]
        assert tk.language_token == want["language_token"]
    # the prompt the benchmarks use: sot, en, transcribe
    assert list(t.sot_sequence) == [50258, 50259, 50359] and t.no_timestamps == 50363 and t.timestamp_begin == 50364
    assert t.timestamp_token(1.0) == 50414 and abs(t.timestamp_seconds(51864) - 30.0) < 1e-9
    g = get_tokenizer(False)                                       # gpt2 vocabulary: one rank fewer, no language / task
    assert g.eot == 50256 and g.sot_sequence == (50257,) and g.language is None
    with pytest.raises(ValueError, match="Unsupported language"):
        Tokenizer(language="xx")
    with pytest.raises(RuntimeError, match="no vocabulary"):
        t.encode("hello")
    with pytest.raises(ValueError, match="language token"):
        _ = Tokenizer().language_token


@pytest.mark.skipif(not os.path.exists(VOCAB), reason="vocabulary file (reference asset) not present")
def test_text_and_suppression_lists_with_the_vocabulary(golden):
    from b200_whisper.tokenizer import get_tokenizer
    t = get_tokenizer(True, vocab_path=VOCAB)
    assert t.encoding.n_vocab == golden["n_vocab"] and t.special_tokens == golden["special_tokens"]
    for case in golden["texts"]:
        assert t.encode(case["text"]) == case["ids"]
        assert t.decode(case["ids"]) == case["text"]
    d = golden["decode"]
    assert t.decode(d["ids"]) == d["plain"] and t.decode_with_timestamps(d["ids"]) == d["with_timestamps"]
    assert list(t.non_speech_tokens) == golden["non_speech_tokens"]
    assert list(t.suppress_tokens("-1")) == golden["suppress_default"]
    assert list(t.suppress_tokens("")) == sorted([t.transcribe, t.translate, t.sot, t.sot_prev, t.sot_lm, t.no_speech])
    assert t.encode(" ") == [220]                                  # the blank SuppressBlank removes (decoding.py:202-209)


def test_suppression_list_feeds_the_device_filter(golden):
    """The bitmap b200_whisper_filtered_argmax consumes, built from the reference's default suppression list."""
    import numpy as np
    ids = golden["suppress_default"]
    bitmap = np.zeros((51865 + 31) // 32, dtype=np.uint32)
    for v in ids:
        bitmap[v // 32] |= np.uint32(1 << (v % 32))
    assert int(sum(bin(int(w)).count("1") for w in bitmap)) == len(ids) == 88
