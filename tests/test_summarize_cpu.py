"""WER harness (b200_whisper.summarize; flow of T/examples/whisper/summarize.py:56-185): jiwer's WER definition on
hand-computed cases, the LibriSpeech chapter loader, the hypothesis clean-up, and `evaluate` over a fake pipeline."""
import types

import numpy as np
import pytest

from b200_whisper import summarize as sm


def test_word_error_rate_matches_the_jiwer_definition():
    assert sm.word_error_rate(["the cat sat"], ["the cat sat"]) == 0.0
    assert sm.word_error_rate(["the cat sat"], ["the bat sat"]) == pytest.approx(1 / 3)          # one substitution
    assert sm.word_error_rate(["the cat sat"], ["the sat"]) == pytest.approx(1 / 3)              # one deletion
    assert sm.word_error_rate(["the cat sat"], ["the big cat sat"]) == pytest.approx(1 / 3)      # one insertion
    assert sm.word_error_rate(["a b c d"], ["d c b a"]) == pytest.approx(1.0)                    # 4 edits (2 sub + ...) / 4
    assert sm.word_error_rate(["a"], ["a b c"]) == pytest.approx(2.0)                            # WER can exceed 1
    # corpus level: errors and words are summed over sentences, not averaged per sentence
    assert sm.word_error_rate(["a b c d e f g h", "x"], ["a b c d e f g h", "y"]) == pytest.approx(1 / 9)
    with pytest.raises(ValueError):
        sm.word_error_rate(["a"], ["a", "b"])
    with pytest.raises(ValueError):
        sm.word_error_rate([""], ["a"])


def test_clean_up_and_normalizer():
    assert sm.clean_hypothesis(" Hello, world! Is it?") == " HELLO WORLD IS IT"                   # summarize.py:128-130
    assert sm.basic_normalizer("  HELLO   World; it's -- fine ") == "hello world it's fine"


def test_load_dataset(tmp_path):
    d = tmp_path / "1272" / "128104"
    d.mkdir(parents=True)
    (d / "1272-128104.trans.txt").write_text("1272-128104-0000 MISTER QUILTER IS THE APOSTLE\n1272-128104-0001 NOR IS HE\n")
    for n in ("1272-128104-0001.flac", "1272-128104-0000.flac"):
        (d / n).write_bytes(b"")
    audio, refs = sm.load_dataset(d)
    assert [a.name for a in audio] == ["1272-128104-0000.flac", "1272-128104-0001.flac"]
    assert refs == ["MISTER QUILTER IS THE APOSTLE", "NOR IS HE"]
    (tmp_path / "empty").mkdir()
    with pytest.raises(FileNotFoundError):
        sm.load_dataset(tmp_path / "empty")


def test_evaluate_over_a_fake_pipeline():
    texts = {0: " Mister Quilter is the apostle.", 1: " Nor is he!", 2: " never transcribed"}

    class FakePipe:
        n_samples = 480000
        B = 2
        shapes = []

        def transcribe(self, batch, tokenizer, sample_len=None):
            assert batch.shape[0] <= self.B and batch.shape[1] == 480000 and batch.dtype == np.float32
            self.shapes.append(batch.shape[0])
            return [{"text": texts[int(row[0])], "tokens": [], "sum_logprob": 0.0} for row in batch]

    waves = [np.full(16000, 0, np.float32), np.full(32000, 1, np.float32), np.full(16000 * 31, 2, np.float32)]
    refs = ["MISTER QUILTER IS THE APOSTLE", "NOR IS SHE", "SKIPPED"]
    tk = types.SimpleNamespace(encoding=object())
    out = sm.evaluate(FakePipe(), tk, waves, refs)
    assert out["skipped"] == 1 and out["hypotheses"] == [" MISTER QUILTER IS THE APOSTLE", " NOR IS HE"]
    assert out["wer"] == pytest.approx(1 / 8)                                         # one substitution in eight words
    assert FakePipe.shapes == [2]                                                     # the 31 s utterance was skipped
    # streamed one decoder batch at a time: five utterances, batch 2 -> slices of 2, 2, 1 (never the whole split at once)
    FakePipe.shapes.clear()
    out = sm.evaluate(FakePipe(), tk, [waves[0], waves[1]] * 2 + [waves[0]], [refs[0], refs[1]] * 2 + [refs[0]])
    assert FakePipe.shapes == [2, 2, 1] and len(out["hypotheses"]) == 5
    with pytest.raises(ValueError, match="no utterance"):
        sm.evaluate(FakePipe(), tk, [waves[2]], ["X"])                                # nothing survives the length filter
    with pytest.raises(RuntimeError, match="vocabulary"):
        sm.evaluate(FakePipe(), types.SimpleNamespace(encoding=None), waves, refs)
