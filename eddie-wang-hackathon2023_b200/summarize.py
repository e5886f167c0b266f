"""Word-error-rate harness with the flow of the reference's T/examples/whisper/summarize.py (:56-70 `load_dataset`,
:113-147 the per-file loop, :166-185 the WER report), on top of `runtime.WhisperPipeline`.

The reference scores with two packages that are not in this image: `jiwer.wer` and whisper's `EnglishTextNormalizer`
(its `normalizers/` directory ships empty in the reference tree).  `word_error_rate` restates jiwer's definition --
(substitutions + deletions + insertions) of a minimum edit-distance word alignment, summed over all sentences, divided by
the number of reference words -- and `basic_normalizer` is the language-independent part of the normaliser (lower case,
punctuation to spaces, collapsed white space); pass the real normaliser as `normalizer=` when it is available.
Audio decoding (`load_audio`, ffmpeg) is likewise the caller's: `evaluate` takes waveforms.
"""
import re
from pathlib import Path
from typing import Callable, Iterable, List, Optional, Sequence, Tuple


def load_dataset(dataset_dir) -> Tuple[List[Path], List[str]]:
    """One LibriSpeech chapter directory: the `*.txt` transcript ("<utterance id> <TEXT>" per line) and the audio files
    next to it (summarize.py:56-70).  Returns (audio files sorted by name, references in transcript order)."""
    label_file, audio = None, []
    for f in Path(dataset_dir).iterdir():
        if str(f).endswith("txt"):
            label_file = f
        else:
            audio.append(f)
    if label_file is None:
        raise FileNotFoundError(f"no transcript (*.txt) in {dataset_dir}")
    references = []
    with open(label_file) as fh:
        for line in fh:
            if line.strip():
                references.append(line.split(" ", 1)[1].replace("\n", ""))
    return sorted(audio), references


def clean_hypothesis(text: str) -> str:
    """summarize.py:128-130: drop the punctuation marks . , ! ? and upper-case, to match LibriSpeech transcripts."""
    return re.sub(r"[.,!?]", "", text).upper()


def basic_normalizer(text: str) -> str:
    text = re.sub(r"[^\w\s']", " ", text.lower())
    return re.sub(r"\s+", " ", text).strip()


def _edit_distance(ref: Sequence[str], hyp: Sequence[str]) -> int:
    prev = list(range(len(hyp) + 1))
    for i, r in enumerate(ref, 1):
        cur = [i] + [0] * len(hyp)
        for j, h in enumerate(hyp, 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (r != h))
        prev = cur
    return prev[-1]


def word_error_rate(references: Iterable[str], hypotheses: Iterable[str]) -> float:
    """jiwer.wer(references, hypotheses): total word edit distance over total reference words."""
    errors = words = 0
    references, hypotheses = list(references), list(hypotheses)
    if len(references) != len(hypotheses):
        raise ValueError("references and hypotheses differ in number")
    for r, h in zip(references, hypotheses):
        rw, hw = r.split(), h.split()
        errors += _edit_distance(rw, hw)
        words += len(rw)
    if words == 0:
        raise ValueError("no reference words")
    return errors / words


def evaluate(pipeline, tokenizer, waveforms, references: Sequence[str], sample_len: Optional[int] = None,
             normalizer: Optional[Callable[[str], str]] = None, max_samples: int = 480000):
    """Transcribes `waveforms` (list of 1-D float32 arrays at 16 kHz) and scores them against `references`.
    Utterances longer than 30 s are skipped like the reference does (summarize.py:118-119).
    Returns {"wer": float, "hypotheses": [...], "references": [...], "skipped": n}."""
    if tokenizer.encoding is None:
        raise RuntimeError("a vocabulary is needed to turn tokens into text: get_tokenizer(..., vocab_path=...)")
    normalizer = normalizer or basic_normalizer
    keep = [i for i, w in enumerate(waveforms) if len(w) <= max_samples]
    if not keep:
        raise ValueError(f"no utterance to score: all {len(waveforms)} are longer than {max_samples} samples (30 s)")
    from . import whisper_utils
    import numpy as np
    # one decoder batch at a time (the reference streams one file at a time, summarize.py:112-140): a LibriSpeech split
    # padded to 30 s per utterance is gigabytes of host AND device memory if stacked in one piece
    results = []
    step = max(1, int(getattr(pipeline, "B", 1)))
    for b0 in range(0, len(keep), step):
        batch = np.stack([whisper_utils.pad_or_trim(np.asarray(waveforms[i], dtype=np.float32), pipeline.n_samples)
                          for i in keep[b0:b0 + step]])
        results.extend(pipeline.transcribe(batch, tokenizer, sample_len=sample_len))
    hyps = [clean_hypothesis(r["text"]) for r in results]
    refs = [references[i] for i in keep]
    wer = word_error_rate([normalizer(r) for r in refs], [normalizer(h) for h in hyps])
    return {"wer": wer, "hypotheses": hyps, "references": refs, "skipped": len(waveforms) - len(keep)}
