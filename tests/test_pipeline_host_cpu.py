"""Host logic of WhisperPipeline (slicing a list of utterances into decoder-sized batches, padding the last slice, cutting
at end-of-text, attaching language / no-speech results) with the device stages replaced by fakes: no GPU involved."""
import types

import torch

from b200_whisper.runtime.pipeline import WhisperPipeline
from b200_whisper.tokenizer import Tokenizer


class FakeDecoder:
    """Emits, for row r of a slice, the tokens [id, id + 1, ...] where id = 1000 + 10 * (utterance tag): the tag is
    carried by the fake mel / encoder output, so the test can tell which utterance each result row came from."""

    def __init__(self, B, eot):
        self.B, self.eot = B, eot
        self.calls = []
        self.logit_filter = types.SimpleNamespace(sum_logprobs=torch.zeros(B))

    def set_encoder_output(self, xa):
        assert xa.shape[0] == self.B            # the decoder batch is fixed: short slices arrive padded
        self.tags = xa[:, 0, 0].long().tolist()

    def decode(self, prompts, n_new, use_graph=True):
        assert len(prompts) == self.B and all(len(p) == len(prompts[0]) for p in prompts)
        self.calls.append((list(self.tags), list(prompts[0]), n_new))
        self.prompt_rows = getattr(self, "prompt_rows", []) + [[list(p) for p in prompts]]
        rows = []
        for tag in self.tags:
            row = [1000 + 10 * tag + i for i in range(n_new)]
            if tag % 2 == 1 and n_new > 3:       # odd utterances end early: eot at position 3, then eot forever
                row[3:] = [self.eot] * (n_new - 3)
            rows.append(row)
        self.logit_filter.sum_logprobs = torch.tensor([-float(t) for t in self.tags])
        return torch.tensor(rows, dtype=torch.int32)

    def detect_language(self, sot, lo, hi, no_speech):
        n = hi - lo
        lang = torch.tensor([lo + (t % n) for t in self.tags], dtype=torch.int32)
        probs = torch.zeros((self.B, n))
        for r, t in enumerate(self.tags):
            probs[r, t % n] = 1.0
        return lang, probs, torch.tensor([0.01 * t for t in self.tags])

    def enable_logit_filters(self, *args):
        self.filters = args


def make_pipe(B, tk):
    pipe = object.__new__(WhisperPipeline)
    pipe.B = B
    pipe.dims = types.SimpleNamespace(n_text_ctx=448, n_audio_ctx=1500, n_mels=80)
    pipe.decoder = FakeDecoder(B, tk.eot)
    pipe.log_mel = lambda audio: torch.as_tensor(audio, dtype=torch.float32).view(-1, 1, 1).expand(-1, 80, 4).contiguous()
    pipe.get_audio_features = lambda mel: mel.clone()
    return pipe


def test_slices_padding_and_order():
    tk = Tokenizer("en", "transcribe")
    pipe = make_pipe(4, tk)
    tags = list(range(10))                       # 10 utterances, decoder batch 4 -> slices of 4, 4, 2 (+2 rows of padding)
    out = pipe.transcribe_tokens(tags, [1, 2, 3], 5)
    assert tuple(out.shape) == (10, 5) and out.dtype == torch.int64
    assert out[:, 0].tolist() == [1000 + 10 * t for t in tags]
    assert [c[0] for c in pipe.decoder.calls] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 8, 8]]   # padded with the slice's first row
    assert all(c[1] == [1, 2, 3] and c[2] == 5 for c in pipe.decoder.calls)


def test_transcribe_cuts_at_eot_and_attaches_language():
    tk = Tokenizer("en", "transcribe")
    pipe = make_pipe(2, tk)
    res = pipe.transcribe([0, 1, 2], tk, sample_len=6, detect_language=True)
    assert pipe.decoder.filters[0] == tk.eot and pipe.decoder.filters[2] == tk.timestamp_begin
    assert pipe.decoder.filters[3] == 220 and pipe.decoder.filters[5] == 50          # blank id, 1.0 s / 0.02 s
    assert set(pipe.decoder.filters[4]) == {tk.transcribe, tk.translate, tk.sot, tk.sot_prev, tk.sot_lm, tk.no_speech}
    # the reference's default (options.language = None): every utterance is decoded from ITS detected language token
    # (decoding.py:738-739); two different languages in one batch
    lang0 = tk.all_language_tokens[0]
    sot, task = tk.sot_sequence[0], tk.sot_sequence[2]
    assert pipe.decoder.prompt_rows == [[[sot, lang0 + 0, task], [sot, lang0 + 1, task]],
                                        [[sot, lang0 + 2, task], [sot, lang0 + 2, task]]]   # slice 2: utterance 2 + padding
    # without detection the tokenizer's language is kept for every row (options.language set)
    pipe2 = make_pipe(2, tk)
    pipe2.transcribe([0, 1], tk, sample_len=4)
    assert pipe2.decoder.prompt_rows == [[list(tk.sot_sequence)] * 2]
    assert [len(r["tokens"]) for r in res] == [6, 3, 6]                                # utterance 1 ended at its eot
    assert res[1]["tokens"] == [1010, 1011, 1012] and tk.eot not in res[1]["tokens"]
    assert [r["sum_logprob"] for r in res] == [0.0, -1.0, -2.0]
    codes = tk.all_language_codes
    assert [r["language"] for r in res] == [codes[0], codes[1], codes[2]]
    assert [round(r["no_speech_prob"], 4) for r in res] == [0.0, 0.01, 0.02]
    assert res[2]["language_probs"][codes[2]] == 1.0 and res[0]["text"] is None
    # sample_len is capped by the text context
    pipe.decoder.calls.clear()
    pipe.transcribe([0], tk, sample_len=10_000)
    assert pipe.decoder.calls[0][2] == 448 - len(tk.sot_sequence)
    pipe.decoder.calls.clear()
    pipe.transcribe([0], tk)
    assert pipe.decoder.calls[0][2] == 224                                            # default: n_text_ctx // 2 (decoding.py:324)
