"""Token-id side of the decoder: special tokens, prompt (sot) sequences, the suppression lists the logit filters
consume, and text <-> ids through tiktoken when a vocabulary file is supplied.

Mirrors the reference's T/examples/whisper/tokenizer.py:125-265 (`Tokenizer`) and the encoding construction of
T/examples/whisper/decoding.py:423-486 (`get_encoding`, `get_tokenizer`): the vocabulary file holds the 50257 (multilingual:
50257, gpt2: 50256) mergeable ranks, and the special tokens are appended after them in a fixed order, so every special id
follows from the number of ranks alone -- endoftext, startoftranscript, one token per language, translate, transcribe,
startoflm, startofprev, nospeech, notimestamps, then 1501 timestamps <|0.00|> ... <|30.00|>.

The vocabulary (assets/multilingual.tiktoken in the reference tree) is data the user supplies by path; without it the
ids, prompts and timestamp arithmetic still work (that is all the GPU decoder needs), only encode/decode and the
vocabulary-derived `non_speech_tokens` need the file."""
import base64
from functools import cached_property, lru_cache
from typing import Dict, List, Optional, Tuple

# order defines the language token ids (tokenizer.py:9-107)
LANGUAGE_CODES = (
    "en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur hr bg lt la mi ml cy "
    "sk te fa lv bn sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si km sn yo so af oc ka be tg sd gu am yi lo "
    "uz fo ht ps tk nn mt sa lb my bo tl mg as tt haw ln ha ba jw su").split()
N_TIMESTAMPS = 1501
MULTILINGUAL_RANKS = 50257
GPT2_PATTERN = r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"""


def special_tokens(n_ranks: int = MULTILINGUAL_RANKS) -> Dict[str, int]:
    """name -> id, appended after the `n_ranks` mergeable tokens (decoding.py:433-450)."""
    names = ["<|endoftext|>", "<|startoftranscript|>"] + [f"<|{c}|>" for c in LANGUAGE_CODES] + [
        "<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>", "<|nospeech|>", "<|notimestamps|>"]
    names += [f"<|{i * 0.02:.2f}|>" for i in range(N_TIMESTAMPS)]
    return {name: n_ranks + i for i, name in enumerate(names)}


@lru_cache(maxsize=None)
def get_encoding(vocab_path: str):
    """tiktoken.Encoding over the ranks of `vocab_path` ("<base64 token> <rank>" per line) plus the special tokens."""
    import tiktoken
    ranks = {}
    with open(vocab_path) as f:
        for line in f:
            if line.strip():
                token, rank = line.split()
                ranks[base64.b64decode(token)] = int(rank)
    specials = special_tokens(len(ranks))
    return tiktoken.Encoding(name=vocab_path.rsplit("/", 1)[-1], explicit_n_vocab=len(ranks) + len(specials),
                             pat_str=GPT2_PATTERN, mergeable_ranks=ranks, special_tokens=specials)


class Tokenizer:
    """Special-token ids of a Whisper vocabulary, and text conversion when an encoding is attached."""

    def __init__(self, language: Optional[str] = None, task: Optional[str] = None, encoding=None,
                 n_ranks: int = MULTILINGUAL_RANKS):
        if language is not None and language not in LANGUAGE_CODES:
            raise ValueError(f"Unsupported language: {language}")
        if task not in (None, "transcribe", "translate"):
            raise ValueError(f"Unsupported task: {task}")
        self.encoding = encoding
        self.language, self.task = language, task
        self.special_tokens = special_tokens(n_ranks if encoding is None else encoding.n_vocab - len(special_tokens(0)))
        seq = [self.sot]
        if language is not None:
            seq.append(self.sot + 1 + LANGUAGE_CODES.index(language))
        if task is not None:
            seq.append(self.transcribe if task == "transcribe" else self.translate)
        self.sot_sequence = tuple(seq)

    # ---- ids ----
    eot = property(lambda self: self.special_tokens["<|endoftext|>"])
    sot = property(lambda self: self.special_tokens["<|startoftranscript|>"])
    translate = property(lambda self: self.special_tokens["<|translate|>"])
    transcribe = property(lambda self: self.special_tokens["<|transcribe|>"])
    sot_lm = property(lambda self: self.special_tokens["<|startoflm|>"])
    sot_prev = property(lambda self: self.special_tokens["<|startofprev|>"])
    no_speech = property(lambda self: self.special_tokens["<|nospeech|>"])
    no_timestamps = property(lambda self: self.special_tokens["<|notimestamps|>"])
    timestamp_begin = property(lambda self: self.special_tokens["<|0.00|>"])
    n_vocab = property(lambda self: self.timestamp_begin + N_TIMESTAMPS)

    @property
    def language_token(self) -> int:
        if self.language is None:
            raise ValueError("This tokenizer does not have language token configured")
        return self.special_tokens[f"<|{self.language}|>"]

    @property
    def all_language_tokens(self) -> Tuple[int, ...]:
        return tuple(self.special_tokens[f"<|{c}|>"] for c in LANGUAGE_CODES)

    @property
    def all_language_codes(self) -> Tuple[str, ...]:
        return tuple(LANGUAGE_CODES)

    @property
    def sot_sequence_including_notimestamps(self) -> Tuple[int, ...]:
        return tuple(list(self.sot_sequence) + [self.no_timestamps])

    def timestamp_token(self, seconds: float) -> int:
        return self.timestamp_begin + int(round(seconds / 0.02))

    def timestamp_seconds(self, token: int) -> float:
        return (token - self.timestamp_begin) * 0.02

    # ---- text (needs the vocabulary) ----
    def _enc(self):
        if self.encoding is None:
            raise RuntimeError("no vocabulary attached: build the tokenizer with get_tokenizer(..., vocab_path=...)")
        return self.encoding

    def encode(self, text, **kwargs) -> List[int]:
        return self._enc().encode(text, **kwargs)

    def decode(self, token_ids, **kwargs) -> str:
        """Text of the ids below the timestamp range (timestamps are dropped, tokenizer.py:157-159)."""
        return self._enc().decode([t for t in token_ids if t < self.timestamp_begin], **kwargs)

    def decode_with_timestamps(self, token_ids, **kwargs) -> str:
        return self._enc().decode(list(token_ids), **kwargs)

    @cached_property
    def non_speech_tokens(self) -> Tuple[int, ...]:
        """Ids suppressed so that speaker tags / music notes / bracketed annotations are never sampled
        (tokenizer.py:231-265): a symbol is suppressed when it is a single token with or without a leading space; the
        musical symbols U+2669..U+266F share their first UTF-8 bytes, so their first token is suppressed either way."""
        enc = self._enc()
        marks = list('"#()*+/:;<=>@[\\]^_`{|}~「」『』')
        marks += "<< >> <<< >>> -- --- -( -[ (' (\" (( )) ((( ))) [[ ]] {{ }} ♪♪ ♪♪♪".split()
        musical = set("♩♪♫♬♭♮♯")
        ids = {enc.encode(" -")[0], enc.encode(" '")[0]}
        for mark in marks + sorted(musical):
            for toks in (enc.encode(mark), enc.encode(" " + mark)):
                if len(toks) == 1 or mark in musical:
                    ids.add(toks[0])
        return tuple(sorted(ids))

    def suppress_tokens(self, suppress="-1") -> Tuple[int, ...]:
        """The SuppressTokens list of decoding.py:391-421: "-1" expands to non_speech_tokens; the task / prompt control
        tokens are always suppressed, and nospeech is too (its probability is read separately)."""
        if isinstance(suppress, str):
            suppress = [int(t) for t in suppress.split(",") if t.strip()]
        suppress = list(suppress or [])
        if -1 in suppress:
            suppress = [t for t in suppress if t >= 0] + list(self.non_speech_tokens)
        suppress += [self.transcribe, self.translate, self.sot, self.sot_prev, self.sot_lm, self.no_speech]
        return tuple(sorted(set(suppress)))


def get_tokenizer(multilingual: bool = True, language: Optional[str] = None, task: Optional[str] = None,
                  vocab_path: Optional[str] = None) -> Tokenizer:
    """decoding.py:452-486: multilingual vocabularies default to language "en" and task "transcribe"; names and the
    aliases of tokenizer.py:110-123 resolve to codes upstream of this call."""
    if language is not None:
        language = language.lower()
    if multilingual:
        language, task = language or "en", task or "transcribe"
    else:
        language = task = None
    encoding = get_encoding(vocab_path) if vocab_path is not None else None
    return Tokenizer(language, task, encoding, n_ranks=MULTILINGUAL_RANKS if multilingual else MULTILINGUAL_RANKS - 1)
