"""run.py -- the reference's `examples/whisper/run.py` entry point (T/examples/whisper/run.py:25-66) over this library:
an engine directory written by build.py + one audio file in, text (or token ids) out.

    audio -> pad_or_trim -> log_mel_spectrogram -> WhisperEncoding.get_audio_features
          -> WhisperDecoding.detect_language -> main_loop -> post_process

`WhisperEncoding` / `WhisperDecoding` keep the names and the call sequence of the reference's encoding.py /
decoding.py sessions; inside they run the module classes of b200_whisper.models (WhisperEncoder, CrossAttn_KV,
WhisperDecoder) with the GPTAttention-plugin contract: a context pass over the prompt, then one generation step per
token against the in-place int8 KV caches.  (The serving path with folded LayerNorm, fused epilogues and one CUDA-graph
replay per token is b200_whisper.runtime.WhisperPipeline; tests check both produce the same tokens.)

Same flags as the reference (`--log_level`, `--engine_dir`, `--input_file`) plus `--vocab` (a tiktoken rank file:
without it the token ids are printed, no vocabulary ships with the repo) and `--max_new_tokens`.  ffmpeg is not part
of this image, so `--input_file` is a 16 kHz mono .wav (PCM16 / float32) or a .npy waveform instead of any container
ffmpeg can read."""
import argparse
import json
import os
import sys
import time
import wave
from pathlib import Path

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

from build import MODEL_CROSSATTN_NAME, MODEL_DECODER_NAME, MODEL_ENCODER_NAME, get_engine_name  # noqa: E402


def parse_arguments(args=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--log_level', type=str, default='error')
    parser.add_argument('--engine_dir', type=str, default='whisper_outputs')
    parser.add_argument('--input_file', type=str, default='test.m4a')
    parser.add_argument('--vocab', type=str, default=None, help='tiktoken rank file (multilingual.tiktoken)')
    parser.add_argument('--max_new_tokens', type=int, default=None, help='default: n_text_ctx // 2 (decoding.py:324)')
    return parser.parse_args(args)


def load_audio(path, sr=16000):
    """whisper_utils.load_audio (T/examples/whisper/whisper_utils.py:24-53) without ffmpeg: .npy or 16 kHz mono .wav"""
    if path.endswith(".npy"):
        return np.load(path).astype(np.float32).reshape(-1)
    if not path.endswith(".wav"):
        raise RuntimeError(f"{path}: only .wav / .npy inputs (the reference shells out to ffmpeg, absent from this image)")
    with wave.open(path, "rb") as w:
        if w.getframerate() != sr or w.getnchannels() != 1:
            raise RuntimeError(f"{path}: expected {sr} Hz mono")
        raw, width = w.readframes(w.getnframes()), w.getsampwidth()
    if width == 2:
        return np.frombuffer(raw, np.int16).astype(np.float32) / 32768.0
    if width == 4:
        return np.frombuffer(raw, np.float32).copy()
    raise RuntimeError(f"{path}: unsupported sample width {width}")


def _load_engine(engine_dir, engine_name, config_name):
    import torch
    engine_dir = Path(engine_dir)
    with open(engine_dir / config_name) as f:
        config = json.load(f)
    tensors = torch.load(engine_dir / engine_name, map_location="cpu", weights_only=True)
    return config, tensors


class WhisperEncoding:
    """encoding.py's session wrapper: `get_audio_features(mel)`."""

    def __init__(self, engine_dir, device="cuda"):
        from b200_whisper.models import WhisperEncoder
        config, tensors = _load_engine(engine_dir, get_engine_name(MODEL_ENCODER_NAME, 'float16', 1, 0), 'encoder_config.json')
        d = config["dims"]
        self.dims = d
        self.model = WhisperEncoder(d['n_mels'], d['n_audio_ctx'], d['n_audio_state'], d['n_audio_head'], d['n_audio_layer'])
        self.model.load_state_dict(tensors)
        self.model.to(device)

    def get_audio_features(self, mel):
        import torch
        with torch.no_grad():
            return self.model(mel)


class WhisperDecoding:
    """decoding.py's session wrapper: `detect_language`, `main_loop`, `post_process` (decoding.py:703-741,743-783,
    823-870) for greedy decoding with the default logit filters."""

    def __init__(self, engine_dir, device="cuda", vocab=None):
        import torch
        from b200_whisper.models import CrossAttn_KV, WhisperDecoder
        from b200_whisper.quantization import QuantMode
        from b200_whisper.tokenizer import Tokenizer, get_encoding, special_tokens
        config, tensors = _load_engine(engine_dir, get_engine_name(MODEL_DECODER_NAME, 'float16', 1, 0), 'decoder_config.json')
        d = config["dims"]
        qm = QuantMode(config["builder_config"]["quant_mode"])
        self.dims, self.device = d, torch.device(device)
        self.decoder = WhisperDecoder(d['n_vocab'], d['n_text_ctx'], d['n_text_state'], d['n_text_head'], d['n_text_layer'],
                                      quant_mode=qm)
        self.decoder.load_state_dict(tensors)
        self.decoder.to(device)
        _, tensors = _load_engine(engine_dir, get_engine_name(MODEL_CROSSATTN_NAME, 'float16', 1, 0), 'crossattn_config.json')
        self.cross_kv = CrossAttn_KV(d['n_text_state'], d['n_text_head'], d['n_text_layer'], quant_mode=qm)
        self.cross_kv.load_state_dict(tensors)
        self.cross_kv.to(device)
        # decoding.py:452-486: every vocabulary but the English-only one (51864 ids) is multilingual and defaults to
        # language "en" / task "transcribe"; the language token is replaced by the detected one in main_loop
        self.multilingual = d['n_vocab'] != 51864
        lang, task = ("en", "transcribe") if self.multilingual else (None, None)
        self.tokenizer = Tokenizer(lang, task, get_encoding(vocab) if vocab else None,
                                   n_ranks=d['n_vocab'] - len(special_tokens(0)))
        self.sample_len = d['n_text_ctx'] // 2
        self.logit_filter = None

    # -- one decoder pass through the module classes ------------------------------------------------------------------
    def _new_caches(self, B):
        import torch
        H, Smax = self.dims['n_text_head'], self.dims['n_text_ctx']
        return [torch.zeros((B, 2, H, Smax, self.dims['n_text_state'] // H), dtype=self.decoder.kv_dtype, device=self.device)
                for _ in range(self.dims['n_text_layer'])]

    def _forward(self, tokens, past_len, caches, cross):
        """tokens [B, S] (S > 1: context phase; S == 1: generation) -> logits of the last position [B, n_vocab] fp32"""
        import torch
        from b200_whisper.layers import RaggedTensor
        B, S = tokens.shape
        is_context = past_len == 0
        lengths = torch.full((B,), S, dtype=torch.int32, device=self.device)
        seq = torch.full((B,), past_len, dtype=torch.int32, device=self.device)
        pkl = torch.tensor([past_len, int(is_context)], dtype=torch.int32)
        x = RaggedTensor.from_row_lengths(tokens, lengths, torch.empty((S,), dtype=torch.int32))
        with torch.no_grad():
            logits = self.decoder(x, positional_embedding=self.decoder.positional_embedding[past_len:past_len + S],
                                  sequence_length=seq, past_key_value_length=pkl, multi_kv_cache=caches, cross_kv_cache=cross)
        return logits[:, -1].float()

    def detect_language(self, audio_features):
        """decoding.py:703-741: one pass over [sot], softmax over the language tokens -> (language tokens, probabilities)"""
        import torch
        B = audio_features.shape[0]
        with torch.no_grad():
            self._cross = self.cross_kv(audio_features)
        tk = self.tokenizer
        if not self.multilingual:
            return None, None
        logits = self._forward(torch.full((B, 1), tk.sot, dtype=torch.int64, device=self.device), 0, self._new_caches(B),
                               self._cross)
        lo, hi = tk.all_language_tokens[0], tk.all_language_tokens[-1] + 1
        probs = logits[:, lo:hi].softmax(-1)
        return (probs.argmax(-1) + lo).tolist(), probs.cpu()

    def main_loop(self, audio_features, prompt=None, n_new=None, use_filters=True, languages=None):
        """decoding.py:743-783: greedy loop -> (tokens [B, n_new], sum_logprobs [B], no_speech_probs [B]).  `languages`:
        per-utterance language tokens from detect_language, written into the prompt's language slot as
        decoding.py:738-739 does when options.language is None."""
        import torch
        from b200_whisper.functional import WhisperLogitFilter
        B = audio_features.shape[0]
        tk = self.tokenizer
        prompt = list(tk.sot_sequence) if prompt is None else list(prompt)
        n_new = self.sample_len if n_new is None else n_new
        n_new = min(n_new, self.dims['n_text_ctx'] - len(prompt))
        if getattr(self, "_cross", None) is None or self._cross[0].shape[0] != B:
            with torch.no_grad():
                self._cross = self.cross_kv(audio_features)
        caches = self._new_caches(B)
        tokens = torch.tensor([prompt] * B, dtype=torch.int64, device=self.device)
        if languages is not None and len(prompt) > 1:
            tokens[:, 1] = torch.tensor(languages, dtype=torch.int64, device=self.device)
        filt = None
        if use_filters:
            has_vocab = tk.encoding is not None
            blank = tk.encode(" ")[0] if has_vocab else 220
            precision = 30.0 / self.dims['n_audio_ctx']
            filt = WhisperLogitFilter(B, self.dims['n_vocab'], tk.eot, tk.no_timestamps, tk.timestamp_begin, blank,
                                      tk.suppress_tokens("-1" if has_vocab else ""), int(round(1.0 / precision)),
                                      device=self.device)
        out, no_speech = [], None
        cur, past = tokens, 0
        for i in range(n_new):
            logits = self._forward(cur, past, caches, self._cross)
            if i == 0 and tk.no_speech is not None and tk.no_speech < logits.shape[1]:
                no_speech = logits.softmax(-1)[:, tk.no_speech].cpu()
            if filt is not None:
                nxt = filt(logits.contiguous(), torch.empty((B,), dtype=torch.int32, device=self.device))
            else:
                nxt = logits.argmax(-1)
            out.append(nxt.long())
            past += cur.shape[1]
            cur = nxt.long().view(B, 1)
        tokens = torch.stack(out, 1).cpu()
        sums = filt.sum_logprobs.cpu() if filt is not None else torch.zeros(B)
        self._cross = None
        return tokens, sums, no_speech

    def post_process(self, tokens, sum_logprobs, no_speech_probs, audio_features=None, languages=None):
        """decoding.py:823-870: cut at end-of-text, decode to text when a vocabulary is attached"""
        tk = self.tokenizer
        results = []
        for b, row in enumerate(tokens.tolist()):
            ids = row[:row.index(tk.eot)] if tk.eot in row else row
            text = tk.decode([t for t in ids if t < tk.eot]).strip() if tk.encoding is not None else None
            results.append({"tokens": ids, "text": text, "sum_logprob": float(sum_logprobs[b]),
                            "no_speech_prob": None if no_speech_probs is None else float(no_speech_probs[b]),
                            "language": None if languages is None else languages[b]})
        return results


def generate(log_level='error', engine_dir='whisper_outputs', input_file='test.m4a', vocab=None, max_new_tokens=None):
    import torch
    from b200_whisper import whisper_utils
    audio = load_audio(input_file)
    whisper_encoding = WhisperEncoding(engine_dir)
    whisper_decoding = WhisperDecoding(engine_dir, vocab=vocab)
    n_samples = 2 * whisper_encoding.dims['n_audio_ctx'] * whisper_utils.HOP_LENGTH
    audio = whisper_utils.pad_or_trim(torch.from_numpy(audio).to('cuda'), n_samples)
    mel = whisper_utils.log_mel_spectrogram(audio, whisper_encoding.dims['n_mels'], dtype=torch.float16)
    mel = mel.unsqueeze(0) if mel.dim() == 2 else mel

    begin_time = time.time()
    audio_features = whisper_encoding.get_audio_features(mel)
    languages, language_probs = whisper_decoding.detect_language(audio_features)
    tokens, sum_logprobs, no_speech_probs = whisper_decoding.main_loop(audio_features, n_new=max_new_tokens,
                                                                       languages=languages)
    result = whisper_decoding.post_process(tokens, sum_logprobs, no_speech_probs, audio_features, languages)
    torch.cuda.synchronize()
    print("transcribe time " + str(time.time() - begin_time))
    result = result[0]
    print(result["text"] if result["text"] is not None else result["tokens"])
    return result


if __name__ == '__main__':
    generate(**vars(parse_arguments()))
