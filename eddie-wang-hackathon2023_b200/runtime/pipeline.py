"""WhisperPipeline -- waveform in, token ids out, everything on one B200: the flow of the reference's run.py:33-66
(load -> pad_or_trim -> log_mel_spectrogram -> WhisperEncoding.get_audio_features -> WhisperDecoding.main_loop) with
each stage replaced by this library's kernels:

    pad_or_trim + log-Mel (fp16)          whisper_utils.log_mel_spectrogram   b200_log_mel_spectrogram
    encoder                               runtime.WhisperEncoder              conv stem + 32 blocks on tcgen05
    CrossAttn_KV (int8 cross-KV caches)   WhisperDecoding.set_encoder_output  weight-only GEMMs + b200_cross_kv_pack
    greedy loop (+ logit filters)         WhisperDecoding.decode              one CUDA graph replay per token

Batches larger than the decoder's batch size run as consecutive slices; across GPUs the utterances are split with
runtime.sharding (one process per GPU, no collective until the final gather of token ids)."""
import numpy as np
import torch

from .. import whisper_utils
from .whisper_decoding import WhisperDecoding
from .whisper_encoder import WhisperEncoder


class WhisperPipeline:

    def __init__(self, dims, state_dict, batch_size, kv_scales, cross_kv_scales, device="cuda"):
        self.dims = dims
        self.B = batch_size
        self.device = torch.device(device)
        self.encoder = WhisperEncoder(dims, state_dict, device=device)
        self.decoder = WhisperDecoding(dims, state_dict, batch_size, kv_scales, cross_kv_scales, device=device)
        self.n_samples = 2 * dims.n_audio_ctx * whisper_utils.HOP_LENGTH  # conv2 has stride 2: 480000 for 1500 frames

    @classmethod
    def from_files(cls, checkpoint_path, quantize_dir, batch_size, device="cuda"):
        from .checkpoint import load_checkpoint, read_kv_scales
        dims, sd = load_checkpoint(checkpoint_path)
        return cls(dims, sd, batch_size, read_kv_scales(quantize_dir, dims.n_text_layer),
                   read_kv_scales(quantize_dir, dims.n_text_layer, cross=True), device=device)

    def log_mel(self, audio):
        """audio: [n] or [B, n] waveform(s), float32 at 16 kHz (numpy or torch) -> fp16 [B, n_mels, 2 * n_audio_ctx]."""
        if not torch.is_tensor(audio):
            audio = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
        if audio.dim() == 1:
            audio = audio.unsqueeze(0)
        audio = whisper_utils.pad_or_trim(audio.to(self.device), self.n_samples)
        return whisper_utils.log_mel_spectrogram(audio, self.dims.n_mels, dtype=torch.float16)

    def get_audio_features(self, mel):
        return self.encoder(mel)

    def transcribe_tokens(self, audio, prompt, n_new, use_graph=True):
        """-> int64 [n_utterances, n_new] greedy token ids (CPU).  prompt: list of token ids shared by all utterances
        (sot, language, task[, no_timestamps]; T/examples/whisper/decoding.py:314-319)."""
        mel = self.log_mel(audio)
        n = mel.shape[0]
        out = []
        for b0 in range(0, n, self.B):
            m = mel[b0:b0 + self.B]
            nb = m.shape[0]
            if nb < self.B:  # the decoder's batch is fixed by its CUDA graph: pad the last slice with its first row
                m = torch.cat([m, m[:1].expand(self.B - nb, -1, -1)], dim=0)
            xa = self.get_audio_features(m.contiguous())
            self.decoder.set_encoder_output(xa)
            tok = self.decoder.decode([list(prompt)] * self.B, n_new, use_graph=use_graph)
            out.append(tok[:nb].long().cpu())
        return torch.cat(out, dim=0)

    def enable_filters(self, tokenizer, suppress="-1", max_initial_timestamp=1.0):
        """The reference's default logit filters (decoding.py:332-348: SuppressBlank, SuppressTokens,
        ApplyTimestampRules with max_initial_timestamp 1.0 s) on the device, from a b200_whisper.tokenizer.Tokenizer.
        Without a vocabulary attached the blank is the multilingual id 220 and "-1" expands to the control tokens only."""
        has_vocab = tokenizer.encoding is not None
        blank = tokenizer.encode(" ")[0] if has_vocab else 220
        if not has_vocab and suppress == "-1":
            suppress = ""
        precision = 30.0 / self.dims.n_audio_ctx  # seconds per timestamp token: 0.02 (decoding.py:338-343)
        mi = None if max_initial_timestamp is None else int(round(max_initial_timestamp / precision))
        self.decoder.enable_logit_filters(tokenizer.eot, tokenizer.no_timestamps, tokenizer.timestamp_begin, blank,
                                          tokenizer.suppress_tokens(suppress), mi)
        self._filter_tokenizer = tokenizer

    def detect_language(self, audio_features, tokenizer):
        """decoding.py:703-741 for the decoder's batch: -> (language codes, [{code: probability}], no-speech
        probabilities).  Installs the cross-KV caches of `audio_features` in the decoder."""
        self.decoder.set_encoder_output(audio_features)
        toks = tokenizer.all_language_tokens
        lang, probs, nsp = self.decoder.detect_language(tokenizer.sot, toks[0], toks[-1] + 1, tokenizer.no_speech)
        codes = tokenizer.all_language_codes
        lang, probs, nsp = lang.cpu().tolist(), probs.cpu().tolist(), nsp.cpu().tolist()
        return [codes[t - toks[0]] for t in lang], [dict(zip(codes, row)) for row in probs], nsp

    def transcribe(self, audio, tokenizer, sample_len=None, detect_language=False):
        """run.py:57-66 for a batch of waveforms: greedy decode from the tokenizer's sot sequence with the logit filters
        on, tokens cut at the first end-of-text (decoding.py:836-840), text when the tokenizer has a vocabulary.
        -> list of {"tokens": [...], "text": str or None, "sum_logprob": float}; with detect_language also "language",
        "language_probs" and "no_speech_prob" (run.py:58), and -- the reference's default, DecodingOptions.language =
        None -- every utterance is decoded from ITS detected language token: decoding.py:738-739 writes it into
        tokens[:, sot_index + 1] before the main loop.  detect_language=False keeps the tokenizer's language (the
        reference with options.language set)."""
        if getattr(self, "_filter_tokenizer", None) is not tokenizer:
            self.enable_filters(tokenizer)
        prompt = list(tokenizer.sot_sequence)
        if sample_len is None:
            sample_len = self.dims.n_text_ctx // 2  # default of decoding.py:324
        sample_len = min(sample_len, self.dims.n_text_ctx - len(prompt))
        mel = self.log_mel(audio)
        results = []
        for b0 in range(0, mel.shape[0], self.B):
            m = mel[b0:b0 + self.B]
            nb = m.shape[0]
            if nb < self.B:
                m = torch.cat([m, m[:1].expand(self.B - nb, -1, -1)], dim=0)
            xa = self.get_audio_features(m.contiguous())
            extra = self.detect_language(xa, tokenizer) if detect_language else None
            if extra is None:
                self.decoder.set_encoder_output(xa)
            prompts = [list(prompt) for _ in range(self.B)]
            if extra is not None and len(prompt) > 1:
                lang_tok0 = tokenizer.all_language_tokens[0]
                codes = tokenizer.all_language_codes
                for i, code in enumerate(extra[0]):
                    prompts[i][1] = lang_tok0 + codes.index(code)
            tok = self.decoder.decode(prompts, sample_len).cpu().tolist()
            lp = self.decoder.logit_filter.sum_logprobs.cpu().tolist()
            for row, s in zip(tok[:nb], lp[:nb]):
                ids = row[:row.index(tokenizer.eot)] if tokenizer.eot in row else row
                text = tokenizer.decode(ids).strip() if tokenizer.encoding is not None else None
                results.append({"tokens": ids, "text": text, "sum_logprob": s})
                if extra is not None:
                    i = len(results) - 1 - b0
                    results[-1].update(language=extra[0][i], language_probs=extra[1][i], no_speech_prob=extra[2][i])
        return results
