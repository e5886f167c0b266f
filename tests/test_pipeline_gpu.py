"""waveform -> tokens through WhisperPipeline (the flow of T/examples/whisper/run.py:33-66) on the GPU, stage by stage
against the oracle: log-Mel (oracle/log_mel.py), encoder and greedy decoder (oracle/whisper_oracle.py)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from log_mel_cases import speech_like  # noqa: E402

from oracle import log_mel as lm  # noqa: E402
from oracle import whisper_oracle as wo  # noqa: E402

pytestmark = pytest.mark.gpu


def test_waveform_to_tokens(tmp_path):
    from b200_whisper.runtime import WhisperPipeline, save_checkpoint, write_kv_scales
    dims = wo.MICRO
    B, n_new, prompt = 2, 6, [3, 7, 11]
    n = 2 * dims.n_audio_ctx * 160
    # three utterances of different lengths (pad_or_trim brings them to n samples), decoder batch 2 -> two slices
    audio = [speech_like(n - 3000, 41), speech_like(n, 42, amp=0.3), speech_like(n + 500, 43, amp=0.05)]
    batch = np.stack([lm.pad_or_trim(a, n) for a in audio])
    sd = wo.synthetic_state_dict(dims, seed=1)
    sdq = wo.quantize_state_dict(sd, dims, decoder_only=False)

    # oracle front end + encoder give the calibration scales (the files a user would bring)
    mel_ref = torch.from_numpy(lm.log_mel_spectrogram(batch)).half().float()
    with torch.no_grad():
        xa_ref = wo.encoder_forward(sdq, dims, mel_ref)
        kv_s, ckv_s = wo.calibrate_kv_scales(sdq, dims, xa_ref[:B], prompt, n_steps=4)
    save_checkpoint(str(tmp_path / "micro.pt"), dims, sd)
    write_kv_scales(str(tmp_path / "q"), kv_s, ckv_s)
    pipe = WhisperPipeline.from_files(str(tmp_path / "micro.pt"), str(tmp_path / "q"), batch_size=B)
    kv_s = [float(np.float32(s)) for s in kv_s]
    ckv_s = [float(np.float32(s)) for s in ckv_s]

    mel = pipe.log_mel(batch)
    assert mel.dtype == torch.float16 and tuple(mel.shape) == (3, 80, 2 * dims.n_audio_ctx)
    assert (mel.float().cpu() - mel_ref).abs().max().item() <= 1.5e-3          # one fp16 ulp at |x| < 2
    xa = pipe.get_audio_features(mel[:B].contiguous())
    assert (xa.float().cpu() - xa_ref[:B]).abs().max().item() <= 3e-2 * max(1.0, xa_ref.abs().max().item())

    tokens = pipe.transcribe_tokens(batch, prompt, n_new)
    assert tokens.dtype == torch.int64 and tuple(tokens.shape) == (3, n_new)
    # the oracle decodes from the pipeline's own encoder output, slice by slice (token identity then tests the decoder;
    # the stages in front of it are covered above)
    for b0 in (0, 2):
        m = mel[b0:b0 + B]
        if m.shape[0] < B:
            m = torch.cat([m, m[:1]], dim=0)
        xa_g = pipe.get_audio_features(m.contiguous()).float().cpu()
        with torch.no_grad():
            ref_tokens, _ = wo.greedy_decode(sdq, dims, xa_g, prompt, n_new, kv_s, ckv_s, act_fp16=True)
        nb = min(B, 3 - b0)
        assert tokens[b0:b0 + nb].tolist() == ref_tokens[:nb].tolist()
    # an utterance decodes to the same tokens whatever slice it lands in
    again = pipe.transcribe_tokens(batch[[2, 0, 1]], prompt, n_new)
    assert again.tolist() == tokens[[2, 0, 1]].tolist()


def test_transcribe_with_filters_and_tokenizer():
    """run.py's flow end to end with the logit filters on: prompt from the tokenizer, tokens cut at end-of-text, the
    timestamp rules visible in the output.  A miniature vocabulary (440 ranks + the 1608 special tokens = 2048)."""
    from b200_whisper.runtime import WhisperDecoding, WhisperPipeline
    from b200_whisper.tokenizer import Tokenizer
    dims = wo.ModelDimensions(80, 96, 128, 2, 2, 2048, 64, 128, 2, 2)
    tk = Tokenizer("en", "transcribe", n_ranks=2048 - 1608)
    assert tk.n_vocab == dims.n_vocab and tk.eot == 440 and tk.sot_sequence == (441, 442, 542)
    B, n = 2, 2 * dims.n_audio_ctx * 160
    sd = wo.synthetic_state_dict(dims, seed=2)
    scales = [0.05] * dims.n_text_layer
    pipe = WhisperPipeline(dims, sd, B, scales, scales)
    audio = np.stack([speech_like(n, 51), speech_like(n, 52, amp=0.4), speech_like(n, 53, amp=0.02)])
    res = pipe.transcribe(audio, tk, sample_len=20)
    assert len(res) == 3
    suppressed = set(tk.suppress_tokens(""))
    for r in res:
        ids = r["tokens"]
        assert 0 < len(ids) <= 20 and tk.eot not in ids and r["text"] is None and np.isfinite(r["sum_logprob"])
        assert not suppressed & set(ids) and tk.no_timestamps not in ids
        # ApplyTimestampRules: the first token is a timestamp no later than 1.0 s; timestamps never decrease
        assert tk.timestamp_begin <= ids[0] <= tk.timestamp_begin + 50
        ts = [t for t in ids if t >= tk.timestamp_begin]
        assert ts == sorted(ts)
    # the same decoder driven by hand gives the same tokens (the pipeline adds no arithmetic of its own)
    dec = pipe.decoder
    mel = pipe.log_mel(audio[:B])
    dec.set_encoder_output(pipe.get_audio_features(mel))
    rows = dec.decode([list(tk.sot_sequence)] * B, 20).cpu().tolist()
    for r, row in zip(res[:B], rows):
        assert r["tokens"] == (row[:row.index(tk.eot)] if tk.eot in row else row)
    assert isinstance(dec, WhisperDecoding)


def test_range_softmax_kernel_against_torch():
    """b200_logits_range_softmax vs the reference's own lines (decoding.py:721-725 and :765-766) evaluated with torch."""
    import b200_whisper
    from b200_whisper import _lib
    lib = b200_whisper.load()
    torch.manual_seed(3)
    rows, V, lo, hi, probe = 5, 51865, 50259, 50358, 50362
    logits = (torch.randn(rows, V) * 4).cuda()
    logits[2, lo + 7] = logits[2, lo + 40] = 30.0          # a tie: the first maximum wins, like torch.argmax
    lang = torch.empty(rows, dtype=torch.int32, device="cuda")
    probs = torch.empty(rows, hi - lo, device="cuda")
    nsp = torch.empty(rows, device="cuda")
    _lib.check(lib.b200_logits_range_softmax(logits.data_ptr(), rows, V, lo, hi, probe, lang.data_ptr(), probs.data_ptr(),
                                             nsp.data_ptr(), _lib.stream_ptr()))
    ref = logits.clone().cpu()
    mask = torch.ones(V, dtype=torch.bool)
    mask[lo:hi] = False
    ref[:, mask] = -float("inf")
    assert lang.cpu().tolist() == ref.argmax(dim=-1).tolist() and lang[2].item() == lo + 7
    torch.testing.assert_close(probs.cpu(), ref.softmax(dim=-1)[:, lo:hi], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(nsp.cpu(), logits.cpu().float().softmax(dim=-1)[:, probe], rtol=1e-4, atol=1e-9)
    assert abs(probs.sum(dim=-1).cpu() - 1).max().item() < 1e-5


def test_detect_language_and_no_speech():
    """WhisperPipeline.detect_language against the oracle decoder's logits for a lone sot token."""
    from b200_whisper.runtime import WhisperPipeline
    from b200_whisper.tokenizer import Tokenizer
    dims = wo.ModelDimensions(80, 96, 128, 2, 2, 2048, 64, 128, 2, 2)
    tk = Tokenizer("en", "transcribe", n_ranks=2048 - 1608)
    B, n = 2, 2 * dims.n_audio_ctx * 160
    sd = wo.synthetic_state_dict(dims, seed=2)
    sdq = wo.quantize_state_dict(sd, dims, decoder_only=False)
    scales = [0.05] * dims.n_text_layer
    pipe = WhisperPipeline(dims, sd, B, scales, scales)
    audio = np.stack([speech_like(n, 61), speech_like(n, 62, amp=0.4)])
    xa = pipe.get_audio_features(pipe.log_mel(audio))
    languages, probs, nsp = pipe.detect_language(xa, tk)
    with torch.no_grad():
        _, ref_logits = wo.greedy_decode(sdq, dims, xa.float().cpu(), [tk.sot], 1, scales, scales, act_fp16=True)
    ref = ref_logits[0].float()
    lo, hi = tk.all_language_tokens[0], tk.all_language_tokens[-1] + 1
    ref_p = ref[:, lo:hi].softmax(dim=-1)
    ref_nsp = ref.softmax(dim=-1)[:, tk.no_speech]
    for b in range(B):
        got = torch.tensor([probs[b][c] for c in tk.all_language_codes])
        assert (got - ref_p[b]).abs().max().item() <= 2e-2 * ref_p[b].max().item() + 1e-4
        top = ref_p[b].topk(2).values
        if (top[0] - top[1]).item() > 1e-2 * top[0].item():
            assert languages[b] == tk.all_language_codes[int(ref_p[b].argmax())]
        assert abs(nsp[b] - ref_nsp[b].item()) <= 2e-2 * ref_nsp[b].item() + 1e-6
    res = pipe.transcribe(audio, tk, sample_len=6, detect_language=True)
    assert [r["language"] for r in res] == languages and all(0.0 <= r["no_speech_prob"] <= 1.0 for r in res)
