"""Diagnostic: is waveform -> tokens bit-reproducible when the caching allocator hands out dirty memory?"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
import numpy as np
import torch
from log_mel_cases import speech_like

from oracle import log_mel as lm
from oracle import whisper_oracle as wo
from b200_whisper.runtime import WhisperPipeline

poison = os.environ.get("POISON", "1") == "1"
dims = wo.MICRO
B, n_new, prompt = 2, 6, [3, 7, 11]
n = 2 * dims.n_audio_ctx * 160
audio = [speech_like(n - 3000, 41), speech_like(n, 42, amp=0.3), speech_like(n + 500, 43, amp=0.05)]
batch = np.stack([lm.pad_or_trim(a, n) for a in audio])
sd = wo.synthetic_state_dict(dims, seed=1)
sdq = wo.quantize_state_dict(sd, dims, decoder_only=False)
mel_ref = torch.from_numpy(lm.log_mel_spectrogram(batch)).half().float()
with torch.no_grad():
    xa_ref = wo.encoder_forward(sdq, dims, mel_ref)
    kv_s, ckv_s = wo.calibrate_kv_scales(sdq, dims, xa_ref[:B], prompt, n_steps=4)


def dirty():
    if poison:
        for sz in (1 << 12, 1 << 16, 1 << 20, 1 << 24, 1 << 27):
            ts = [torch.full((sz,), 0xFF, dtype=torch.uint8, device="cuda") for _ in range(4)]
            del ts


dirty()
pipe = WhisperPipeline(dims, sd, B, kv_s, ckv_s)
runs = []
for r in range(3):
    dirty()
    mel = pipe.log_mel(batch)
    xa = pipe.get_audio_features(mel[:B].contiguous()).clone()
    tok = pipe.transcribe_tokens(batch, prompt, n_new)
    runs.append((mel.clone(), xa, tok))
    print("run", r, tok.tolist(), "xa finite", bool(torch.isfinite(xa).all()))
for r in (1, 2):
    print("mel equal", torch.equal(runs[0][0], runs[r][0]), "xa equal", torch.equal(runs[0][1], runs[r][1]),
          "tokens equal", torch.equal(runs[0][2], runs[r][2]))
with torch.no_grad():
    ref_tokens, ref_logits = wo.greedy_decode(sdq, dims, runs[0][1].float().cpu(), prompt, n_new, kv_s, ckv_s, act_fp16=True)
print("oracle", ref_tokens.tolist())
for t, l in enumerate(ref_logits):
    top = l.topk(2).values
    print("step", t, "margins", (top[:, 0] - top[:, 1]).tolist())
# decoder alone, repeated on the same encoder output
dec = pipe.decoder
for r in range(3):
    dirty()
    dec.set_encoder_output(runs[0][1])
    print("decode", r, dec.decode([prompt] * B, n_new).cpu().tolist())
    print("   eager", dec.decode([prompt] * B, n_new, use_graph=False).cpu().tolist())
