"""Synthetic 16 kHz waveforms shared by make_log_mel_golden.py and the log-mel tests (no audio files offline).
numpy's PCG64 Generator streams are stable across numpy versions, so the cases regenerate identically everywhere."""
import numpy as np

SR = 16000


def speech_like(n, seed, amp=0.1):
    """A few drifting harmonics with a syllable-rate envelope over a noise floor: wide dynamic range across the 80 mel
    bands, so both the clamp at max - 8 and the unclamped region are exercised."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / SR
    f0 = 120.0 + 40.0 * np.sin(2 * np.pi * 0.7 * t + rng.uniform(0, 6.28))
    phase = 2 * np.pi * np.cumsum(f0) / SR
    x = np.zeros(n)
    for h in range(1, 24):
        x += (1.0 / h) * np.sin(h * phase + rng.uniform(0, 6.28))
    env = 0.5 * (1 + np.sin(2 * np.pi * 3.1 * t + rng.uniform(0, 6.28))) ** 2
    x = amp * env * x / np.abs(x).max() + 1e-3 * amp * rng.standard_normal(n)
    return x.astype(np.float32)


def cases():
    """name -> (audio float32 [n], padding)."""
    out = {}
    out["speech_1p5s"] = (speech_like(24000, 11), 0)
    burst = np.concatenate([speech_like(8000, 12, amp=0.8), np.zeros(24000, np.float32)])   # loud burst, then digital silence
    out["burst_then_silence"] = (burst, 0)
    out["ragged_padded"] = (speech_like(16000 + 37, 13), 123)                                 # n % 160 != 0, right padding
    out["noise_quiet"] = ((1e-4 * np.random.default_rng(14).standard_normal(4000)).astype(np.float32), 0)
    out["full_30s"] = (speech_like(480000, 15), 0)
    return out
