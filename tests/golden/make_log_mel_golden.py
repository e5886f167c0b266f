"""Generates tests/golden/log_mel_golden.npz from the REFERENCE front end
(T/examples/whisper/whisper_utils.py:99-145 `log_mel_spectrogram`, imported from /root/reference -- build container
only; CPU torch.stft).  Inputs are the synthetic waveforms of log_mel_cases.py.  Stored per case: the reference output
(float32; the 30 s case keeps every 25th frame) and, once, the reference's mel filterbank asset digest.
Run:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_log_mel_golden.py
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/tensorrt_llm_july-release-v1/examples/whisper")

from log_mel_cases import cases  # noqa: E402


def main():
    import whisper_utils as wu  # the reference
    out = {}
    for name, (audio, padding) in cases().items():
        ref = wu.log_mel_spectrogram(torch.from_numpy(audio), padding=padding).numpy()
        assert ref.dtype == np.float32 and ref.shape == (80, (audio.shape[0] + padding) // 160)
        out[name] = ref[:, ::25].copy() if name == "full_30s" else ref
        print(name, ref.shape, float(ref.min()), float(ref.max()))
    filt = wu.mel_filters("cpu", 80).numpy()
    out["mel_filters_sha256"] = np.frombuffer(hashlib.sha256(filt.tobytes()).digest(), dtype=np.uint8)
    out["mel_filters_row_sums"] = filt.sum(axis=1)
    # pad_or_trim (whisper_utils.py:56-79)
    a = np.arange(10, dtype=np.float32)
    assert np.array_equal(wu.pad_or_trim(a, 4), a[:4]) and np.array_equal(wu.pad_or_trim(a, 12), np.pad(a, (0, 2)))
    np.savez_compressed(os.path.join(HERE, "log_mel_golden.npz"), **out)


if __name__ == "__main__":
    main()
