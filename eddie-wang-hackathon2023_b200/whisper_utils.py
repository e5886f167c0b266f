"""Audio front end with the names and argument meaning of the reference's T/examples/whisper/whisper_utils.py
(`pad_or_trim` :56-79, `mel_filters` :81-97, `log_mel_spectrogram` :99-145); the spectrogram is computed on the GPU by
`b200_log_mel_spectrogram` (csrc/log_mel.cu) instead of torch.stft on the host.  There is no CPU path.

`load_audio` (:17-54) shells out to ffmpeg, which is neither in the reference tree nor in this image; waveforms are
passed as arrays / tensors (float32, 16 kHz, mono, the format `load_audio` returns).
"""
from functools import lru_cache
from typing import Optional, Union

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib

SAMPLE_RATE = 16000
N_FFT = 400
N_MELS = 80
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE  # 480000 samples in a 30-second chunk
N_FRAMES = N_SAMPLES // HOP_LENGTH  # 3000 frames in a mel spectrogram input


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """Pad (zeros on the right) or trim the audio array to `length` samples, as expected by the encoder."""
    if torch.is_tensor(array):
        if array.shape[axis] > length:
            array = array.narrow(axis, 0, length)
        if array.shape[axis] < length:
            pad = [0, 0] * array.ndim
            pad[2 * (array.ndim - 1 - (axis % array.ndim)) + 1] = length - array.shape[axis]
            array = F.pad(array, pad)
        return array
    array = np.asarray(array)
    if array.shape[axis] > length:
        array = array.take(indices=range(length), axis=axis)
    if array.shape[axis] < length:
        pad_widths = [(0, 0)] * array.ndim
        pad_widths[axis] = (0, length - array.shape[axis])
        array = np.pad(array, pad_widths)
    return array


def _mel_filterbank(n_mels: int, sr: int = SAMPLE_RATE, n_fft: int = N_FFT) -> np.ndarray:
    """librosa.filters.mel(sr=16000, n_fft=400, n_mels=80) -- the matrix the reference ships as
    assets/mel_filters.npz (whisper_utils.py:86-90 names the call).  Slaney mel scale (linear below 1 kHz, logarithmic
    above), triangular filters, Slaney area normalisation; float64 construction stored as float32.  Bit-identical with
    the reference asset (tests/test_log_mel_cpu.py)."""
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    top = min_log_mel + np.log((sr / 2.0) / min_log_hz) / logstep if sr / 2.0 >= min_log_hz else (sr / 2.0) / f_sp
    mels = np.linspace(0.0, top, n_mels + 2)
    hz = np.where(mels >= min_log_mel, min_log_hz * np.exp(logstep * (mels - min_log_mel)), f_sp * mels)
    bins = np.linspace(0, sr / 2, 1 + n_fft // 2)
    width = np.diff(hz)
    offset = hz[:, None] - bins[None, :]
    bank = np.zeros((n_mels, bins.size), dtype=np.float32)
    for m in range(n_mels):
        rising = -offset[m] / width[m]
        falling = offset[m + 2] / width[m + 1]
        bank[m] = np.maximum(0, np.minimum(rising, falling))
    bank *= (2.0 / (hz[2:] - hz[:-2]))[:, None]
    bank[bank == 0] = 0.0  # no negative zeros: bit-identical with the asset
    return bank


@lru_cache(maxsize=None)
def mel_filters(device, n_mels: int = N_MELS) -> torch.Tensor:
    """The mel filterbank matrix [n_mels, 201] for projecting the STFT power into a mel spectrogram."""
    assert n_mels == 80, f"Unsupported n_mels: {n_mels}"
    return torch.from_numpy(_mel_filterbank(n_mels)).to(device)


def log_mel_spectrogram(audio: Union[np.ndarray, torch.Tensor], n_mels: int = N_MELS, padding: int = 0,
                        device: Optional[Union[str, torch.device]] = None, dtype: torch.dtype = torch.float32,
                        out: Optional[torch.Tensor] = None):
    """Log-Mel spectrogram of a 16 kHz waveform.

    audio: shape (n,) -> (80, n_frames) like the reference, or (B, n) -> (B, 80, n_frames) with every utterance
    normalised by its own maximum (the reference handles one utterance per call).  padding: zero samples appended on the
    right.  device: CUDA device for the computation (default: the tensor's own, or the current CUDA device).
    dtype (B200 extension): torch.float32 (reference result) or torch.float16 (what `run.py:45` casts to).
    """
    if isinstance(audio, str):
        raise NotImplementedError("load_audio needs ffmpeg, which this image does not have: pass the waveform")
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(np.ascontiguousarray(audio))
    if device is not None:
        audio = audio.to(device)
    elif not audio.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("log_mel_spectrogram runs on the GPU (there is no CPU fallback)")
        audio = audio.cuda()
    if not audio.is_cuda:
        raise RuntimeError("log_mel_spectrogram runs on the GPU (there is no CPU fallback)")
    squeeze = audio.dim() == 1
    a = audio.reshape(-1, audio.shape[-1]).float().contiguous()
    B, n = a.shape
    lib = _lib.load()
    n_frames = lib.b200_log_mel_frames(n, padding)
    if dtype not in (torch.float32, torch.float16):
        raise ValueError("dtype must be torch.float32 or torch.float16")
    with torch.cuda.device(a.device):
        filters = mel_filters(a.device, n_mels)
        if out is None:
            out = torch.empty((B, n_mels, n_frames), dtype=dtype, device=a.device)
        else:
            assert out.is_contiguous() and out.dtype == dtype and out.numel() == B * n_mels * n_frames
        ws = torch.empty((lib.b200_log_mel_workspace_bytes(B, n, padding, n_mels),), dtype=torch.uint8, device=a.device)
        rc = lib.b200_log_mel_spectrogram(a.data_ptr(), B, n, padding, filters.data_ptr(), n_mels, out.data_ptr(),
                                          _lib.DTYPE_F16 if dtype == torch.float16 else _lib.DTYPE_F32,
                                          ws.data_ptr(), ws.numel(), _lib.stream_ptr())
        _lib.check(rc, "log_mel_spectrogram")
    return out[0] if squeeze else out
