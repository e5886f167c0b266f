"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's log-Mel front end.  Only tests/, smoke() and the
cpu_baseline leg of bench.py may import this; the product path (csrc/log_mel.cu) never does.

Restates T/examples/whisper/whisper_utils.py:
  * `pad_or_trim`            :56-79
  * `mel_filters`            :81-97   (the asset assets/mel_filters.npz = librosa.filters.mel(sr=16000, n_fft=400,
                                       n_mels=80); librosa is a third-party dependency that is not vendored, so its
                                       published Slaney-scale construction is restated here and checked bit-for-bit
                                       against the asset, tests/test_log_mel_cpu.py)
  * `log_mel_spectrogram`    :99-145  (torch.stft with a periodic Hann window, centre = True / reflect padding, power,
                                       mel projection, log10 clamp, max - 8 floor, (x + 4) / 4)

The arithmetic is float64 numpy (a direct DFT through numpy's rfft), i.e. the exact value the reference's fp32 FFT
approximates.  Pinned by tests/golden/log_mel_golden.npz, generated from the reference module itself
(tests/golden/make_log_mel_golden.py).
"""
import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
N_MELS = 80
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE
N_BINS = N_FFT // 2 + 1


def pad_or_trim(array, length=N_SAMPLES):
    """whisper_utils.py:56-79 (last axis)."""
    array = np.asarray(array)
    if array.shape[-1] > length:
        array = array[..., :length]
    if array.shape[-1] < length:
        pad = [(0, 0)] * array.ndim
        pad[-1] = (0, length - array.shape[-1])
        array = np.pad(array, pad)
    return array


def mel_filters(n_mels=N_MELS, sr=SAMPLE_RATE, n_fft=N_FFT):
    """librosa.filters.mel(sr, n_fft, n_mels) with its defaults (fmin 0, fmax sr/2, Slaney mel scale, Slaney area
    normalisation, float32 result): the matrix stored in the reference's assets/mel_filters.npz
    (whisper_utils.py:81-97)."""
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fftfreqs = np.linspace(0, sr / 2, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    w = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None]
    w[w == 0] = 0.0  # one ramp product is -0.0; the asset holds +0.0 there
    return w


def hann_window(n=N_FFT):
    """torch.hann_window(n) (periodic)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def log_mel_spectrogram(audio, n_mels=N_MELS, padding=0, filters=None):
    """whisper_utils.py:99-145 for one utterance (1-D audio) or a batch [B, n] treated utterance by utterance (the
    reference is only ever called with one utterance, run.py:44-46).  Returns float64 [..., n_mels, n // 160]."""
    audio = np.asarray(audio, dtype=np.float64)
    if audio.ndim == 2:
        return np.stack([log_mel_spectrogram(a, n_mels, padding, filters) for a in audio])
    if padding > 0:
        audio = np.pad(audio, (0, padding))
    n = audio.shape[0]
    assert n > N_FFT // 2, "reflect padding needs more than n_fft / 2 samples"
    x = np.pad(audio, (N_FFT // 2, N_FFT // 2), mode="reflect")       # torch.stft(center=True, pad_mode='reflect')
    n_frames = 1 + n // HOP_LENGTH
    idx = np.arange(N_FFT)[None, :] + HOP_LENGTH * np.arange(n_frames)[:, None]
    frames = x[idx] * hann_window()[None, :]
    stft = np.fft.rfft(frames, axis=-1).T                               # [201, n_frames]
    magnitudes = np.abs(stft[:, :-1]) ** 2                              # drop the last frame (:136)
    if filters is None:
        filters = mel_filters(n_mels)
    mel_spec = np.asarray(filters, dtype=np.float64) @ magnitudes
    log_spec = np.log10(np.maximum(mel_spec, 1e-10))
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    return (log_spec + 4.0) / 4.0
