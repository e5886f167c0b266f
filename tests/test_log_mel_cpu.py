"""CPU tests of the log-Mel front end's checker and host logic:
  * the oracle (oracle/log_mel.py, float64) against golden outputs of the reference's own `log_mel_spectrogram`
    (T/examples/whisper/whisper_utils.py:99-145; tests/golden/make_log_mel_golden.py), and live against the reference
    module when /root/reference is present;
  * the mel filterbank built by the product (b200_whisper.whisper_utils._mel_filterbank) and by the oracle, bit-for-bit
    against the reference asset assets/mel_filters.npz (sha256 in the golden file; the file itself when present);
  * pad_or_trim, argument validation of the C entry point, no CPU fallback."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from log_mel_cases import cases  # noqa: E402

from oracle import log_mel as lm  # noqa: E402

REF_W = "/root/reference/tensorrt_llm_july-release-v1/examples/whisper"
# the reference computes the STFT with an fp32 FFT; against exact arithmetic its output moves by up to 3.6e-5
# (in (log10 + 4) / 4 units) on these cases
TOL_REF = 1e-4


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "log_mel_golden.npz"))


@pytest.mark.parametrize("name", ["speech_1p5s", "burst_then_silence", "ragged_padded", "noise_quiet", "full_30s"])
def test_oracle_matches_reference_golden(golden, name):
    audio, padding = cases()[name]
    out = lm.log_mel_spectrogram(audio, padding=padding)
    assert out.shape == (80, (audio.shape[0] + padding) // 160)
    if name == "full_30s":
        assert out.shape == (80, 3000)
        out = out[:, ::25]
    assert np.abs(out - golden[name]).max() <= TOL_REF
    # the floor at max - 8 decades is active in the cases built for it
    if name == "burst_then_silence":
        assert (golden[name] == golden[name].min()).mean() > 0.5
        assert np.isclose(golden[name].max() - golden[name].min(), 2.0, atol=1e-6)


def test_mel_filterbank_is_bit_identical_with_the_reference_asset(golden):
    from b200_whisper import whisper_utils as wu
    digest = bytes(golden["mel_filters_sha256"])
    for bank in (wu._mel_filterbank(80), lm.mel_filters(80)):
        assert bank.dtype == np.float32 and bank.shape == (80, 201)
        assert hashlib.sha256(bank.tobytes()).digest() == digest
        assert np.array_equal(bank.sum(axis=1), golden["mel_filters_row_sums"])
    assert int((wu._mel_filterbank(80) != 0).sum()) == 391   # what the kernel's zero-skipping relies on being small
    asset = os.path.join(REF_W, "assets", "mel_filters.npz")
    if os.path.exists(asset):
        assert np.array_equal(np.load(asset)["mel_80"], wu._mel_filterbank(80))
    f = wu.mel_filters("cpu", 80)
    assert f.dtype == torch.float32 and tuple(f.shape) == (80, 201)
    with pytest.raises(AssertionError):
        wu.mel_filters("cpu", 128)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_W, "whisper_utils.py")), reason="reference tree not present")
def test_oracle_matches_reference_live():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF_W)
    try:
        import whisper_utils as ref
    finally:
        sys.path.remove(REF_W)
    rng = np.random.default_rng(5)
    for n, padding in [(3200, 0), (16001, 0), (5000, 777), (201, 0)]:
        a = (0.3 * rng.standard_normal(n)).astype(np.float32)
        want = ref.log_mel_spectrogram(torch.from_numpy(a), padding=padding).numpy()
        got = lm.log_mel_spectrogram(a, padding=padding)
        assert got.shape == want.shape and np.abs(got - want).max() <= TOL_REF
    a = np.arange(10, dtype=np.float32)
    for length in (4, 10, 13):
        assert np.array_equal(ref.pad_or_trim(a, length), lm.pad_or_trim(a, length))


def test_pad_or_trim():
    from b200_whisper import whisper_utils as wu
    assert wu.N_SAMPLES == 480000 and wu.N_FRAMES == 3000 and wu.HOP_LENGTH == 160 and wu.N_FFT == 400
    a = np.arange(10, dtype=np.float32)
    assert np.array_equal(wu.pad_or_trim(a, 4), a[:4])
    assert np.array_equal(wu.pad_or_trim(a, 13), np.concatenate([a, np.zeros(3, np.float32)]))
    assert wu.pad_or_trim(a).shape == (480000,)
    t = torch.arange(12, dtype=torch.float32).reshape(2, 6)
    assert torch.equal(wu.pad_or_trim(t, 4), t[:, :4])
    assert torch.equal(wu.pad_or_trim(t, 8), torch.nn.functional.pad(t, (0, 2)))
    assert torch.equal(wu.pad_or_trim(t, 3, axis=0), torch.nn.functional.pad(t, (0, 0, 0, 1)))
    b = np.arange(12, dtype=np.float32).reshape(2, 6)
    assert np.array_equal(wu.pad_or_trim(b, 3, axis=0), np.pad(b, ((0, 1), (0, 0))))


def test_entry_point_validates_arguments_without_a_gpu():
    import b200_whisper
    lib = b200_whisper.load()
    assert lib.b200_log_mel_frames(480000, 0) == 3000 and lib.b200_log_mel_frames(16037, 123) == 101
    assert lib.b200_log_mel_workspace_bytes(16, 480000, 0, 80) == 256 + 16 * 80 * 3000 * 4
    assert lib.b200_log_mel_workspace_bytes(0, 480000, 0, 80) == 0
    x = torch.zeros(512)
    p = x.data_ptr()
    assert lib.b200_log_mel_spectrogram(None, 1, 512, 0, p, 80, p, 0, None, 0, None) == 1            # null audio
    assert lib.b200_log_mel_spectrogram(p, 1, 150, 50, p, 80, p, 0, None, 0, None) == 1              # <= 200 samples
    assert b"reflect" in lib.b200_last_error()
    assert lib.b200_log_mel_spectrogram(p, 1, 512, 0, p, 80, p, 2, None, 0, None) == 1               # int8 output
    assert lib.b200_log_mel_spectrogram(p, 1, 512, 0, p, 80, p, 0, None, 0, None) == 4               # no workspace
    if not torch.cuda.is_available():
        from b200_whisper import whisper_utils as wu
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            wu.log_mel_spectrogram(np.zeros(16000, np.float32))
