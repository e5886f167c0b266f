"""Times the GPU log-Mel front end (b200_log_mel_spectrogram) for a batch of 30 s utterances, beside the reference's
host path restated with torch.stft on the CPU cores (whisper_utils.py:99-145).  CUDA events, after warm-up."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

from b200_whisper.whisper_utils import log_mel_spectrogram, mel_filters

B = int(os.environ.get("BATCH", "16"))
torch.manual_seed(0)
audio = (0.1 * torch.randn(B, 480000)).cuda()
out = torch.empty((B, 80, 3000), dtype=torch.float16, device="cuda")
for _ in range(3):
    log_mel_spectrogram(audio, dtype=torch.float16, out=out)
torch.cuda.synchronize()
reps = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    log_mel_spectrogram(audio, dtype=torch.float16, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flop = B * 3000 * 201 * 400 * 4
print(f"gpu  batch {B}: {ms:7.3f} ms per call ({B / ms * 1e3:9.0f} utterances/s, {flop / ms / 1e9:6.1f} TFLOP/s fp32 DFT, "
      f"{B * 480000 * 4 / ms / 1e6:6.1f} GB/s of audio in)")
# the reference's host path, one utterance per call
a = audio[0].cpu()
filt = mel_filters("cpu", 80)
win = torch.hann_window(400)
t0 = time.perf_counter()
n = 4
for _ in range(n):
    stft = torch.stft(a, 400, 160, window=win, return_complex=True)
    mag = stft[..., :-1].abs() ** 2
    ls = torch.clamp(filt @ mag, min=1e-10).log10()
    ls = (torch.maximum(ls, ls.max() - 8.0) + 4.0) / 4.0
dt = (time.perf_counter() - t0) / n
print(f"cpu  torch.stft path, {torch.get_num_threads()} threads: {dt * 1e3:7.3f} ms per utterance ({1 / dt:7.1f} utterances/s)")
